"""Generates tests/golden/* from the UNMODIFIED reference (container only; needs /root/reference).

    python -m oracle.make_golden

Fixtures
  core_*.npz          full episodes of model/ffm_core.py FloorFieldModel under keyed Philox draws
                      (oracle/inject.py): inputs, per-step positions, per-step DFF, step count
  stock_main_seed42.npz  the stock `main.py` run (config/default_config.yaml, seed 42, the
                      reference's own MT19937 streams) with every consumed draw recorded, so the run
                      can be replayed through the injected-draw buffers of the CUDA path
  shipped_assets.json sha256 of the arrays the reference ships (data/maps, data/sff)
"""
import hashlib
import json
import os
import random
import sys
import tempfile

import numpy as np

from . import assets, inject

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
REF = inject.REFERENCE_ROOT


def _flatten(traj):
    counts = np.array([len(p) for p in traj], dtype=np.int32)
    flat = np.concatenate([p.reshape(-1, 2) for p in traj]).astype(np.int16) if traj else np.zeros((0, 2), np.int16)
    return flat, counts


def core_case(name, h, w, N, nbh, metric, dtype, seed, episode, extra=None, dff_every=1):
    ref = inject.import_reference("ffm_core")
    m = assets.room_map(h, w)
    sff = assets.sff_norm_min(m, metric, dtype)
    params = {"neighborhood": nbh, **(extra or {})}
    with tempfile.TemporaryDirectory() as tmp:
        p = os.path.join(tmp, "sff.npy")
        np.save(p, sff)
        np.random.seed(seed)
        model = ref.FloorFieldModel(m, p, N, params)
        pos0 = np.array(model.positions, dtype=np.int16)
        r = inject.run_reference(model, inject.PhiloxSource(seed, episode), log_probs=True)
    flat, counts = _flatten(r["traj"])
    dff = np.stack(r["dff"][::dff_every]).astype(np.float32)
    # probabilities of the first 64 draws, for the 1e-6 relative probability bar
    pl = r["probs_log"][:64]
    probs = np.zeros((len(pl), 9), np.float64)
    probs_meta = np.zeros((len(pl), 3), np.int32)
    for i, (t, idx, pr) in enumerate(pl):
        probs[i, :len(pr)] = pr
        probs_meta[i] = (t, idx, len(pr))
    np.savez_compressed(
        os.path.join(OUT, name + ".npz"), map=m, sff=sff, pos0=pos0, params=json.dumps(params),
        seed=np.uint64(seed), episode=np.uint32(episode), steps=np.int32(r["steps"]), traj=flat,
        traj_counts=counts, dff=dff, dff_every=np.int32(dff_every), min_margin=np.float64(r["min_margin"]),
        n_move=np.int32(r["n_move"]), n_coin=np.int32(r["n_coin"]), n_winner=np.int32(r["n_winner"]),
        probs=probs, probs_meta=probs_meta)
    print(name, "steps", r["steps"], "min_margin %.2e" % r["min_margin"], "draws", r["n_move"], r["n_coin"], r["n_winner"])


def stock_main():
    """main.py:17-46 with its own generators; records the draws it consumes."""
    import yaml
    ref = inject.import_reference("ffm_core")
    with open(os.path.join(REF, "config", "default_config.yaml")) as f:
        config = yaml.safe_load(f)
    np.random.seed(config["seed"])
    random.seed(config["seed"])
    m = np.load(os.path.join(REF, config["map"]))
    model = ref.FloorFieldModel(m, os.path.join(REF, config["sff"]), config["N"], config["params"])
    pos0 = np.array(model.positions, dtype=np.int16)
    H, W = m.shape
    rec = inject.RecordingSource(inject.StockSource(), 600, config["N"], H * W)
    r = inject.run_reference(model, rec, keep_dff=False)
    T = r["steps"]
    cf = rec.conflict_buf[:T]
    tt, cc = np.nonzero(~np.isnan(cf[:, :, 0]))
    flat, counts = _flatten(r["traj"])
    np.savez_compressed(
        os.path.join(OUT, "stock_main_seed42.npz"), params=json.dumps(config["params"]), pos0=pos0,
        steps=np.int32(T), traj=flat, traj_counts=counts, move=rec.move_buf[:T],
        conflict_t=tt.astype(np.int32), conflict_cell=cc.astype(np.int32), conflict_u=cf[tt, cc],
        min_margin=np.float64(r["min_margin"]), final_dff=np.array(model.dff, dtype=np.float32))
    print("stock main.py seed 42:", T, "steps,", int(counts[:-1].sum()) + config["N"], "ped-steps, min_margin %.2e" % r["min_margin"])


UNI_PARAMS = dict(k_S=10, k_D=1, k_A=10, alpha_v=0.01, alpha_h=0.1, gamma=0.99, exit_reward=100.0, step_penalty=-1.0,
                  collision_penalty=-1.0, neighborhood="neumann", block_size=1)   # run_unified_*_training.py MODEL_PARAMS


def _tables_to_arrays(o_like, vdict, hdict, A):
    """dict tables -> (ids, values) arrays using the dense state id of oracle/unified_numpy.py."""
    vid = np.array(sorted(o_like.key_to_id(k) for k in vdict), np.int64)
    vval = np.array([vdict[o_like.id_to_key(i)] for i in vid], np.float64)
    hid = np.array(sorted(o_like.key_to_id(k) for k in hdict), np.int64) if hdict else np.zeros(0, np.int64)
    hval = np.array([hdict[o_like.id_to_key(i)] for i in hid], np.float64).reshape(-1, A)
    return vid, vval, hid, hval


def unified_case(name, mode, h, w, N, episodes, seed, params, eps=0.0, radius=None, max_steps=300, sff_dtype=np.float32,
                 v_from=None):
    """Multi-episode learning run of the reference FloorFieldModelUnified under keyed draws: tables carry
    over the episodes exactly as in run_unified_critic_training.py:216-222."""
    import contextlib, io, pickle
    from . import unified_numpy
    ref = inject.import_reference("ffm_unified")
    m = assets.room_map(h, w)
    sff = assets.sff_norm_min(m, "L1", sff_dtype)
    vtab = None
    with tempfile.TemporaryDirectory() as tmp:
        p = os.path.join(tmp, "sff.npy")
        np.save(p, sff)
        vpath = None
        if v_from is not None:
            z = np.load(os.path.join(OUT, v_from + ".npz"))
            helper = unified_numpy.UnifiedOracle(m, sff, np.zeros((0, 2)), "critic_only", params)
            vtab = {helper.id_to_key(i): float(v) for i, v in zip(z["v_ids"], z["v_vals"])}
            vpath = os.path.join(tmp, "v.pkl")
            with open(vpath, "wb") as f:      # the reference expects pickled-bytes keys (ffm_unified.py:91-107)
                pickle.dump({pickle.dumps(k): v for k, v in vtab.items()}, f)
        np.random.seed(seed)
        with contextlib.redirect_stdout(io.StringIO()):
            model = ref.FloorFieldModelUnified(m, p, N, learning_mode=mode, pretrained_v_path=vpath, params=params)
        if mode != "critic_only":
            model.set_epsilon(eps)
        exit_pos = tuple(int(v) for v in np.argwhere(m == 3)[0])
        pos0s, trajs, counts, steps, margins = [], [], [], [], []
        for ep in range(episodes):
            model.reset(exit_pos=exit_pos if radius else None, radius=radius)
            pos0s.append(np.array(model.positions, dtype=np.int16).reshape(-1, 2))
            r = inject.run_reference(model, inject.PhiloxSource(seed, ep), max_steps=max_steps, keep_dff=False)
            flat, cnt = _flatten(r["traj"])
            trajs.append(flat); counts.append(cnt); steps.append(r["steps"]); margins.append(r["min_margin"])
    helper = unified_numpy.UnifiedOracle(m, sff, np.zeros((0, 2)), mode, params)
    A = helper.A
    vid, vval, hid, hval = _tables_to_arrays(helper, dict(model.V), dict(model.H) if model.H is not None else {}, A)
    save = dict(map=m, sff=sff, params=json.dumps(params), mode=mode, seed=np.uint64(seed), eps=np.float64(eps),
                max_steps=np.int32(max_steps), episodes=np.int32(episodes), steps=np.array(steps, np.int32),
                min_margin=np.array(margins), v_ids=vid, v_vals=vval, h_ids=hid, h_vals=hval,
                final_dff=np.array(model.dff, np.float32), v_from=v_from or "")
    for ep in range(episodes):
        save[f"pos0_{ep}"] = pos0s[ep]; save[f"traj_{ep}"] = trajs[ep]; save[f"counts_{ep}"] = counts[ep]
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **save)
    print(name, "steps", steps, "min_margin %.1e" % min(margins), "|V|", len(vid), "|H|", len(hid))


def trained_case(name, h, w, N, seed, params, h_from, max_steps=120):
    """model/ffm_trained_core.py FloorFieldModel with the H table learned in fixture `h_from`."""
    import contextlib, io, pickle
    from . import unified_numpy
    ref = inject.import_reference("ffm_trained_core")
    m = assets.room_map(h, w)
    sff = assets.sff_norm_min(m, "L1", np.float32)
    z = np.load(os.path.join(OUT, h_from + ".npz"))
    helper = unified_numpy.UnifiedOracle(m, sff, np.zeros((0, 2)), "trained", params)
    htab = {helper.id_to_key(i): [float(x) for x in row] for i, row in zip(z["h_ids"], z["h_vals"])}
    with tempfile.TemporaryDirectory() as tmp:
        p = os.path.join(tmp, "sff.npy"); np.save(p, sff)
        hp = os.path.join(tmp, "h.pkl")
        with open(hp, "wb") as f:
            pickle.dump({pickle.dumps(k): v for k, v in htab.items()}, f)   # ffm_trained_core.py:55-59
        np.random.seed(seed)
        with contextlib.redirect_stdout(io.StringIO()):
            model = ref.FloorFieldModel(m, p, N, hp, params)
        pos0 = np.array(model.positions, dtype=np.int16)
        r = inject.run_reference(model, inject.PhiloxSource(seed, 0), max_steps=max_steps, keep_dff=False)
    flat, cnt = _flatten(r["traj"])
    np.savez_compressed(os.path.join(OUT, name + ".npz"), map=m, sff=sff, params=json.dumps(params), seed=np.uint64(seed),
                        h_from=h_from, pos0=pos0, traj=flat, counts=cnt, steps=np.int32(r["steps"]), max_steps=np.int32(max_steps),
                        min_margin=np.float64(r["min_margin"]), final_dff=np.array(model.dff, np.float32))
    print(name, "steps", r["steps"], "min_margin %.1e" % r["min_margin"])


def unified_all():
    unified_case("uni_critic_12x12", "critic_only", 12, 12, 30, 5, 21, UNI_PARAMS, radius=7)
    unified_case("uni_critic_moore_bs5", "critic_only", 12, 12, 45, 3, 22, {**UNI_PARAMS, "neighborhood": "moore", "block_size": 5})
    unified_case("uni_critic_20x20_f64", "critic_only", 20, 20, 40, 2, 23, {**UNI_PARAMS, "block_size": 3}, sff_dtype=np.float64)
    unified_case("uni_actor_eps", "actor_only", 12, 12, 16, 3, 24, UNI_PARAMS, eps=0.2, max_steps=60, v_from="uni_critic_12x12")
    unified_case("uni_actor_greedy", "actor_only", 12, 12, 16, 2, 25, UNI_PARAMS, eps=0.0, max_steps=60)
    unified_case("uni_both", "both", 12, 12, 20, 3, 26, UNI_PARAMS, eps=0.1, max_steps=60)
    unified_case("uni_both_moore", "both", 12, 12, 20, 2, 27, {**UNI_PARAMS, "neighborhood": "moore", "block_size": 2}, eps=0.05, max_steps=50)
    trained_case("trained_12x12", 12, 12, 30, 28, {"k_D": 1, "k_A": 10, "neighborhood": "neumann", "block_size": 1}, "uni_actor_eps")


def mcq_case(name, h, w, N, seed, params, betas, sff_dtype=np.float64):
    """Multi-episode run of model/ffm_learning_core.py (shared Q across episodes like main_learning.py:79-106)."""
    from . import mcq_numpy
    ref = inject.import_reference("ffm_learning_core")
    m = assets.room_map(h, w)
    sff = assets.sff_norm_min(m, "L1", sff_dtype)
    with tempfile.TemporaryDirectory() as tmp:
        p = os.path.join(tmp, "sff.npy")
        np.save(p, sff)
        np.random.seed(seed)
        model = ref.FloorFieldModel(m, p, N, params)
    Q, save, steps, margins = {}, {}, [], []
    for ep, beta in enumerate(betas):
        model.Q = Q
        model.reset()
        save[f"pos0_{ep}"] = np.array(model.positions, dtype=np.int16)
        traj = []
        with inject.injected(inject.PhiloxSource(seed, ep), w) as st:
            t = 0
            while model.positions.shape[0] > 0:
                st.step = t
                model.step(beta)
                traj.append(np.array(model.positions, dtype=np.int64).reshape(-1, 2))
                t += 1
        flat, cnt = _flatten(traj)
        save[f"traj_{ep}"], save[f"counts_{ep}"] = flat, cnt
        steps.append(t); margins.append(st.min_margin)
        Q = model.Q
    helper = mcq_numpy.McqOracle(m, sff, np.zeros((0, 2)), params)
    ids = np.array(sorted(helper.id_of(k) for k in Q), np.int64)
    rows = np.stack([Q[helper.key_of(i)] for i in ids]).astype(np.float32)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), map=m, sff=sff, params=json.dumps(params), seed=np.uint64(seed),
                        betas=np.array(betas, np.float64), alpha=np.float64(model.alpha), gamma=np.float64(model.gamma),
                        steps=np.array(steps, np.int32), min_margin=np.array(margins), q_ids=ids, q_rows=rows,
                        final_dff=np.array(model.dff, np.float32), **save)
    print(name, "steps", steps, "min_margin %.1e" % min(margins), "|Q|", len(ids))


def pretrain_case(name, h, w, seed, params, shuffle_seed=5):
    """coverage_pretrain_empty of the UNMODIFIED run_coverage_pretrain_and_training.py (:173-216) on a small room: the
    module's own loop, shuffles and force_first_step_and_roll; only the draw call sites are keyed (mini-episode k =
    Philox episode k, CA step = the model's own _step_count) and the order of the patterns is logged."""
    import importlib
    import random as pyrandom
    from . import mcq_numpy
    if inject.REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, inject.REFERENCE_ROOT)
    drv = importlib.import_module("run_coverage_pretrain_and_training")
    cls = drv.FloorFieldModel
    m = assets.room_map(h, w)
    sff = assets.sff_norm_min(m, "L1", np.float64)
    order, steps = [], []
    with tempfile.TemporaryDirectory() as tmp:
        p = os.path.join(tmp, "sff.npy")
        np.save(p, sff)
        with inject.injected(inject.PhiloxSource(seed, 0), w) as st:
            o_step, o_force = cls.step, drv.force_first_step_and_roll

            def step(self, beta):
                st.step = self._step_count              # the CA step about to run (incremented inside, :158)
                steps[-1] += 1
                return o_step(self, beta)

            def force(map_array, sff_path, params, shared_Q, T, from_dir, step_buffer=10, save_episode_path=None):
                st.source = inject.PhiloxSource(seed, len(order))
                order.append((int(T[0]), int(T[1]), int(from_dir)))
                steps.append(0)
                return o_force(map_array=map_array, sff_path=sff_path, params=params, shared_Q=shared_Q, T=T, from_dir=from_dir,
                               step_buffer=step_buffer, save_episode_path=save_episode_path)

            cls.step, drv.force_first_step_and_roll = step, force
            try:
                np.random.seed(seed)
                pyrandom.seed(shuffle_seed)
                Q = {}
                drv.coverage_pretrain_empty(map_array=m, sff_path=p, params=dict(params), shared_Q=Q, shuffle=True, save_dir=None)
            finally:
                cls.step, drv.force_first_step_and_roll = o_step, o_force
    helper = mcq_numpy.McqOracle(m, sff, np.zeros((0, 2)), params)
    ids = np.array(sorted(helper.id_of(k) for k in Q), np.int64)
    rows = np.stack([Q[helper.key_of(i)] for i in ids]).astype(np.float32)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), map=m, sff=sff, params=json.dumps(params), seed=np.uint64(seed),
                        order=np.array(order, np.int32), steps=np.array(steps, np.int32), q_ids=ids, q_rows=rows,
                        alpha=np.float64(params.get("alpha", 0.1)), gamma=np.float64(params.get("gamma", 0.99)),
                        min_margin=np.float64(st.min_margin))
    print(name, "patterns", len(order), "CA steps", int(np.sum(steps)), "|Q|", len(ids), "min_margin %.1e" % st.min_margin)


def mcq_all():
    pretrain_case("mcq_pretrain_12x12", 12, 12, 41, {"max_steps": 500, "alpha": 0.1, "gamma": 0.99, "step_penalty": 0.01, "stop_penalty": 0.3})
    pretrain_case("mcq_pretrain_9x14", 9, 14, 42, {"max_steps": 500, "alpha": 0.2, "gamma": 0.9, "timeout_penalty": 20.0})
    mcq_case("mcq_12x12_penalties", 12, 12, 20, 31, {"max_steps": 60, "step_penalty": 0.01, "stop_penalty": 0.3, "collision_penalty": 0.7},
             [1.0, 0.6, 0.2, 0.0], np.float32)
    mcq_case("mcq_12x12_default", 12, 12, 40, 32, {"max_steps": 500}, [1.0, 0.5, 0.1])
    mcq_case("mcq_20x20_kq", 20, 20, 30, 33, {"max_steps": 80, "k_Q": 2.0, "step_penalty": 0.05}, [0.8, 0.3])


def legacy_ac_case(name, h, w, N, episodes, seed, params, sff_dtype=np.float32, metric="L1", max_steps=400, set_v_after=None):
    """Multi-episode run of the legacy TD critic model/ffm_ac_core.py under keyed draws: V carries over the episodes like
    run_critic_training.py does (reset() between episodes).  `set_v_after` = episode index after which the table is passed
    through get_v_table() / set_v_table(), which switches the default of unseen states to -1.0 (ffm_ac_core.py:340)."""
    import pickle
    from . import legacy_numpy
    ref = inject.import_reference("ffm_ac_core")
    m = assets.room_map(h, w)
    sff = assets.sff_norm_min(m, metric, sff_dtype)
    with tempfile.TemporaryDirectory() as tmp:
        p = os.path.join(tmp, "sff.npy")
        np.save(p, sff)
        np.random.seed(seed)
        model = ref.FloorFieldModel(m, p, N, params)
        pos0s, trajs, counts, steps, margins = [], [], [], [], []
        for ep in range(episodes):
            if ep > 0:
                model.reset()
            pos0s.append(np.array(model.positions, dtype=np.int16).reshape(-1, 2))
            r = inject.run_reference(model, inject.PhiloxSource(seed, ep), max_steps=max_steps, keep_dff=False)
            flat, cnt = _flatten(r["traj"])
            trajs.append(flat); counts.append(cnt); steps.append(r["steps"]); margins.append(r["min_margin"])
            if set_v_after is not None and ep == set_v_after:
                model.set_v_table(model.get_v_table())
    bs = model.block_size
    nby = (w + bs - 1) // bs
    tab = {legacy_numpy.state_to_key(pickle.loads(k), nby): float(v) for k, v in model.V.items()}
    keys = np.array(sorted(tab), np.uint64)
    vals = np.array([tab[int(k)] for k in keys], np.float64)
    save = dict(map=m, sff=sff, params=json.dumps(params), seed=np.uint64(seed), max_steps=np.int32(max_steps),
                episodes=np.int32(episodes), steps=np.array(steps, np.int32), min_margin=np.array(margins), v_keys=keys, v_vals=vals,
                final_dff=np.array(model.dff, np.float32), set_v_after=np.int32(-1 if set_v_after is None else set_v_after))
    for ep in range(episodes):
        save[f"pos0_{ep}"] = pos0s[ep]; save[f"traj_{ep}"] = trajs[ep]; save[f"counts_{ep}"] = counts[ep]
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **save)
    print(name, "steps", steps, "min_margin %.1e" % min(margins), "|V|", len(keys))


def legacy_actor_case(name, h, w, N, episodes, seed, params, eps=0.0, max_steps=60, obstacles=()):
    """Multi-episode run of the legacy actor model/ffm_actor_only.py under keyed draws (entity = idx * 8 + i: the class decides
    once per neighbour slot, ffm_actor_only.py:214-355); V and H carry over the episodes like run_actor_only_training.py."""
    import contextlib, io, pickle
    from . import legacy_numpy
    ref = inject.import_reference("ffm_actor_only")
    m = assets.room_map(h, w)
    for r, c in obstacles:
        m[r, c] = 2
    sff = assets.sff_norm_min(m, "L1", np.float32)
    with tempfile.TemporaryDirectory() as tmp:
        p = os.path.join(tmp, "sff.npy")
        np.save(p, sff)
        np.random.seed(seed)
        with contextlib.redirect_stdout(io.StringIO()):
            model = ref.FloorFieldModelActorOnly(m, p, N, params=params)
        model.set_epsilon(eps)
        pos0s, trajs, counts, steps, margins = [], [], [], [], []
        for ep in range(episodes):
            if ep > 0:
                model.reset()
            pos0s.append(np.array(model.positions, dtype=np.int16).reshape(-1, 2))
            r = inject.run_reference(model, inject.PhiloxSource(seed, ep), max_steps=max_steps, keep_dff=False, sub_key=True)
            flat, cnt = _flatten(r["traj"])
            trajs.append(flat); counts.append(cnt); steps.append(r["steps"]); margins.append(r["min_margin"])
    nby = (w + 4) // 5
    vt = {legacy_numpy.state_to_key(pickle.loads(k), nby): float(v) for k, v in model.V.items()}
    ht = {legacy_numpy.state_to_key(pickle.loads(k), nby): [float(x) for x in v] for k, v in model.H.items()}
    vk = np.array(sorted(vt), np.uint64)
    hk = np.array(sorted(ht), np.uint64)
    save = dict(map=m, sff=sff, params=json.dumps(params), seed=np.uint64(seed), eps=np.float64(eps), max_steps=np.int32(max_steps),
                episodes=np.int32(episodes), steps=np.array(steps, np.int32), min_margin=np.array(margins),
                v_keys=vk, v_vals=np.array([vt[int(k)] for k in vk], np.float64),
                h_keys=hk, h_vals=np.array([ht[int(k)] for k in hk], np.float64).reshape(len(hk), -1),
                final_dff=np.array(model.dff, np.float32), set_v_after=np.int32(-1))
    for ep in range(episodes):
        save[f"pos0_{ep}"] = pos0s[ep]; save[f"traj_{ep}"] = trajs[ep]; save[f"counts_{ep}"] = counts[ep]
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **save)
    print(name, "steps", steps, "min_margin %.1e" % min(margins), "|V|", len(vk), "|H|", len(hk), "draws", r["n_move"], r["n_eps"], r["n_winner"])


def legacy_all():
    legacy_ac_case("legacy_ac_12x12", 12, 12, 30, 4, 51, {"neighborhood": "neumann"})
    legacy_ac_case("legacy_ac_moore_f64", 16, 20, 60, 3, 52, {"neighborhood": "moore", "block_size": 5, "k_S": 3, "k_D": 0.5, "gamma": 0.9,
                                                             "alpha_v": 0.2, "step_penalty": -1.0, "collision_penalty": -2.5},
                   sff_dtype=np.float64, metric="L2", set_v_after=0)
    legacy_ac_case("legacy_ac_12x12_full", 12, 12, 100, 2, 53, {"neighborhood": "neumann", "block_size": 1}, max_steps=120)
    legacy_actor_case("legacy_actor_12x12", 12, 12, 12, 4, 66, {"neighborhood": "neumann"}, eps=0.0, max_steps=50)
    legacy_actor_case("legacy_actor_eps_moore", 14, 16, 24, 3, 62, {"neighborhood": "moore", "k_A": 4, "k_D": 0.5, "alpha_h": 0.3, "alpha_v": 0.2,
                                                                   "gamma": 0.9, "step_penalty": -1.0, "collision_penalty": -0.5},
                      eps=0.15, max_steps=40, obstacles=((5, 5), (5, 6), (8, 10)))
    legacy_actor_case("legacy_actor_crowded", 10, 10, 40, 2, 63, {"neighborhood": "neumann", "step_penalty": -1.0}, eps=0.05, max_steps=40)


def shipped():
    out = {}
    for rel in ("data/maps/simple_room.npy", "data/sff/distance_L1.npy", "data/sff/distance_L2.npy", "data/sff/distance_Linf.npy"):
        a = np.load(os.path.join(REF, rel))
        out[rel] = dict(shape=list(a.shape), dtype=str(a.dtype), sha256=hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest(),
                        n_inf=int(np.isinf(a).sum()) if a.dtype.kind == "f" else 0)
    with open(os.path.join(OUT, "shipped_assets.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    print("shipped assets:", {k: v["sha256"][:12] for k, v in out.items()})


def main():
    import sys
    os.makedirs(OUT, exist_ok=True)
    only = [a for a in sys.argv[1:] if not a.startswith("-")]
    if only:                      # python -m oracle.make_golden core_50x50_moore_f64 ...: regenerate the named core fixtures
        for name in only:
            if name == "legacy":
                legacy_all()
            elif name == "core_50x50_moore_f64":
                core_case(name, 50, 50, 100, "moore", "L2", np.float64, 13, 5, dff_every=20)
            else:
                raise SystemExit(f"no single-fixture recipe for {name}; run without arguments")
        return
    core_case("core_12x12_neumann_f32", 12, 12, 50, "neumann", "L1", np.float32, 11, 3)
    core_case("core_12x12_moore_f32_full", 12, 12, 100, "moore", "L1", np.float32, 12, 0)
    core_case("core_50x50_moore_f64", 50, 50, 100, "moore", "L2", np.float64, 13, 5, dff_every=20)
    core_case("core_50x50_neumann_f64", 50, 50, 100, "neumann", "L1", np.float64, 14, 1, dff_every=20)
    core_case("core_20x20_moore_params", 20, 20, 150, "moore", "Linf", np.float32, 15, 9,
              extra={"k_S": 2.5, "k_D": 0.7, "diffuse": 0.3, "decay": 0.1}, dff_every=10)
    core_case("core_16x24_moore_kd0", 16, 24, 60, "moore", "L2", np.float32, 16, 2, extra={"k_D": 0}, dff_every=10)
    stock_main()
    shipped()
    unified_all()
    mcq_all()
    legacy_all()


if __name__ == "__main__":
    main()
