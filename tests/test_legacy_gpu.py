"""Legacy 13-cell TD critic (model/ffm_ac_core.py) on the GPU, through the C ABI, against fixtures written
by the UNMODIFIED reference under keyed draws and against the NumPy restatement."""
import os
import pickle

import numpy as np
import pytest

from helpers import LEGACY_AC_FIXTURES, MARGIN_GUARD, load_legacy, random_positions
from oracle import assets, legacy_numpy
from oracle.inject import PhiloxSource

pytestmark = pytest.mark.gpu


def _pack(pos, cap):
    buf = np.full((1, cap, 2), -1, np.int32)
    buf[0, :len(pos)] = pos
    return buf, np.array([len(pos)], np.int32)


@pytest.mark.parametrize("name", LEGACY_AC_FIXTURES)
def test_ac_fixture_bit_exact(name):
    from ffm_b200.legacy import LegacySim
    g = load_legacy(name)
    assert float(np.min(g["min_margin"])) > MARGIN_GUARD
    H, W = g["map"].shape
    cap = max(len(e["pos0"]) for e in g["ep"])
    sim = LegacySim(g["map"], g["sff"], 1, cap, model="ac", learn="exact", params=g["params"], seed=int(g["seed"]))
    T = int(g["max_steps"])
    for ep, e in enumerate(g["ep"]):
        sim.set_episode_base(ep)
        sim.set_positions(*_pack(e["pos0"], cap))
        sim.set_dff(np.zeros((1, H, W), np.float32))
        cells, cnt = sim.rollout(T, record=T)
        steps = int(sim.counters()[0][0])
        assert steps == len(e["traj"])
        for t in range(steps):
            ref = e["traj"][t][:, 0] * W + e["traj"][t][:, 1]
            assert cnt[0, t] == len(ref), (ep, t)
            assert np.array_equal(cells[0, t, :cnt[0, t]].astype(np.int64), ref), (ep, t)
        if ep == int(g["set_v_after"]):                     # get_v_table() -> set_v_table(): unseen states read -1.0 from now on
            k, r = sim.get_table("V")
            sim.set_table(k, r, "V", default=-1.0)
    keys, rows = sim.get_table("V")
    order = np.argsort(keys)
    assert np.array_equal(keys[order], g["v_keys"])
    assert np.array_equal(rows[order, 0], g["v_vals"])      # bit for bit
    assert sim.table_size("V") == len(g["v_keys"])
    assert np.array_equal(sim.get_dff()[0], g["final_dff"])


def test_ac_random_configurations_vs_oracle():
    """Random rooms with obstacles, both neighbourhoods and SFF dtypes, random parameters: trajectories, V and DFF bit for bit."""
    from ffm_b200.legacy import LegacySim
    rng = np.random.RandomState(77)
    checked = 0
    for case in range(10):
        h, w = int(rng.randint(8, 22)), int(rng.randint(8, 22))
        m = assets.room_map(h, w)
        m[rng.randint(2, h - 2, 6), rng.randint(1, w - 1, 6)] = 2          # obstacles
        if case % 3 == 0:
            m[h - 1, w // 3] = 3                                            # a second exit
        dtype = np.float32 if case % 2 == 0 else np.float64
        sff = assets.sff_norm_min(m, ["L1", "L2", "Linf"][case % 3], dtype)
        params = {"neighborhood": "moore" if case % 2 else "neumann", "block_size": int(rng.randint(1, 6)),
                  "k_S": float(rng.uniform(1, 10)), "k_D": float(rng.uniform(0, 2)), "alpha_v": float(rng.uniform(0.05, 0.5)),
                  "gamma": float(rng.uniform(0.8, 1.0)), "step_penalty": float(-rng.uniform(0, 1)),
                  "collision_penalty": float(-rng.uniform(0, 3)), "diffuse": float(rng.uniform(0.05, 0.4)), "decay": float(rng.uniform(0.05, 0.4))}
        n = int(rng.randint(1, max(2, int((m == 0).sum() * 0.6))))
        seed = 1000 + case
        sim = LegacySim(m, sff, 1, n, model="ac", learn="exact", params=params, seed=seed)
        V, bad = {}, False
        for ep in range(2):
            pos0 = random_positions(m, n, rng)
            o = legacy_numpy.AcOracle(m, sff, pos0, params, PhiloxSource(seed, ep), v_table=V)
            traj = o.run(max_steps=80)
            V = o.V
            sim.set_episode_base(ep)
            sim.set_positions(*_pack(pos0, n))
            sim.set_dff(np.zeros((1, h, w), np.float32))
            cells, cnt = sim.rollout(80, record=80)
            if o.min_margin < MARGIN_GUARD:
                bad = True                                                  # a draw on a CDF boundary: excused (helpers.MARGIN_GUARD)
                break
            assert int(sim.counters()[0][0]) == len(traj)
            for t, p in enumerate(traj):
                assert np.array_equal(cells[0, t, :cnt[0, t]].astype(np.int64), p[:, 0] * w + p[:, 1]), (case, ep, t)
            assert np.array_equal(sim.get_dff()[0], o.dff), (case, ep)
        if bad:
            continue
        keys, rows = sim.get_table("V")
        order = np.argsort(keys)
        ok = np.array(sorted(V), np.uint64)
        assert np.array_equal(keys[order], ok), case
        assert np.array_equal(rows[order, 0], np.array([V[int(k)] for k in ok])), case
        checked += 1
    assert checked >= 8


def test_ac_frozen_batch_matches_single_episodes():
    """learn="none": many episodes per launch; V never steers the moves, so every episode equals its oracle run."""
    from ffm_b200.legacy import LegacySim
    m = assets.room_map(14, 14)
    sff = assets.sff_norm_min(m, "L1", np.float32)
    rng = np.random.RandomState(5)
    B, n = 12, 40
    pos = [random_positions(m, n - e, rng) for e in range(B)]                 # ragged
    buf = np.full((B, n, 2), -1, np.int32)
    for e in range(B):
        buf[e, :len(pos[e])] = pos[e]
    sim = LegacySim(m, sff, B, n, model="ac", learn="none", params={"neighborhood": "moore"}, seed=9, episode_base=100)
    sim.set_positions(buf, np.array([len(p) for p in pos], np.int32))
    cells, cnt = sim.rollout(100, record=100)
    steps, ped_steps = sim.counters()
    assert sim.table_size("V") == 0
    for e in range(B):
        o = legacy_numpy.AcOracle(m, sff, pos[e], {"neighborhood": "moore"}, PhiloxSource(9, 100 + e))
        traj = o.run(max_steps=100)
        if o.min_margin < MARGIN_GUARD:
            continue
        assert steps[e] == len(traj)
        assert ped_steps[e] == len(pos[e]) + sum(len(p) for p in traj[:-1])
        for t, p in enumerate(traj):
            assert np.array_equal(cells[e, t, :cnt[e, t]].astype(np.int64), p[:, 0] * 14 + p[:, 1]), (e, t)


def test_ac_dropin_interface(tmp_path):
    """The drop-in class: constructor, step / run / reset, table accessors, .npz output (ffm_ac_core.py:320-390)."""
    from ffm_b200.model.ffm_ac_core import FloorFieldModel
    m = assets.room_map(12, 12)
    sff = assets.sff_norm_min(m, "L1", np.float32)
    p = os.path.join(tmp_path, "sff.npy")
    np.save(p, sff)
    params = {"neighborhood": "neumann", "seed": 4242}
    np.random.seed(3)
    model = FloorFieldModel(m, p, 25, params)
    pos0 = np.array(model.positions)
    assert pos0.shape == (25, 2) and model.dff.shape == (12, 12) and model.get_v_table_size() == 0
    o = legacy_numpy.AcOracle(m, sff, pos0, params, PhiloxSource(4242, 0))
    for t in range(5):
        model.step()
        o.step()
        assert np.array_equal(model.positions, o.positions)
        assert np.array_equal(model.dff, o.dff)
    assert model.run(max_steps=23) == 23
    for _ in range(23):
        o.step()
    assert np.array_equal(model.positions, o.positions)
    vt = model.get_v_table()
    nby = (12 + 3 - 1) // 3
    assert {legacy_numpy.state_to_key(pickle.loads(k), nby): v for k, v in vt.items()} == o.V
    assert model.get_v_table_size() == len(o.V) == len(model.V)
    model.set_v_table(vt)
    assert model.V[b"unseen"] == -1.0                                           # ffm_ac_core.py:345
    np.random.seed(8)
    model.reset()
    assert model.positions.shape == (25, 2) and not model.dff.any() and model.get_v_table_size() == len(o.V)
    o2 = legacy_numpy.AcOracle(m, sff, np.array(model.positions), params, PhiloxSource(4242, 1), v_table=o.V, v_default=-1.0)
    assert model.run() == len(o2.run())
    vt2 = model.get_v_table()
    assert {legacy_numpy.state_to_key(pickle.loads(k), nby): v for k, v in vt2.items()} == o2.V
    model.update_dff()
    # .npz output of run(save_prefix=...) (:378-388); like the reference's np.array(buffer) it needs a constant head count
    # per file, so: three pedestrians far from the exit, six steps
    big = assets.room_map(20, 20)
    p2 = os.path.join(tmp_path, "sff20.npy")
    np.save(p2, assets.sff_norm_min(big, "L1", np.float32))
    np.random.seed(1)
    walker = FloorFieldModel(big, p2, 3, {"seed": 7})
    walker.positions = np.array([[15, 5], [16, 10], [17, 15]])
    assert walker.run(save_prefix=os.path.join(tmp_path, "ep"), save_interval=4, max_steps=6) == 6
    a = np.load(os.path.join(tmp_path, "ep_4.npz"))["positions"]
    b = np.load(os.path.join(tmp_path, "ep_final.npz"))["positions"]
    assert a.shape == (4, 3, 2) and a.dtype == np.int32 and b.shape == (2, 3, 2)
    assert np.array_equal(b[-1], walker.positions)


def test_ac_errors():
    from ffm_b200.legacy import LegacySim
    m = assets.room_map(8, 8)
    sff = assets.sff_norm_min(m, "L1", np.float32)
    with pytest.raises(ValueError):
        LegacySim(m, sff, 2, 4, model="ac", learn="exact")                      # sequential learning is one episode
    sim = LegacySim(m, sff, 1, 4, model="ac")
    with pytest.raises(ValueError):
        sim.set_positions(np.array([[[0, 0], [1, 1], [-1, -1], [-1, -1]]], np.int32), [2])   # a wall cell
    sim.set_positions(np.array([[[1, 1], [1, 1], [-1, -1], [-1, -1]]], np.int32), [2])
    with pytest.raises(ValueError):
        sim.rollout(1)                                                           # two pedestrians on one cell
    bad = m.copy(); bad[0, 1] = 0
    with pytest.raises(ValueError):
        LegacySim(bad, sff, 1, 4, model="ac")


# ---- model/ffm_actor_only.py -------------------------------------------------------------------------------------------------
from helpers import LEGACY_ACTOR_FIXTURES  # noqa: E402


def _sorted_table(sim, which):
    keys, rows = sim.get_table(which)
    order = np.argsort(keys)
    return keys[order], rows[order]


@pytest.mark.parametrize("name", LEGACY_ACTOR_FIXTURES)
def test_actor_only_fixture_bit_exact(name):
    from ffm_b200.legacy import LegacySim
    g = load_legacy(name)
    assert float(np.min(g["min_margin"])) > MARGIN_GUARD
    H, W = g["map"].shape
    cap = max(len(e["pos0"]) for e in g["ep"])
    sim = LegacySim(g["map"], g["sff"], 1, cap, model="actor_only", learn="exact", params=g["params"], seed=int(g["seed"]))
    sim.set_epsilon(float(g["eps"]))
    T = int(g["max_steps"])
    for ep, e in enumerate(g["ep"]):
        sim.set_episode_base(ep)
        sim.set_positions(*_pack(e["pos0"], cap))
        sim.set_dff(np.zeros((1, H, W), np.float32))
        cells, cnt = sim.rollout(T, record=T)
        steps = int(sim.counters()[0][0])
        assert steps == len(e["traj"])
        for t in range(steps):
            ref = e["traj"][t][:, 0] * W + e["traj"][t][:, 1]
            assert cnt[0, t] == len(ref), (ep, t)
            assert np.array_equal(cells[0, t, :cnt[0, t]].astype(np.int64), ref), (ep, t)
    vk, vr = _sorted_table(sim, "V")
    hk, hr = _sorted_table(sim, "H")
    assert np.array_equal(vk, g["v_keys"]) and np.array_equal(vr[:, 0], g["v_vals"])
    assert np.array_equal(hk, g["h_keys"]) and np.array_equal(hr, g["h_vals"])
    assert np.array_equal(sim.get_dff()[0], g["final_dff"])


def test_actor_only_random_configurations_vs_oracle():
    """Random rooms, both neighbourhoods, random parameters and exploration rates, tables carried over episodes and passed
    through table_get / table_set in between: trajectories, V, H and DFF bit for bit."""
    from ffm_b200.legacy import LegacySim
    rng = np.random.RandomState(99)
    checked = 0
    for case in range(6):
        h, w = int(rng.randint(8, 16)), int(rng.randint(8, 16))
        m = assets.room_map(h, w)
        m[rng.randint(2, h - 2, 4), rng.randint(1, w - 1, 4)] = 2
        if case % 3 == 1:
            m[h - 1, w // 3] = 3
        sff = assets.sff_norm_min(m, "L1", np.float32)
        params = {"neighborhood": "moore" if case % 2 else "neumann", "k_A": float(rng.uniform(1, 10)), "k_D": float(rng.uniform(0, 2)),
                  "alpha_v": float(rng.uniform(0.05, 0.5)), "alpha_h": float(rng.uniform(0.05, 0.5)), "gamma": float(rng.uniform(0.8, 1.0)),
                  "step_penalty": float(-rng.uniform(0, 1)), "collision_penalty": float(-rng.uniform(0, 3))}
        eps = [0.0, 0.1, 0.3][case % 3]
        n = int(rng.randint(1, max(2, int((m == 0).sum() * 0.4))))
        seed = 2000 + case
        sim = LegacySim(m, sff, 1, n, model="actor_only", learn="exact", params=params, seed=seed)
        sim.set_epsilon(eps)
        V, Ht, bad = {}, {}, False
        for ep in range(3):
            pos0 = random_positions(m, n, rng)
            o = legacy_numpy.ActorOnlyOracle(m, sff, pos0, params, PhiloxSource(seed, ep), v_table=V, h_table=Ht, epsilon=eps)
            traj = o.run(max_steps=30)
            V, Ht = o.V, o.H
            sim.set_episode_base(ep)
            sim.set_positions(*_pack(pos0, n))
            sim.set_dff(np.zeros((1, h, w), np.float32))
            cells, cnt = sim.rollout(30, record=30)
            if o.min_margin < MARGIN_GUARD:
                bad = True
                break
            assert int(sim.counters()[0][0]) == len(traj)
            for t, p in enumerate(traj):
                assert np.array_equal(cells[0, t, :cnt[0, t]].astype(np.int64), p[:, 0] * w + p[:, 1]), (case, ep, t)
            assert np.array_equal(sim.get_dff()[0], o.dff), (case, ep)
            if ep == 0:                                          # round trip of both tables through the host
                for which in ("V", "H"):
                    k, r = sim.get_table(which)
                    sim.set_table(k, r, which)
        if bad:
            continue
        vk, vr = _sorted_table(sim, "V")
        hk, hr = _sorted_table(sim, "H")
        ok = np.array(sorted(V), np.uint64)
        assert np.array_equal(vk, ok) and np.array_equal(vr[:, 0], np.array([V[int(k)] for k in ok])), case
        ok = np.array(sorted(Ht), np.uint64)
        assert np.array_equal(hk, ok) and np.array_equal(hr, np.array([Ht[int(k)] for k in ok]).reshape(len(ok), -1)), case
        checked += 1
    assert checked >= 4


def test_actor_only_dropin_interface(tmp_path):
    from ffm_b200.model.ffm_actor_only import FloorFieldModelActorOnly
    m = assets.room_map(12, 12)
    sff = assets.sff_norm_min(m, "L1", np.float32)
    p = os.path.join(tmp_path, "sff.npy")
    np.save(p, sff)
    # a pretrained critic table in the format run_critic_training.py pickles (bytes keys): loaded, counted, never read
    vp = os.path.join(tmp_path, "v.pkl")
    with open(vp, "wb") as f:
        pickle.dump({pickle.dumps(((0,) * 13, (1, 1))): 3.5, pickle.dumps(((2,) * 13, (0, 0))): -1.0}, f)
    params = {"neighborhood": "neumann", "seed": 777}
    np.random.seed(3)
    model = FloorFieldModelActorOnly(m, p, 10, pretrained_v_path=vp, params=params)
    model.set_epsilon(0.1)
    assert model.initial_v_size == 2 and model.get_v_table_size() == (2, 2, 0) and model.get_h_table_size() == (0, 0)
    pos0 = np.array(model.positions)
    o = legacy_numpy.ActorOnlyOracle(m, sff, pos0, params, PhiloxSource(777, 0), epsilon=0.1)
    for t in range(4):
        model.step()
        o.step()
        assert np.array_equal(model.positions, o.positions)
        assert np.array_equal(model.dff, o.dff)
    steps, traj = model.run(max_steps=20, return_trajectory=True)
    ot = [o.step() or o.positions.copy() for _ in range(steps)]
    assert steps == 20 and len(traj) == 20
    for a, b in zip(traj, ot):
        assert np.array_equal(np.asarray(a, dtype=np.int64).reshape(-1, 2), b)
    nby = (12 + 4) // 5
    ht = model.get_h_table()
    assert {legacy_numpy.state_to_key(pickle.loads(k), nby): v for k, v in ht.items()} == o.H
    vt = model.get_v_table()
    dev = {legacy_numpy.state_to_key(pickle.loads(k), nby): v for k, v in vt.items() if isinstance(k, bytes)}
    assert dev == o.V and len(vt) == len(o.V) + 2
    assert model.get_v_table_size() == (2, len(o.V) + 2, len(o.V))
    assert model.get_h_table_size() == (len(o.H), len(o.H) * 5)
    assert model.H[b"unseen"] == [] and model.V[b"unseen"] == 0.0
    np.random.seed(4)
    model.reset()
    assert not model.dff.any() and model.get_h_table_size()[0] == len(o.H)


def test_actor_only_frozen_batch():
    """learn="none": a batch of episodes with frozen tables; rows that do not exist read as zeros and are not inserted."""
    from ffm_b200.legacy import LegacySim
    m = assets.room_map(12, 12)
    sff = assets.sff_norm_min(m, "L1", np.float32)
    rng = np.random.RandomState(6)
    # train one episode, then evaluate a batch with the learned tables
    tr = LegacySim(m, sff, 1, 20, model="actor_only", learn="exact", seed=3)
    tr.set_epsilon(0.2)
    tr.set_positions(*_pack(random_positions(m, 20, rng), 20))
    tr.rollout(60)
    B, n = 16, 20
    buf = np.stack([np.concatenate([random_positions(m, n, rng)]) for _ in range(B)]).astype(np.int32)
    ev = LegacySim(m, sff, B, n, model="actor_only", learn="none", seed=11)
    for which in ("V", "H"):
        k, r = tr.get_table(which)
        ev.set_table(k, r, which)
    sizes = (ev.table_size("V"), ev.table_size("H"))
    ev.set_positions(buf, np.full((B,), n, np.int32))
    ev.rollout(400)
    steps, ped_steps = ev.counters()
    pos, left = ev.get_positions()
    assert (ev.table_size("V"), ev.table_size("H")) == sizes
    assert (left >= 0).all() and (steps > 0).all() and (ped_steps >= n).all()
    ev2 = LegacySim(m, sff, B, n, model="actor_only", learn="none", seed=11)          # run to run identical
    for which in ("V", "H"):
        k, r = tr.get_table(which)
        ev2.set_table(k, r, which)
    ev2.set_positions(buf, np.full((B,), n, np.int32))
    ev2.rollout(400)
    assert np.array_equal(ev2.counters()[0], steps) and np.array_equal(ev2.get_positions()[0], pos)


def test_legacy_dropins_driven_like_the_training_scripts(tmp_path):
    """The loops of run_critic_training.py:125-226 (N patterns x episodes on ONE model, V pickled at the end) and of
    run_actor_only_training.py:151-299 (pretrained V, epsilon ramp, `model.N = N`, a trajectory .npz every few episodes,
    H pickled per N), written out here with the drop-in classes in place of the reference's."""
    from ffm_b200.model.ffm_ac_core import FloorFieldModel
    from ffm_b200.model.ffm_actor_only import FloorFieldModelActorOnly
    m = assets.room_map(12, 12)
    p = os.path.join(tmp_path, "sff.npy")
    np.save(p, assets.sff_norm_min(m, "L1", np.float64))
    params = {"k_S": 10, "k_D": 1, "alpha_v": 0.01, "gamma": 0.99, "exit_reward": 100.0, "step_penalty": -1.0,
              "collision_penalty": -1.0, "neighborhood": "neumann", "block_size": 5}          # run_critic_training.py:34-44
    np.random.seed(0)
    model = FloorFieldModel(map_array=m, sff_path=p, N=1, params=params)
    sizes = []
    for N in (1, 10, 30):
        for episode in range(3):
            model.N = N
            model.reset()
            steps = model.run(max_steps=500)
            assert 0 < steps <= 500 and model.positions.shape[0] == 0                         # everybody out well before the cap
            sizes.append(model.get_v_table_size())
    assert sizes == sorted(sizes) and sizes[-1] > 100                                         # the shared table only grows
    v_table = model.get_v_table()
    vp = os.path.join(tmp_path, "V_integrated_total9ep.pkl")
    with open(vp, "wb") as f:
        pickle.dump(v_table, f)
    vals = np.array(list(v_table.values()))
    assert np.isfinite(vals).all() and vals.max() <= 100.0 + 1e-9 and vals.max() > 0.9        # bounded by the exit reward

    actor = FloorFieldModelActorOnly(map_array=m, sff_path=p, N=1, pretrained_v_path=vp,
                                     params={**params, "k_A": 10, "alpha_h": 0.1})
    assert actor.initial_v_size == len(v_table)
    n_list, per_n = (1, 5), 6
    total, cur = len(n_list) * per_n, 0
    for N in n_list:
        actor.N = N
        for episode in range(1, per_n + 1):
            cur += 1
            actor.set_epsilon(float(np.clip(0.5 + (0.01 - 0.5) * (cur - 1) / (total - 1), 0.0, 1.0)))
            actor.reset()
            if episode % 3 == 0:
                steps, trajectory = actor.run(max_steps=200, return_trajectory=True)
                fn = os.path.join(tmp_path, f"trajectory_N{N}_ep{episode:05d}_total{cur:05d}.npz")
                np.savez_compressed(fn, positions=trajectory, episode=episode, N=N, total_episode=cur, steps=steps)
                z = np.load(fn, allow_pickle=True)
                assert len(z["positions"]) == steps and z["positions"][0].shape[1] == 2
            else:
                steps = actor.run(max_steps=200, return_trajectory=False)
            assert 0 < steps <= 200
            v0, v1, vnew = actor.get_v_table_size()
            assert v0 == len(v_table) and v1 == v0 + vnew
            h_states, h_actions = actor.get_h_table_size()
            assert h_actions == 5 * h_states
        with open(os.path.join(tmp_path, f"H_actor_N{N}_total{per_n}ep.pkl"), "wb") as f:
            pickle.dump(actor.get_h_table(), f)
    assert actor.get_h_table_size()[0] > 10 and abs(actor.epsilon - 0.01) < 1e-12
