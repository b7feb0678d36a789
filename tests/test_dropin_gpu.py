"""The drop-in class (ffm_b200/model/ffm_core.py) driven the way the reference's main.py drives
model/ffm_core.py (main.py:39-46), compared with the oracle step by step."""
import os

import numpy as np
import pytest

from helpers import MARGIN_GUARD
from oracle import assets, ffm_numpy
from oracle.inject import PhiloxSource

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("nbh,dtype,hw,N", [("neumann", np.float32, 12, 50), ("moore", np.float64, 50, 100)])
def test_main_py_loop(cuda_device, tmp_path, nbh, dtype, hw, N):
    from ffm_b200.model.ffm_core import FloorFieldModel

    m = assets.room_map(hw, hw).astype(np.int64)          # np.load(config["map"]) is int64 (main.py:34)
    sff = assets.sff_norm_min(m, "L1", dtype)
    p = os.path.join(tmp_path, "sff.npy")
    np.save(p, sff)
    np.random.seed(42)
    model = FloorFieldModel(m, p, N, {"neighborhood": nbh, "seed": 777, "k_A": 3, "max_steps": 500})
    assert model.positions.shape == (N, 2) and model.positions.dtype == np.int64
    assert model.dff.shape == m.shape and model.dff.dtype == np.float32 and not model.dff.any()
    assert model.map_array.dtype == np.uint8 and len(model.neighbors) == (4 if nbh == "neumann" else 8)
    # same global-generator placement as the reference constructor (ffm_core.py:24-25)
    np.random.seed(42)
    free = np.argwhere(m == 0)
    assert np.array_equal(model.positions, free[np.random.choice(len(free), N, replace=False)])

    o = ffm_numpy.CoreOracle(m, sff, model.positions, model.params, PhiloxSource(777, 0))
    log, step = [], 0
    while model.positions.shape[0] > 0:                     # main.py:44-46
        model.step()
        log.append(np.copy(model.positions))
        step += 1
        o.step()
        if o.min_margin >= MARGIN_GUARD:
            assert np.array_equal(log[-1], o.positions), f"step {step}"
            if step % 25 == 0:
                assert np.array_equal(model.dff.view(np.uint32), o.dff.view(np.uint32))
    assert o.positions.shape[0] == 0 or o.min_margin < MARGIN_GUARD


def test_run_and_attribute_assignment(cuda_device, tmp_path):
    from ffm_b200.model.ffm_core import FloorFieldModel

    m = assets.room_map(12, 12)
    sff = assets.sff_norm_min(m, "L1", np.float32)
    p = os.path.join(tmp_path, "sff.npy")
    np.save(p, sff)
    model = FloorFieldModel(m, p, 30, {"seed": 5})
    model.step()
    # run_trained_ffm.py:235-236 assigns positions and dff
    model.positions = np.array([[5, 5], [6, 6], [3, 2]])
    model.dff = np.zeros_like(model.map_array, dtype=np.float32)
    assert model.positions.shape == (3, 2) and not model.dff.any()
    model.update_dff()
    assert model.run() is None and model.positions.shape[0] == 0
    with pytest.raises(ValueError):
        model.positions = np.array([[0, 0]])                # a wall
    with pytest.raises(ValueError):
        FloorFieldModel(m, p, 500, {"seed": 5})            # more pedestrians than free cells (ffm_core.py:25)
    with pytest.raises(FileNotFoundError):
        FloorFieldModel(m, os.path.join(tmp_path, "missing.npy"), 3)


@pytest.mark.parametrize("h,w,nbh", [(12, 12, "neumann"), (10, 14, "moore"), (16, 36, "moore"), (9, 64, "neumann")])
def test_standalone_update_dff_runs_on_the_device_bit_exact(cuda_device, tmp_path, h, w, nbh):
    """update_dff() outside step() (ffm_core.py:106-117): the library's stencil, not a host restatement; widths that
    take the vectorised (w % 4 == 0, w >= 32) and the scalar walk."""
    from ffm_b200.model.ffm_core import FloorFieldModel
    from oracle import ffm_numpy

    m = assets.room_map(h, w)
    p = os.path.join(tmp_path, "sff.npy")
    np.save(p, assets.sff_norm_min(m, "L1", np.float32))
    model = FloorFieldModel(m, p, 5, {"seed": 1, "neighborhood": nbh, "diffuse": 0.3, "decay": 0.15})
    rng = np.random.RandomState(3)
    d = (rng.rand(h, w) * (rng.rand(h, w) < 0.4) * 3).astype(np.float32)
    d[2, 3] = 1.2e-4                                         # decays below the 1e-4 cut
    model.dff = d
    want = d.copy()
    for _ in range(3):
        model.update_dff()
        want = ffm_numpy.update_dff(want, model.params, model.neighbors)
        assert np.array_equal(model.dff.view(np.uint32), want.view(np.uint32))


def test_unified_dropin_like_training_driver(cuda_device, tmp_path):
    """run_unified_critic_training.py:164-225 / run_unified_actor_training.py:193-268 usage pattern: one
    model object, `.N = n`, reset(exit_pos, radius), run(max_steps), table accessors; compared with
    the oracle driven through the same episodes."""
    import pickle
    from ffm_b200.model.ffm_unified import FloorFieldModelUnified
    from oracle import unified_numpy

    m = assets.room_map(12, 12)
    sff = assets.sff_norm_min(m, "L1", np.float32)
    p = os.path.join(tmp_path, "sff.npy")
    np.save(p, sff)
    params = dict(k_S=10, k_D=1, k_A=10, alpha_v=0.01, alpha_h=0.1, gamma=0.99, exit_reward=100.0, step_penalty=-1.0,
                  collision_penalty=-1.0, neighborhood="neumann", block_size=1, seed=4242)
    with pytest.raises(ValueError):
        FloorFieldModelUnified(m, p, 5, learning_mode="bogus", params=params)
    np.random.seed(3)
    model = FloorFieldModelUnified(m, p, 10, learning_mode="critic_only", params=params)
    o = unified_numpy.UnifiedOracle(m, sff, np.zeros((0, 2)), "critic_only", params)
    ep = 1                                            # episode 0 was the constructor's placement
    for N in (1, 10, 25):
        model.N = N
        for _ in range(2):
            model.reset(exit_pos=(0, 6), radius=7)
            o.positions, o.t, o.source = model.positions.copy(), 0, PhiloxSource(4242, ep)
            o.dff[:] = 0
            steps = model.run(max_steps=300)
            r = o.run(max_steps=300)
            ep += 1
            assert steps == r["steps"] and model.positions.shape[0] == 0
    assert model.get_v_table_size() == int(o.v_seen.sum()) and model.get_h_table() is None
    v = model.get_v_table()
    assert v == o.v_dict()
    vp = os.path.join(tmp_path, "V.pkl")
    with open(vp, "wb") as f:
        pickle.dump({pickle.dumps(k): x for k, x in v.items()}, f)      # legacy bytes-key format (ffm_unified.py:91-107)

    actor = FloorFieldModelUnified(m, p, 8, learning_mode="actor_only", pretrained_v_path=vp, params=params)
    assert actor.initial_v_size == len(v)
    actor.set_epsilon(0.3)
    actor.reset(exit_pos=(0, 6), radius=5)
    steps, traj = actor.run(max_steps=40, return_trajectory=True)
    assert steps == len(traj) and traj.dtype == object
    init, cur, new = actor.get_v_table_size()
    assert init == len(v) and cur == init + new
    rows, total = actor.get_h_table_size()
    assert total == rows * 5 and len(actor.get_h_table()) == rows and all(len(x) == 5 for x in actor.get_h_table().values())


def test_trained_dropin_like_run_trained_ffm(cuda_device, tmp_path):
    """run_trained_ffm.py:199-243: positions / dff assigned per episode, run(max_steps)."""
    import json
    import pickle
    from helpers import GOLDEN
    from ffm_b200.model.ffm_trained_core import FloorFieldModel
    from oracle import unified_numpy

    z = np.load(os.path.join(GOLDEN, "trained_12x12.npz"))
    hz = np.load(os.path.join(GOLDEN, str(z["h_from"]) + ".npz"))
    params = {**json.loads(str(z["params"])), "seed": int(z["seed"])}
    helper = unified_numpy.UnifiedOracle(z["map"], z["sff"], np.zeros((0, 2)), "trained", params)
    htab = {helper.id_to_key(i): [float(x) for x in row] for i, row in zip(hz["h_ids"], hz["h_vals"])}
    hp, sp = os.path.join(tmp_path, "H.pkl"), os.path.join(tmp_path, "sff.npy")
    with open(hp, "wb") as f:
        pickle.dump({pickle.dumps(k): v for k, v in htab.items()}, f)
    np.save(sp, z["sff"])
    model = FloorFieldModel(z["map"], sp, len(z["pos0"]), hp, params)
    model._episode = 0                                 # replay the fixture's episode 0
    model.positions = z["pos0"].astype(np.int64)
    model.dff = np.zeros_like(model.map_array, dtype=np.float32)
    steps = model.run(save_prefix=None, save_interval=100, max_steps=int(z["max_steps"]))
    assert steps == int(z["steps"])
    assert np.array_equal(model.dff.view(np.uint32), z["final_dff"].view(np.uint32))


def test_mcq_dropin_like_main_learning(cuda_device, tmp_path):
    """main_learning.py:60-106: a NEW model per episode, alpha / gamma assigned after construction, the Q dict handed
    from episode to episode, finalize_timeouts() when the driver's own step limit hits first."""
    from ffm_b200.model.ffm_learning_core import FloorFieldModel
    from oracle import mcq_numpy

    m = assets.room_map(12, 12)
    sff = assets.sff_norm_min(m, "L1", np.float64)
    p = os.path.join(tmp_path, "sff.npy")
    np.save(p, sff)
    params = {"max_steps": 500, "step_penalty": 0.01, "stop_penalty": 0.3, "collision_penalty": 0.7, "seed": 99}
    o, shared_Q = None, {}
    for ep, (beta, limit) in enumerate([(1.0, 0), (0.5, 30), (0.2, 0)]):
        np.random.seed(100 + ep)
        model = FloorFieldModel(m, p, 15, {**params, "episode": 10 * ep})
        model.alpha, model.gamma = 0.2, 0.95
        model.Q = shared_Q
        model.reset()
        pos0 = np.array(model.positions, dtype=np.int64)
        if o is None:
            o = mcq_numpy.McqOracle(m, sff, pos0, params, None, alpha=0.2, gamma=0.95)
        o.reset(pos0)
        o.source = PhiloxSource(99, 10 * ep + 1)                 # reset() advanced the episode key once
        step = 0
        while model.positions.shape[0] > 0 and (limit == 0 or step < limit):
            model.step(beta)
            o.step(beta)
            step += 1
            if o.min_margin >= MARGIN_GUARD:
                assert np.array_equal(np.asarray(model.positions, dtype=np.int64).reshape(-1, 2), o.positions), (ep, step)
        if model.positions.shape[0] > 0:
            model.finalize_timeouts()
            o.finalize_timeouts()
        assert model.positions.shape[0] == 0
        shared_Q = model.Q
        want = o.q_dict()
        if o.min_margin >= MARGIN_GUARD:
            assert set(shared_Q) == set(want)
            assert all(np.array_equal(shared_Q[k].view(np.uint32), want[k].view(np.uint32)) for k in want), ep
    model.save_Q(os.path.join(tmp_path, "Q.pkl"))
