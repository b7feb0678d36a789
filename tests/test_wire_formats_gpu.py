"""Wire formats: what the reference's drivers write and its inspection / visualisation
scripts read back must come out of the GPU path unchanged.

  * main.py:44-52               positions_log -> np.save(..., np.array(log, dtype=object))  ("positions.npy")
  * ffm_core.py:119-133         run(save_prefix=...) -> {prefix}_{step}.npz / {prefix}_final.npz, key "positions", int32
  * run_actor_only_training.py:205-218   np.savez_compressed(file, positions=trajectory (object array), episode=..., N=..., ...)
  * ffm_learning_core.py:364-367         save_Q -> pickle of {(bytes, (bx, by)): float32[5]}
"""
import os
import pickle

import numpy as np
import pytest

from helpers import MARGIN_GUARD
from oracle import assets, ffm_numpy
from oracle.inject import PhiloxSource

pytestmark = pytest.mark.gpu


def _assets(tmp_path, h=12, w=12):
    m = assets.room_map(h, w)
    sff = assets.sff_norm_min(m, "L1", np.float32)
    p = os.path.join(tmp_path, "sff.npy")
    np.save(p, sff)
    return m, sff, p


def test_main_py_positions_npy_object_array(cuda_device, tmp_path):
    """The loop of main.py:42-52 over the drop-in: file written with np.save(dtype=object), read back the way
    inspect / visualisation scripts do (allow_pickle), equal to the oracle's per-step position arrays."""
    from ffm_b200.model.ffm_core import FloorFieldModel
    m, sff, p = _assets(tmp_path)
    np.random.seed(11)
    model = FloorFieldModel(m, p, 40, {"seed": 77, "neighborhood": "neumann"})
    o = ffm_numpy.CoreOracle(m, sff, model.positions.copy(), model.params, PhiloxSource(77, 0))
    positions_log, step = [], 0
    while model.positions.shape[0] > 0:
        model.step()
        positions_log.append(np.copy(model.positions))
        step += 1
    out = os.path.join(tmp_path, "positions.npy")
    np.save(out, np.array(positions_log, dtype=object))
    back = np.load(out, allow_pickle=True)
    r = o.run(keep_dff=False)
    assert r["min_margin"] >= MARGIN_GUARD
    assert back.dtype == object and len(back) == r["steps"] == step
    for t in range(step):
        assert back[t].dtype == np.int64 and back[t].shape == r["traj"][t].shape and np.array_equal(back[t], r["traj"][t])


def test_compact_record_gives_the_same_positions_npy(cuda_device, tmp_path, monkeypatch):
    """The batched path's compact record, unpacked, is byte-for-byte the positions.npy of the stepping loop."""
    import io
    import torch
    from ffm_b200 import BatchSim
    from ffm_b200.sim import unpack_trajectory
    monkeypatch.setenv("FFM_KERNEL", "cell")
    m, sff, _ = _assets(tmp_path, 20, 28)
    rng = np.random.RandomState(5)
    free = np.argwhere(m == 0)
    pos0 = free[rng.choice(len(free), 60, replace=False)]
    params = {"neighborhood": "moore"}
    o = ffm_numpy.CoreOracle(m, sff, pos0, params, PhiloxSource(9, 0))
    r = o.run(keep_dff=False)
    assert r["min_margin"] >= MARGIN_GUARD
    sim = BatchSim(m, sff, 1, 60, params, seed=9)
    sim.set_positions(pos0[None].astype(np.int32), np.array([60], np.int32))
    rec = sim.rollout(1000, record=1000, compact_cap=60 * 1000)
    torch.cuda.synchronize()
    steps = int(sim.counters()[0][0])
    rows = unpack_trajectory(rec["ctraj"][0].cpu().numpy(), rec["off"][0].cpu().numpy(), rec["n"][0].cpu().numpy(), steps=steps)
    a, b = io.BytesIO(), io.BytesIO()
    np.save(a, np.array(rows, dtype=object))
    np.save(b, np.array(r["traj"], dtype=object))
    assert steps == r["steps"] and a.getvalue() == b.getvalue()


def test_run_save_prefix_npz_files(cuda_device, tmp_path):
    """run(save_prefix, save_interval) (ffm_core.py:119-133): while nobody has left, every interval's rows are equally long
    and {prefix}_{step}.npz holds int32 [interval, n, 2]; the first ragged interval raises ValueError exactly where
    np.array(buffer, dtype=np.int32) raises in the reference on current NumPy."""
    from ffm_b200.model.ffm_core import FloorFieldModel
    m = assets.room_map(30, 12)
    sff = assets.sff_norm_min(m, "L1", np.float32)
    p = os.path.join(tmp_path, "sff.npy")
    np.save(p, sff)
    far = np.argwhere(m == 0)
    far = far[far[:, 0] >= 22][:20]                          # 20 pedestrians at least 21 rows from the exit in row 0
    np.random.seed(0)
    model = FloorFieldModel(m, p, 20, {"seed": 3, "neighborhood": "neumann"})
    model.positions = far
    o = ffm_numpy.CoreOracle(m, sff, far, model.params, PhiloxSource(3, 0))
    r = o.run(keep_dff=False)
    assert r["min_margin"] >= MARGIN_GUARD
    first_exit = next(t for t, q in enumerate(r["traj"]) if len(q) < 20)
    assert first_exit >= 20
    prefix = os.path.join(tmp_path, "run")
    with pytest.raises(ValueError):
        model.run(save_prefix=prefix, save_interval=5)
    full = first_exit // 5                                   # complete intervals before anybody left
    for k in range(1, full + 1):
        z = np.load(f"{prefix}_{5 * k}.npz")
        assert z["positions"].dtype == np.int32 and z["positions"].shape == (5, 20, 2)
        assert np.array_equal(z["positions"], np.array(r["traj"][5 * (k - 1):5 * k], dtype=np.int32))
    assert not os.path.exists(f"{prefix}_{5 * (full + 1)}.npz")


def test_unified_trajectory_npz_like_the_training_drivers(cuda_device, tmp_path):
    """run(return_trajectory=True) -> np.savez_compressed(positions=trajectory, episode=, N=, total_episode=, steps=)
    (run_actor_only_training.py:205-218), read back with allow_pickle."""
    from ffm_b200.model.ffm_unified import FloorFieldModelUnified
    from oracle import unified_numpy
    m, sff, p = _assets(tmp_path)
    params = dict(k_S=10, k_D=1, alpha_v=0.01, gamma=0.99, exit_reward=100.0, step_penalty=-1.0, collision_penalty=-1.0,
                  neighborhood="neumann", block_size=1, seed=515)
    np.random.seed(2)
    model = FloorFieldModelUnified(m, p, 12, learning_mode="critic_only", params=params)
    model.reset(exit_pos=(0, 6), radius=9)
    o = unified_numpy.UnifiedOracle(m, sff, model.positions.copy(), "critic_only", params, PhiloxSource(515, 1))
    steps, trajectory = model.run(max_steps=200, return_trajectory=True)
    r = o.run(max_steps=200)
    f = os.path.join(tmp_path, "trajectory_N12_ep00000_total00000.npz")
    np.savez_compressed(f, positions=trajectory, episode=0, N=12, total_episode=0, steps=steps)
    z = np.load(f, allow_pickle=True)
    assert int(z["steps"]) == steps == r["steps"] and int(z["N"]) == 12
    assert z["positions"].dtype == object and len(z["positions"]) == steps
    if r["min_margin"] >= MARGIN_GUARD:
        for t in range(steps):
            assert np.array_equal(np.asarray(z["positions"][t], dtype=np.int64), r["traj"][t])


def test_mcq_save_q_pickle(cuda_device, tmp_path):
    """save_Q (ffm_learning_core.py:364-367): a pickled dict {(bytes of the 3x3 window, (bx, by)): float32[5]}."""
    from ffm_b200.model.ffm_learning_core import FloorFieldModel
    m, sff, p = _assets(tmp_path)
    np.random.seed(4)
    model = FloorFieldModel(m, p, 15, {"seed": 8, "max_steps": 60})
    model.alpha, model.gamma = 0.1, 0.99
    model.reset()
    while model.positions.shape[0] > 0:
        model.step(beta=1.0)
    f = os.path.join(tmp_path, "Q.pkl")
    model.save_Q(f)
    with open(f, "rb") as fh:
        Q = pickle.load(fh)
    assert len(Q) > 0
    for (cells, blk), row in Q.items():
        assert isinstance(cells, bytes) and len(cells) == 9 and len(blk) == 2
        assert isinstance(row, np.ndarray) and row.dtype == np.float32 and row.shape == (5,)
    assert any(row.any() for row in Q.values())
