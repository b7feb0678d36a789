"""Edge cases of the rollout kernels against the C / NumPy oracles: empty and single-pedestrian episodes, ragged
batches, the maximum pedestrian capacity, maps above 65 536 cells (32-bit cell ids), obstacle maps with several
exits, non-square maps, walls coded 1 (create_test_map of run_unified_critic_training.py:58-66)."""
import numpy as np
import pytest

from helpers import MARGIN_GUARD, pack_positions, random_positions
from oracle import assets, c_oracle

pytestmark = pytest.mark.gpu


def _recorded(m, sff, pos, n, params, seed, T, base=0, threads=4, track=True):
    return c_oracle.run_core_batch(m, sff, pos, n, params, seed=seed, episode_base=base, max_steps=T, threads=threads,
                                   track_dff=track, traj_steps=T, want_state=True, guard=1e-5, record_moves=T)


def _check_recorded(sim, ref, T, W):
    import torch
    cells, cnt = sim.rollout(T, draws=dict(move=torch.from_numpy(ref["move_draws"]).cuda()), record=T)
    torch.cuda.synchronize()
    steps, ped = sim.counters()
    assert np.array_equal(steps, np.minimum(ref["steps"], T)) and np.array_equal(ped, ref["ped_steps"])
    cells, cnt = cells.cpu().numpy(), cnt.cpu().numpy()
    for e in range(len(steps)):
        s = int(steps[e])
        assert np.array_equal(cnt[e, :s], ref["traj_n"][e, :s]), e
        mask = np.arange(cells.shape[2])[None, :] < cnt[e, :s, None]
        assert np.array_equal(cells[e, :s][mask], ref["traj"][e, :s][mask]), e


def test_empty_single_and_ragged_episodes(cuda_device, core_kernel):
    from ffm_b200 import BatchSim
    m = assets.room_map(16, 20)
    sff = assets.sff_norm_min_fast(m, "L2", np.float32)
    rng = np.random.RandomState(0)
    counts = [0, 1, 2, 33, 64, 65, 100]
    pos0 = [random_positions(m, k, rng) if k else np.zeros((0, 2), np.int64) for k in counts]
    pos, n = pack_positions(pos0, 100)
    ref = _recorded(m, sff, pos, n, {"neighborhood": "moore"}, 3, 400)
    sim = BatchSim(m, sff, len(counts), 100, {"neighborhood": "moore"}, seed=3)
    sim.set_positions(pos, n)
    _check_recorded(sim, ref, 400, 20)
    assert sim.counters()[0][0] == 0 and (sim.get_positions()[1] == 0).all()
    assert np.array_equal(sim.get_dff().view(np.uint32), ref["final_dff"].view(np.uint32))


def test_32bit_cell_ids_above_65536_cells(cuda_device, core_kernel):
    """300 x 260 = 78 000 cells: PosT = uint32 variant, fields in global memory."""
    from ffm_b200 import BatchSim
    m = assets.obstacle_map_c5(300, 260, index=1, n_exits=4)
    sff = c_oracle.geodesic(m, "bfs8")
    rng = np.random.RandomState(1)
    reach = np.argwhere((m == 0) & np.isfinite(sff))
    pos0 = [reach[rng.choice(len(reach), 700, replace=False)] for _ in range(2)]
    pos, n = pack_positions(pos0, 700)
    params = {"neighborhood": "moore", "k_S": 2.0, "k_D": 1.0}
    ref = _recorded(m, sff, pos, n, params, 5, 150)
    sim = BatchSim(m, sff, 2, 700, params, seed=5)
    sim.set_positions(pos, n)
    _check_recorded(sim, ref, 150, 260)
    assert np.array_equal(sim.get_dff().view(np.uint32), ref["final_dff"].view(np.uint32))


def test_maximum_capacity(cuda_device, core_kernel):
    """n_max = 16 380 pedestrians (the owner grid's 14-bit id space) in a 130 x 130 room: 96.9 % of the cells
    occupied at t = 0.  (256 x 256 maps hold at most ~10 100 pedestrians: one SM's shared memory.)"""
    from ffm_b200 import BatchSim
    from ffm_b200.workloads import place
    m = assets.room_map(130, 130)
    sff = assets.sff_norm_min_fast(m, "Linf", np.float32)
    N = 16380
    pos = place(m, N, 1, 0, 9)
    n = np.array([N], np.int32)
    params = {"neighborhood": "moore"}
    ref = _recorded(m, sff, pos, n, params, 9, 60, threads=1)
    sim = BatchSim(m, sff, 1, N, params, seed=9)
    sim.set_positions(pos, n)
    _check_recorded(sim, ref, 60, 130)
    assert np.array_equal(sim.get_dff().view(np.uint32), ref["final_dff"].view(np.uint32))
    with pytest.raises(Exception):
        BatchSim(m, sff, 1, N + 1, params)


def test_maps_beyond_one_sm_run_as_clusters(cuda_device):
    """520 x 400 = 208 000 cells with 12 000 pedestrians, DFF on: neither kernel's per-episode state fits one SM's
    227 KB, so ffm_create picks the cell-centric kernel as a thread-block cluster (owner grid + claim masks in
    distributed shared memory, DFF and score left in L2 at this size).  Recorded-draw parity, 120 steps."""
    from ffm_b200 import BatchSim
    m = assets.obstacle_map_c5(520, 400, index=2, n_exits=8)
    sff = c_oracle.geodesic(m, "dijkstra8")
    rng = np.random.RandomState(4)
    reach = np.argwhere((m == 0) & np.isfinite(sff))
    N = 12000
    pos0 = [reach[rng.choice(len(reach), N - 1000 * e, replace=False)] for e in range(2)]
    pos, n = pack_positions(pos0, N)
    params = {"neighborhood": "moore", "k_S": 1.5, "k_D": 1.0}
    ref = _recorded(m, sff, pos, n, params, 11, 120)
    sim = BatchSim(m, sff, 2, N, params, seed=11)
    info = sim.kernel_info()
    assert info["name"] == "ffm_cell_rollout_kernel" and info["cluster"] >= 2, info
    sim.set_positions(pos, n)
    _check_recorded(sim, ref, 120, 400)
    assert np.array_equal(sim.get_dff().view(np.uint32), ref["final_dff"].view(np.uint32))


@pytest.mark.parametrize("variant", ["cell", "cl2", "cl4"])
def test_compact_trajectory_record(cuda_device, monkeypatch, variant):
    """The compact record (int16 (row, col) pairs, CSR offsets per step, 4 bytes per pedestrian-step) holds exactly
    what run() collects (ffm_core.py:125): compared with the dense record of the same run and with the oracle;
    ragged batch incl. an empty episode; an undersized buffer reports the overflow instead of writing past it."""
    import torch
    from ffm_b200 import BatchSim
    from ffm_b200.sim import unpack_trajectory
    monkeypatch.setenv("FFM_KERNEL", "cell")
    if variant.startswith("cl"):
        monkeypatch.setenv("FFM_CLUSTER", variant[2:])
    m = assets.room_map(40, 72)
    sff = assets.sff_norm_min_fast(m, "Linf", np.float32)
    rng = np.random.RandomState(7)
    counts = [0, 1, 37, 300, 513]
    pos0 = [random_positions(m, k, rng) if k else np.zeros((0, 2), np.int64) for k in counts]
    pos, n = pack_positions(pos0, 513)
    params = {"neighborhood": "moore"}
    T = 500
    ref = _recorded(m, sff, pos, n, params, 21, T)
    sim = BatchSim(m, sff, len(counts), 513, params, seed=21)
    sim.set_positions(pos, n)
    cap = int(ref["ped_steps"].max()) + 4 * T
    rec = sim.rollout(T, draws=dict(move=torch.from_numpy(ref["move_draws"]).cuda()), record=T, compact_cap=cap)
    torch.cuda.synchronize()
    steps, ped = sim.counters()
    assert np.array_equal(ped, ref["ped_steps"])
    ct, off, cnt = rec["ctraj"].cpu().numpy(), rec["off"].cpu().numpy(), rec["n"].cpu().numpy()
    for e in range(len(counts)):
        s = int(steps[e])
        assert np.array_equal(cnt[e, :s], ref["traj_n"][e, :s])
        assert off[e, 0] == 0 and np.array_equal(np.diff(off[e, :s + 1]), (cnt[e, :s] + 3) // 4 * 4)
        rows = unpack_trajectory(ct[e], off[e], cnt[e], steps=s)
        for t in range(s):
            assert rows[t].dtype == np.int64 and rows[t].shape == (cnt[e, t], 2)
            assert np.array_equal(rows[t][:, 0] * 72 + rows[t][:, 1], ref["traj"][e, t, :cnt[e, t]]), (e, t)
        # padding entries are (-1, -1)
        if s:
            pad = ct[e, off[e, 0] + cnt[e, 0]:off[e, 1]]
            assert (pad == -1).all()
    # bytes: 4 per (padded) pedestrian-step
    assert int(off[np.arange(len(counts)), steps].sum()) * 4 <= 4 * int(ped.sum()) + 16 * int(steps.sum())
    # overflow: a buffer of half the size stops recording and says so
    sim.set_positions(pos, n)
    rec = sim.rollout(T, draws=dict(move=torch.from_numpy(ref["move_draws"]).cuda()), record=T, compact_cap=cap // 2)
    torch.cuda.synchronize()
    off2 = rec["off"].cpu().numpy()
    assert (off2[4] == -1).any() and off2[4, 1] > 0 and np.array_equal(sim.counters()[1], ref["ped_steps"])
    with pytest.raises(Exception):
        unpack_trajectory(rec["ctraj"][4].cpu().numpy(), off2[4], rec["n"][4].cpu().numpy(), steps=int(steps[4]))


@pytest.mark.parametrize("nbh", ["neumann", "moore"])
def test_obstacles_and_several_exits_full_episodes(cuda_device, core_kernel, nbh):
    from ffm_b200 import BatchSim
    m = assets.obstacle_map_c5(96, 80, index=3, n_exits=8)
    sff = c_oracle.geodesic(m, "bfs4" if nbh == "neumann" else "dijkstra8").astype(np.float64)
    rng = np.random.RandomState(2)
    reach = np.argwhere((m == 0) & np.isfinite(sff))
    pos0 = [reach[rng.choice(len(reach), 600 - 50 * e, replace=False)] for e in range(3)]
    pos, n = pack_positions(pos0, 600)
    params = {"neighborhood": nbh, "k_S": 1.5}
    T = 1200
    ref = _recorded(m, sff, pos, n, params, 21, T, base=5)
    assert (ref["steps"] < T).all()
    sim = BatchSim(m, sff, 3, 600, params, seed=21, episode_base=5)
    sim.set_positions(pos, n)
    _check_recorded(sim, ref, T, 80)
    assert np.array_equal(sim.get_dff().view(np.uint32), ref["final_dff"].view(np.uint32))


def test_unified_map_with_walls_coded_1(cuda_device):
    """create_test_map (run_unified_critic_training.py:58-66) codes walls as 1: blocked for moves AND seen as
    'pedestrian' by the diagonal test of _encode_state."""
    import torch
    from ffm_b200 import UnifiedSim
    from oracle import unified_numpy
    from oracle.inject import PhiloxSource
    m = np.zeros((14, 14), np.uint8)
    m[0, :] = 1; m[-1, :] = 1; m[:, 0] = 1; m[:, -1] = 1
    m[7, 13] = 3
    m[5:8, 6] = 1                                          # an interior wall coded 1
    sff = np.ones(m.shape, np.float32) * 999                # create_test_sff: 999 on walls, L1 to the first exit
    for i in range(14):
        for j in range(14):
            if m[i, j] in (0, 3):
                sff[i, j] = abs(i - 7) + abs(j - 13)
    P = dict(k_S=10, k_D=1, alpha_v=0.05, gamma=0.95, exit_reward=100.0, step_penalty=-1.0, collision_penalty=-1.0,
             neighborhood="neumann", block_size=5)
    rng = np.random.RandomState(3)
    pos0 = random_positions(m, 30, rng)
    o = unified_numpy.UnifiedOracle(m, sff, pos0, "critic_only", P, PhiloxSource(8, 0))
    r = o.run(max_steps=300)
    assert r["min_margin"] >= MARGIN_GUARD
    sim = UnifiedSim(m, sff, 1, 30, mode="critic_only", learn="exact", params=P, seed=8)
    sim.set_positions(*pack_positions([pos0], 30))
    cells, cnt = sim.rollout(300, record=300)
    torch.cuda.synchronize()
    assert sim.counters()[0][0] == r["steps"]
    cells, cnt = cells.cpu().numpy()[0], cnt.cpu().numpy()[0]
    for t, want in enumerate(r["traj"]):
        assert np.array_equal(cells[t, :cnt[t]], want[:, 0] * 14 + want[:, 1]), t
    V, vs, _, _ = sim.get_tables()
    assert np.array_equal(vs, o.v_seen) and np.array_equal(V.view(np.uint64), o.V.view(np.uint64))


def test_handles_of_one_kernel_variant_with_different_shared_memory(cuda_device):
    """The opt-in dynamic shared-memory size is per-function state: a later, smaller handle must not break an
    earlier, larger one that runs the same kernel instantiation."""
    from ffm_b200 import BatchSim
    m = assets.room_map(40, 40)
    sff = assets.sff_norm_min_fast(m, "Linf", np.float32)
    rng = np.random.RandomState(0)
    big = BatchSim(m, sff, 2, 1400, {"k_D": 0}, seed=1, track_dff=False)
    small = BatchSim(m, sff, 2, 300, {"k_D": 0}, seed=1, track_dff=False)
    assert big.kernel_info()["smem_bytes"] > small.kernel_info()["smem_bytes"]
    assert big.kernel_info()["threads"] == small.kernel_info()["threads"]
    for sim, N in ((big, 1400), (small, 300), (big, 1400)):
        sim.set_positions(*pack_positions([random_positions(m, N, rng) for _ in range(2)], N))
        sim.rollout(3000)
        assert (sim.get_positions()[1] == 0).all()


def test_two_pedestrians_on_one_cell_are_rejected(cuda_device, core_kernel):
    """ffm_set_positions validates map codes and bounds per pedestrian; a DUPLICATE cell is caught by the rollout kernel's
    placement (the second arrival finds the cell taken) and surfaces as ValueError at the next host read."""
    from ffm_b200 import BatchSim
    m = assets.room_map(16, 20)
    sff = assets.sff_norm_min_fast(m, "L2", np.float32)
    pos = np.full((2, 8, 2), -1, np.int32)
    pos[0, :3] = [(3, 3), (4, 4), (5, 5)]
    pos[1, :3] = [(3, 3), (7, 7), (3, 3)]                   # episode 1 holds a duplicate
    sim = BatchSim(m, sff, 2, 8, {"neighborhood": "moore"}, seed=1)
    sim.set_positions(pos, np.array([3, 3], np.int32))
    sim.rollout(1)
    with pytest.raises(ValueError, match="same cell"):
        sim.counters()


def test_duplicates_rejected_in_the_table_models(cuda_device):
    from ffm_b200 import McqSim, UnifiedSim
    m = assets.room_map(12, 12)
    sff = assets.sff_norm_min(m, "L1", np.float32)
    pos = np.full((1, 4, 2), -1, np.int32)
    pos[0, :3] = [(3, 3), (5, 5), (3, 3)]
    for sim in (UnifiedSim(m, sff, 1, 4, mode="critic_only", learn="none", params={"block_size": 1}, seed=1),
                McqSim(m, sff, 1, 4, learn="none", params={"max_steps": 10}, seed=1)):
        sim.set_positions(pos, np.array([3], np.int32))
        sim.rollout(1)
        with pytest.raises(ValueError, match="same cell"):
            sim.counters()
