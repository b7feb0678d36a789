"""Device-side initial placement (ffm_place) against the host placement with the same Philox keys and against
the semantics of initialize_agents (ffm_core.py:23-26, ffm_unified.py:131-171)."""
import numpy as np
import pytest

from oracle import assets, philox

pytestmark = pytest.mark.gpu


def _expected(m, n, episode, seed, exit_pos=None, radius=None):
    free = np.argwhere(m == 0)
    if radius is not None:
        free = free[np.abs(free[:, 0] - exit_pos[0]) + np.abs(free[:, 1] - exit_pos[1]) <= radius]
    keys, _ = philox.draw2(seed, episode, 0, philox.STREAM_PLACE, np.arange(len(free)))
    return free[np.argsort(keys, kind="stable")[:min(n, len(free))]]


def test_place_matches_keyed_host_placement(cuda_device):
    from ffm_b200 import BatchSim
    m = assets.room_map(64, 64)
    sff = assets.sff_norm_min_fast(m, "Linf", np.float32)
    B, N, seed, base = 5, 1024, 0x5EED0002, 4090
    sim = BatchSim(m, sff, B, N, {"k_D": 0}, seed=seed, episode_base=base, track_dff=False)
    sim.place(N)
    pos, n = sim.get_positions()
    assert (n == N).all()
    for e in range(B):
        assert np.array_equal(pos[e], _expected(m, N, base + e, seed)), e
    # ragged counts, and the bench's host generator gives the same thing
    from ffm_b200.workloads import place
    assert np.array_equal(pos, place(m, N, B, base, seed))
    sim.place(np.array([1, 10, 0, 1024, 500], np.int32))
    pos, n = sim.get_positions()
    assert n.tolist() == [1, 10, 0, 1024, 500] and np.array_equal(pos[3], _expected(m, 1024, base + 3, seed))
    with pytest.raises(ValueError):
        sim.place(N + 5000)                              # more pedestrians than free cells (ffm_core.py:25)


def test_place_radius_limited_and_clamped(cuda_device):
    from ffm_b200 import UnifiedSim
    m = assets.room_map(12, 12)
    sff = assets.sff_norm_min(m, "L1", np.float32)
    sim = UnifiedSim(m, sff, 3, 100, mode="critic_only", learn="none", params={"block_size": 1}, seed=5, episode_base=9)
    sim.place(50, exit_pos=(0, 6), radius=3)
    pos, n = sim.get_positions()
    want = _expected(m, 50, 9, 5, (0, 6), 3)
    assert n[0] == len(want) < 50                      # clamped to the cells inside the radius (ffm_unified.py:160-162)
    assert np.array_equal(pos[0, :n[0]], want)
    sim.place(4, exit_pos=(0, 6), radius=15)
    pos, n = sim.get_positions()
    assert (n == 4).all() and np.array_equal(pos[2, :4], _expected(m, 4, 11, 5, (0, 6), 15))
    sim.rollout(300)
    assert (sim.get_positions()[1] == 0).all()


def test_place_on_large_maps(cuda_device):
    """C3 floor plan (60 000+ free cells): the histogram select keeps the sort in one SM's shared memory."""
    from ffm_b200 import BatchSim
    from ffm_b200.workloads import place, rooms_map_c3
    m = rooms_map_c3()
    sim = BatchSim(m, np.zeros(m.shape, np.float32), 3, 10000, {"k_D": 0}, seed=0x5EED0003, episode_base=7, track_dff=False)
    sim.place(10000)
    pos, n = sim.get_positions()
    assert (n == 10000).all() and np.array_equal(pos, place(m, 10000, 3, 7, 0x5EED0003))
    sim.place(np.array([10000, 17, 0], np.int32))
    pos, n = sim.get_positions()
    assert n.tolist() == [10000, 17, 0] and np.array_equal(pos[1, :17], _expected(m, 17, 8, 0x5EED0003))
