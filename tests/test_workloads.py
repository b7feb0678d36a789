"""The product-side workload generators (ffm_b200/workloads.py) and the checker's independent copies
(oracle/assets.py) describe the same inputs; placement keys match the oracle's Philox."""
import numpy as np

from ffm_b200 import workloads
from oracle import assets, philox


def test_maps_agree():
    assert np.array_equal(workloads.room_map(64, 64), assets.room_map(64, 64))
    assert np.array_equal(workloads.rooms_map_c3(), assets.rooms_map_c3())
    assert np.array_equal(workloads.obstacle_map_c5(128, 160, index=2), assets.obstacle_map_c5(128, 160, index=2))
    m = workloads.room_map(40, 30)
    assert np.array_equal(workloads.sff_room(m, "moore"), assets.sff_norm_min_fast(m, "Linf", np.float32))
    assert np.array_equal(workloads.sff_room(m, "neumann"), assets.sff_norm_min_fast(m, "L1", np.float32))


def test_c3_plan_properties():
    m = workloads.rooms_map_c3()
    assert m.shape == (256, 256) and (m == 3).sum() == 4
    border = np.concatenate([m[0], m[-1], m[:, 0], m[:, -1]])
    assert (border != 0).all() and (m == 0).sum() > 60000


def test_placement_is_keyed_by_global_episode():
    m = workloads.room_map(20, 20)
    a = workloads.place(m, 50, 6, 100, 7)
    b = workloads.place(m, 50, 3, 103, 7)
    assert np.array_equal(a[3:], b)                       # same global ids -> same placement
    free = np.argwhere(m == 0)
    keys, _ = philox.draw2(7, 104, 0, philox.STREAM_PLACE, np.arange(len(free)))
    assert np.array_equal(a[4], free[np.argsort(keys, kind="stable")[:50]])
    for e in range(6):
        assert len({(int(r), int(c)) for r, c in a[e]}) == 50 and (m[a[e][:, 0], a[e][:, 1]] == 0).all()
