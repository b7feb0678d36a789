"""SFF oracles: the vectorised norm-min equals the literal Create_SFF.py loop; the C Dijkstra equals
the reference's fields on obstacle-free rooms and SciPy's Dijkstra on obstacle maps."""
import numpy as np
import pytest

from oracle import assets, c_oracle


@pytest.mark.parametrize("metric", ["L1", "L2", "Linf"])
def test_fast_norm_min_equals_literal_loop(metric):
    m = assets.room_map(23, 31)
    m[22, 7] = 3; m[10, 0] = 3       # several exits
    m[8:12, 9:14] = 2                # obstacle (ignored by the obstacle-blind generator)
    for dt in (np.float32, np.float64):
        a, b = assets.sff_norm_min(m, metric, dt), assets.sff_norm_min_fast(m, metric, dt)
        assert a.dtype == b.dtype and np.array_equal(a, b)


def test_geodesic_equals_reference_fields_on_open_room():
    m = assets.room_map(50, 50)
    assert np.array_equal(c_oracle.geodesic(m, "bfs4"), assets.sff_norm_min_fast(m, "L1", np.float32))
    assert np.array_equal(c_oracle.geodesic(m, "bfs8"), assets.sff_norm_min_fast(m, "Linf", np.float32))


def test_geodesic_matches_scipy_on_obstacle_map():
    from scipy.sparse import coo_matrix
    from scipy.sparse.csgraph import dijkstra
    m = assets.obstacle_map_c5(96, 80, index=3, n_exits=4)
    H, W = m.shape
    walk = (m == 0) | (m == 3)
    idx = np.arange(H * W).reshape(H, W)
    for mode, diag in (("bfs4", None), ("bfs8", 1.0), ("dijkstra8", float(np.float32(np.sqrt(2))))):
        rows, cols, wts = [], [], []
        for dr in (-1, 0, 1):
            for dc in (-1, 0, 1):
                if (dr, dc) == (0, 0) or (dr and dc and diag is None):
                    continue
                a = idx[max(0, -dr):H - max(0, dr), max(0, -dc):W - max(0, dc)]
                b = idx[max(0, dr):H - max(0, -dr), max(0, dc):W - max(0, -dc)]
                ok = walk.ravel()[a] & walk.ravel()[b]
                rows.append(a[ok]); cols.append(b[ok]); wts.append(np.full(ok.sum(), diag if (dr and dc) else 1.0))
        g = coo_matrix((np.concatenate(wts), (np.concatenate(rows), np.concatenate(cols))), shape=(H * W, H * W)).tocsr()
        d = dijkstra(g, indices=np.flatnonzero(m.ravel() == 3), min_only=True).reshape(H, W)
        d[~walk] = np.inf
        got = c_oracle.geodesic(m, mode)
        assert np.array_equal(np.isinf(got), np.isinf(d))
        f = np.isfinite(d)
        if mode == "dijkstra8":
            np.testing.assert_allclose(got[f], d[f], rtol=2e-6)
        else:
            assert np.array_equal(got[f], d[f].astype(np.float32))
