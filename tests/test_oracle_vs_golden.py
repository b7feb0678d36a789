"""Pins the oracle (oracle/ffm_numpy.py) to fixtures produced by the unmodified reference
(oracle/make_golden.py ran model/ffm_core.py under the injected-draw protocol)."""
import hashlib
import json
import os

import numpy as np
import pytest

from helpers import CORE_FIXTURES, GOLDEN, load_golden
from oracle import assets, ffm_numpy
from oracle.inject import BufferSource, PhiloxSource


@pytest.mark.parametrize("name", CORE_FIXTURES)
def test_core_oracle_reproduces_reference(name):
    g = load_golden(name)
    o = ffm_numpy.CoreOracle(g["map"], g["sff"], g["pos0"], g["params"],
                             PhiloxSource(int(g["seed"]), int(g["episode"])))
    o.probs_log = []
    r = o.run()
    assert r["steps"] == int(g["steps"])
    for t, want in enumerate(g["traj_list"]):
        assert np.array_equal(r["traj"][t], want), f"step {t}"
    every = int(g["dff_every"])
    for k, want in enumerate(g["dff"]):
        assert np.array_equal(r["dff"][k * every].view(np.uint32), want.view(np.uint32)), f"dff at step {k * every}"
    assert r["min_margin"] == float(g["min_margin"])
    # move probabilities: same machine + same NumPy kernels => identical; bar is 1e-6 relative
    for (t, idx, p), want, meta in zip(o.probs_log, g["probs"], g["probs_meta"]):
        assert (t, idx, len(p)) == tuple(meta)
        np.testing.assert_allclose(p, want[:len(p)], rtol=1e-6, atol=0)


def test_stock_seeded_main_run_replays():
    """main.py with config/default_config.yaml, seed 42 -> 272 steps / 14 079 ped-steps."""
    z = np.load(os.path.join(GOLDEN, "stock_main_seed42.npz"))
    params = json.loads(str(z["params"]))
    m = assets.room_map(50, 50)
    sff = assets.sff_norm_min(m, "L1", np.float64)
    conflict = {(int(t), int(c)): tuple(u) for t, c, u in zip(z["conflict_t"], z["conflict_cell"], z["conflict_u"])}
    o = ffm_numpy.CoreOracle(m, sff, z["pos0"], params, BufferSource(z["move"], conflict))
    r = o.run()
    assert r["steps"] == int(z["steps"]) == 272
    counts = z["traj_counts"]
    assert 100 + int(counts[:-1].sum()) == 14079
    offs = np.concatenate([[0], np.cumsum(counts)])
    for t in range(272):
        assert np.array_equal(r["traj"][t], z["traj"][offs[t]:offs[t + 1]].astype(np.int64))
    assert np.array_equal(o.dff.view(np.uint32), z["final_dff"].view(np.uint32))


def test_asset_generators_reproduce_shipped_files():
    """Create_Map.py / Create_SFF.py restatements regenerate data/maps + data/sff bit for bit."""
    with open(os.path.join(GOLDEN, "shipped_assets.json")) as f:
        want = json.load(f)
    m = assets.room_map(50, 50, dtype=np.int64)
    got = {"data/maps/simple_room.npy": m}
    for metric in ("L1", "L2", "Linf"):
        got[f"data/sff/distance_{metric}.npy"] = assets.sff_norm_min(m, metric, np.float64)
    for rel, meta in want.items():
        a = got[rel]
        assert list(a.shape) == meta["shape"]
        if meta["dtype"] != str(a.dtype):
            a = a.astype(meta["dtype"])
        assert hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest() == meta["sha256"], rel


def test_12x12_generator_summary():
    """create_12x12_map_and_sff.py prints range [0, 15] and 101 valid cells."""
    m = assets.room_map(12, 12)
    s = assets.sff_norm_min(m, "L1", np.float32)
    assert s.dtype == np.float32 and int(np.isfinite(s).sum()) == 101
    assert s[np.isfinite(s)].min() == 0 and s[np.isfinite(s)].max() == 15
    assert (m == 0).sum() == 100 and tuple(np.argwhere(m == 3)[0]) == (0, 6)


def test_update_dff_constants_are_float32():
    """c0 = f32(0.64), c1 = f32(0.02) Moore / f32(0.04) von Neumann and separate mul/add."""
    rng = np.random.RandomState(0)
    d = (rng.rand(9, 11) * 3).astype(np.float32)
    for nb, c1 in ((ffm_numpy.MOORE, 0.02), (ffm_numpy.NEUMANN, 0.04)):
        got = ffm_numpy.update_dff(d.copy(), ffm_numpy.CORE_DEFAULTS, nb)
        c0f, c1f = np.float32(0.2 * 0.8 / len(nb)), None
        s = np.float32((1 - 0.2) * (1 - 0.2)) * d
        acc = s.copy()
        pad = np.pad(s, 1)
        for dx, dy in nb:
            acc = acc + np.float32(0.2 * (1 - 0.2) / len(nb)) * pad[1 + dx:10 + dx, 1 + dy:12 + dy]
        acc[acc < np.float32(1e-4)] = 0
        assert got.dtype == np.float32 and np.array_equal(got.view(np.uint32), acc.view(np.uint32))
        assert abs(float(c0f) - c1) < 1e-6
