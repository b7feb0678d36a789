"""Pins the C restatement (oracle/c/ffm_oracle.c: CPU baseline + full-size checker) to the NumPy
oracle, which is itself pinned to the unmodified reference (tests/test_oracle_vs_golden.py)."""
import numpy as np
import pytest

from helpers import CORE_FIXTURES, MARGIN_GUARD, load_golden, oracle_core_episode, pack_positions, random_positions
from oracle import assets, c_oracle


@pytest.mark.parametrize("name", CORE_FIXTURES)
def test_c_oracle_reproduces_reference_fixture(name):
    g = load_golden(name)
    N, W = len(g["pos0"]), g["map"].shape[1]
    pos, n = pack_positions([g["pos0"].astype(np.int32)], N)
    T = int(g["steps"])
    r = c_oracle.run_core_batch(g["map"], g["sff"], pos, n, g["params"], seed=int(g["seed"]), episode_base=int(g["episode"]),
                                traj_steps=T, want_state=True)
    assert r["steps"][0] == T
    for t, want in enumerate(g["traj_list"]):
        assert r["traj_n"][0, t] == len(want)
        assert np.array_equal(r["traj"][0, t, :len(want)], want[:, 0] * W + want[:, 1]), f"step {t}"
    # final DFF: the fixture's last snapshot is only the final one when (T-1) % every == 0
    if (T - 1) % int(g["dff_every"]) == 0:
        assert np.array_equal(r["final_dff"][0].view(np.uint32), g["dff"][-1].view(np.uint32))
    assert abs(r["min_margin"][0] - float(g["min_margin"])) < 1e-6


def test_c_oracle_matches_numpy_oracle_batch():
    m = assets.room_map(18, 22)
    sff = assets.sff_norm_min(m, "Linf", np.float32)
    params = {"k_S": 2, "k_D": 1.5, "neighborhood": "moore"}
    rng = np.random.RandomState(3)
    pos0 = [random_positions(m, 90 - 7 * e, rng) for e in range(4)]
    pos, n = pack_positions(pos0, 90)
    r = c_oracle.run_core_batch(m, sff, pos, n, params, seed=99, episode_base=10, threads=3, want_state=True)
    for e in range(4):
        o = oracle_core_episode(m, sff, pos0[e], params, 99, 10 + e)
        if o["min_margin"] < MARGIN_GUARD:
            continue
        assert r["steps"][e] == o["steps"]
        assert r["ped_steps"][e] == len(pos0[e]) + sum(len(p) for p in o["traj"][:-1])
        assert np.array_equal(r["final_dff"][e].view(np.uint32), o["final_dff"].view(np.uint32))
