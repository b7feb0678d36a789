"""CUDA unified / trained models (C ABI -> ffm_unified_rollout_kernel) against fixtures from the
unmodified reference and against the oracle: trajectories, V / H tables (bit-exact float64), DFF."""
import json
import os

import numpy as np
import pytest

from helpers import GOLDEN, MARGIN_GUARD, UNIFIED_FIXTURES, load_unified, pack_positions

pytestmark = pytest.mark.gpu


def _run_episodes(sim, g, N):
    import torch
    W = g["map"].shape[1]
    T = int(g["max_steps"])
    for ep, E in enumerate(g["ep"]):
        sim.set_positions(*pack_positions([E["pos0"].astype(np.int32)], N))
        cells, cnt = sim.rollout(T, record=T)
        torch.cuda.synchronize()
        cells, cnt = cells.cpu().numpy()[0], cnt.cpu().numpy()[0]
        steps = sim.counters()[0][0]
        assert steps == int(g["steps"][ep]), (ep, steps)
        for t, want in enumerate(E["traj"]):
            assert cnt[t] == len(want), (ep, t)
            assert np.array_equal(cells[t, :cnt[t]], want[:, 0] * W + want[:, 1]), (ep, t)


@pytest.mark.parametrize("name", UNIFIED_FIXTURES)
def test_unified_reference_fixture(cuda_device, name):
    """Multi-episode learning run: tables carry over episodes (run_unified_critic_training.py:216-222).
    The draw key's episode id is the episode counter, as in the fixture generator."""
    from ffm_b200 import UnifiedSim

    g = load_unified(name)
    assert float(np.min(g["min_margin"])) >= MARGIN_GUARD
    N = max(len(E["pos0"]) for E in g["ep"])
    sims = []
    vinit = None
    if g["v_from"]:
        z = np.load(os.path.join(GOLDEN, g["v_from"] + ".npz"))
        vinit = (z["v_ids"], z["v_vals"])
    tables = None
    W = g["map"].shape[1]
    T = int(g["max_steps"])
    import torch
    for ep, E in enumerate(g["ep"]):
        # one handle per episode id (episode_base = ep) sharing the tables through get/set: exercises
        # the table round trip as well
        sim = UnifiedSim(g["map"], g["sff"], 1, N, mode=g["mode"], learn="exact", params={**g["params"], "epsilon": float(g["eps"])},
                         seed=int(g["seed"]), episode_base=ep)
        if tables is not None:
            sim.set_tables(*tables)
        elif vinit is not None:
            V = np.zeros(sim.S); vs = np.zeros(sim.S, np.uint8)
            V[vinit[0]] = vinit[1]; vs[vinit[0]] = 1
            sim.set_tables(V=V, v_seen=vs)
        sim.set_positions(*pack_positions([E["pos0"].astype(np.int32)], N))
        cells, cnt = sim.rollout(T, record=T)
        torch.cuda.synchronize()
        cells, cnt = cells.cpu().numpy()[0], cnt.cpu().numpy()[0]
        assert sim.counters()[0][0] == int(g["steps"][ep]), ep
        for t, want in enumerate(E["traj"]):
            assert cnt[t] == len(want), (ep, t)
            assert np.array_equal(cells[t, :cnt[t]], want[:, 0] * W + want[:, 1]), (ep, t)
        tables = sim.get_tables()
        last = sim
    V, vs, H, hs = tables
    assert np.array_equal(np.flatnonzero(vs), g["v_ids"])
    assert np.array_equal(V[g["v_ids"]].view(np.uint64), g["v_vals"].view(np.uint64)), "V table bits"
    assert np.array_equal(np.flatnonzero(hs), g["h_ids"])
    assert np.array_equal(H[g["h_ids"]].view(np.uint64), g["h_vals"].view(np.uint64)), "H table bits"
    assert np.array_equal(last.get_dff()[0].view(np.uint32), g["final_dff"].view(np.uint32))


def test_trained_reference_fixture(cuda_device):
    import torch
    from ffm_b200 import UnifiedSim

    z = np.load(os.path.join(GOLDEN, "trained_12x12.npz"))
    params = json.loads(str(z["params"]))
    hz = np.load(os.path.join(GOLDEN, str(z["h_from"]) + ".npz"))
    N, T, W = len(z["pos0"]), int(z["max_steps"]), z["map"].shape[1]
    sim = UnifiedSim(z["map"], z["sff"], 1, N, mode="trained", learn="none", params=params, seed=int(z["seed"]))
    H = np.zeros((sim.S, sim.A)); hs = np.zeros(sim.S, np.uint8)
    H[hz["h_ids"]] = hz["h_vals"]; hs[hz["h_ids"]] = 1
    sim.set_tables(H=H, h_seen=hs)
    sim.set_positions(*pack_positions([z["pos0"].astype(np.int32)], N))
    cells, cnt = sim.rollout(T, record=T)
    torch.cuda.synchronize()
    cells, cnt = cells.cpu().numpy()[0], cnt.cpu().numpy()[0]
    assert sim.counters()[0][0] == int(z["steps"])
    offs = np.concatenate([[0], np.cumsum(z["counts"])])
    for t in range(int(z["steps"])):
        want = z["traj"][offs[t]:offs[t + 1]].astype(np.int64)
        assert np.array_equal(cells[t, :cnt[t]], want[:, 0] * W + want[:, 1]), t
    assert np.array_equal(sim.get_dff()[0].view(np.uint32), z["final_dff"].view(np.uint32))


def test_frozen_policy_batch_matches_oracle(cuda_device):
    """learn="none": B independent episodes on frozen tables (critic_only dynamics) == oracle per episode."""
    import torch
    from ffm_b200 import UnifiedSim
    from helpers import random_positions
    from oracle import assets, unified_numpy
    from oracle.inject import PhiloxSource

    m = assets.room_map(14, 16)
    sff = assets.sff_norm_min(m, "L1", np.float32)
    params = {"k_S": 4, "k_D": 1, "neighborhood": "moore", "block_size": 2}
    rng = np.random.RandomState(1)
    B, N = 5, 40
    pos0 = [random_positions(m, N - 2 * e, rng) for e in range(B)]
    sim = UnifiedSim(m, sff, B, N, mode="critic_only", learn="none", params=params, seed=31, episode_base=7)
    sim.set_positions(*pack_positions(pos0, N))
    cells, cnt = sim.rollout(200, record=200)
    torch.cuda.synchronize()
    cells, cnt = cells.cpu().numpy(), cnt.cpu().numpy()
    steps = sim.counters()[0]
    for e in range(B):
        o = unified_numpy.UnifiedOracle(m, sff, pos0[e], "trained" if False else "critic_only", params, PhiloxSource(31, 7 + e))
        # frozen tables: the oracle's critic update never feeds back into the dynamics in critic_only mode
        r = o.run(max_steps=200)
        if r["min_margin"] < MARGIN_GUARD:
            continue
        assert steps[e] == r["steps"]
        for t, want in enumerate(r["traj"]):
            assert np.array_equal(cells[e, t, :cnt[e, t]], want[:, 0] * 16 + want[:, 1]), (e, t)
