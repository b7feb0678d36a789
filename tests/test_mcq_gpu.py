"""CUDA Monte-Carlo Q-learning model (C ABI -> ffm_mcq_rollout_kernel) against fixtures from the unmodified
model/ffm_learning_core.py: trajectories of multi-episode runs with a shared Q table, the final Q rows (float32 bits)
and the DFF."""
import numpy as np
import pytest

from helpers import MARGIN_GUARD, MCQ_FIXTURES, load_mcq, pack_positions

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", MCQ_FIXTURES)
def test_mcq_reference_fixture(cuda_device, name):
    import torch
    from ffm_b200 import McqSim
    g = load_mcq(name)
    assert float(np.min(g["min_margin"])) >= MARGIN_GUARD
    N = max(len(E["pos0"]) for E in g["ep"])
    W = g["map"].shape[1]
    cap = int(g["params"]["max_steps"])
    sim = McqSim(g["map"], g["sff"], 1, N, learn="exact", params=g["params"], seed=int(g["seed"]),
                 alpha=float(g["alpha"]), gamma=float(g["gamma"]))
    for ep, E in enumerate(g["ep"]):
        sim.set_episode_base(ep)
        sim.set_beta(float(g["betas"][ep]))
        sim.set_positions(*pack_positions([E["pos0"].astype(np.int32)], N))
        T = cap + 2
        half = int(g["steps"][ep]) // 2                       # two launches: the path bookkeeping survives a relaunch
        c1, n1 = sim.rollout(half, record=max(half, 1))
        c2, n2 = sim.rollout(T, record=T)
        torch.cuda.synchronize()
        cells = np.concatenate([c1.cpu().numpy()[0][:half], c2.cpu().numpy()[0]])
        cnt = np.concatenate([n1.cpu().numpy()[0][:half], n2.cpu().numpy()[0]])
        assert sim.counters()[0][0] == int(g["steps"][ep]), ep
        for t, want in enumerate(E["traj"]):
            assert cnt[t] == len(want), (ep, t)
            assert np.array_equal(cells[t, :cnt[t]], want[:, 0] * W + want[:, 1]), (ep, t)
    ids, rows = sim.get_q()
    assert np.array_equal(ids, g["q_ids"])
    assert np.array_equal(rows.view(np.uint32), g["q_rows"].view(np.uint32)), "Q rows (float32 bits)"
    assert np.array_equal(sim.get_dff()[0].view(np.uint32), g["final_dff"].view(np.uint32))


def test_mcq_frozen_table_batch(cuda_device):
    """learn="none": many episodes on a frozen Q table == the oracle per episode (which never backs up here because
    the comparison stops before anybody arrives)."""
    import torch
    from ffm_b200 import McqSim
    from helpers import random_positions
    from oracle import assets, mcq_numpy
    from oracle.inject import PhiloxSource
    m = assets.room_map(15, 15)
    sff = assets.sff_norm_min(m, "L1", np.float64)
    params = {"max_steps": 40, "k_S": 1.0}
    rng = np.random.RandomState(0)
    B, N = 4, 25
    pos0 = [random_positions(m, N, rng) for _ in range(B)]
    sim = McqSim(m, sff, B, N, learn="none", params=params, seed=77, episode_base=3)
    sim.set_beta(0.7)
    sim.set_positions(*pack_positions(pos0, N))
    cells, cnt = sim.rollout(45, record=45)
    torch.cuda.synchronize()
    cells, cnt = cells.cpu().numpy(), cnt.cpu().numpy()
    steps = sim.counters()[0]
    for e in range(B):
        o = mcq_numpy.McqOracle(m, sff, pos0[e], params, PhiloxSource(77, 3 + e))
        o.alpha = 0.0                                          # frozen table: backups change nothing
        r = o.run(0.7)
        if r["min_margin"] < MARGIN_GUARD:
            continue
        assert steps[e] == r["steps"] == 40                    # nobody is left after the timeout
        for t, want in enumerate(r["traj"]):
            assert np.array_equal(cells[e, t, :cnt[e, t]], want[:, 0] * 15 + want[:, 1]), (e, t)
