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


from helpers import PRETRAIN_FIXTURES, load_pretrain


@pytest.mark.parametrize("name", PRETRAIN_FIXTURES)
def test_coverage_pretrain_matches_the_reference_driver(cuda_device, name):
    """All patterns of coverage_pretrain_empty (run_coverage_pretrain_and_training.py:173-216) as ONE launch + the ordered
    backup pass: the shared Q of the unmodified driver, keys and float32 bits, and every mini-episode's step count."""
    from ffm_b200.mcq_training import coverage_pretrain
    g = load_pretrain(name)
    assert float(g["min_margin"]) >= MARGIN_GUARD
    Q, steps = coverage_pretrain(g["map"], g["sff"], g["params"], {}, order=g["order"], seed=int(g["seed"]), return_steps=True)
    assert np.array_equal(steps, g["steps"])
    from oracle import mcq_numpy
    helper = mcq_numpy.McqOracle(g["map"], g["sff"], np.zeros((0, 2)), g["params"])
    ids = np.array(sorted(helper.id_of(k) for k in Q), np.int64)
    assert np.array_equal(ids, g["q_ids"])
    rows = np.stack([Q[helper.key_of(i)] for i in ids])
    assert np.array_equal(rows.view(np.uint32), g["q_rows"].view(np.uint32)), "Q rows (float32 bits)"
    # a second pretrain pass continues from the shared dict (Q is loaded, updated, returned)
    Q2 = coverage_pretrain(g["map"], g["sff"], g["params"], dict(Q), order=g["order"][:50], seed=int(g["seed"]))
    o, _ = mcq_numpy.coverage_pretrain(g["map"], g["sff"], g["params"], g["order"], int(g["seed"]), float(g["alpha"]), float(g["gamma"]))
    from oracle.inject import PhiloxSource
    for k, (tx, ty, a) in enumerate(g["order"][:50]):
        o.source = PhiloxSource(int(g["seed"]), k)
        mcq_numpy.force_first_step_and_roll(o, (int(tx), int(ty)), int(a))
    want = o.q_dict()
    assert set(Q2) == set(want) and all(np.array_equal(Q2[k].view(np.uint32), want[k].view(np.uint32)) for k in want)


def test_batched_mc_learning_fold(cuda_device):
    """learn="batched" + accumulate + fold: per (state, action) the table moves by (1 - (1 - alpha)^n)(mean G - Q); checked
    against returns collected from the oracle on the same keyed episodes (beta = 1: the policy does not read Q)."""
    from ffm_b200 import McqSim
    from ffm_b200.mcq_training import McqBatchedLearner
    from helpers import random_positions
    from oracle import assets, mcq_numpy
    from oracle.inject import PhiloxSource
    m = assets.room_map(12, 12)
    sff = assets.sff_norm_min(m, "L1", np.float64)
    params = {"max_steps": 60, "step_penalty": 0.02, "stop_penalty": 0.1, "collision_penalty": 0.5}
    alpha, gamma, B, N = 0.25, 0.97, 24, 12
    rng = np.random.RandomState(1)
    pos0 = [random_positions(m, N, rng) for _ in range(B)]
    sim = McqSim(m, sff, B, N, learn="batched", params=params, seed=99, alpha=alpha, gamma=gamma, q_log2_capacity=14)
    sim.set_beta(1.0)
    sim.set_positions(*pack_positions(pos0, N))
    sim.rollout(61)
    McqBatchedLearner(sim, distributed=False).sync()
    got = sim.q_dict()

    class Collect(mcq_numpy.McqOracle):
        def _backup(self, path):
            G = 0.0
            for sid, ac, r in reversed(path):
                G = r + self.gamma * G
                self.q_seen[sid] = True
                self.sums.setdefault((sid, ac), []).append(G)

    sums, knife = {}, False
    for e in range(B):
        o = Collect(m, sff, pos0[e], params, PhiloxSource(99, e), alpha, gamma)
        o.sums = sums
        knife |= o.run(1.0)["min_margin"] < MARGIN_GUARD
    assert not knife
    helper = mcq_numpy.McqOracle(m, sff, np.zeros((0, 2)), params)
    want = {}
    for (sid, ac), gs in sums.items():
        row = want.setdefault(helper.key_of(sid), np.zeros(5))
        row[ac] = (1.0 - (1.0 - alpha) ** len(gs)) * np.mean(gs)
    assert set(want) <= set(got)                                  # rows ensured at decision time may stay zero rows
    for k, row in want.items():
        assert np.allclose(got[k], row, rtol=2e-6, atol=1e-6), k
    assert all(not got[k].any() for k in set(got) - set(want))
    # a second sync without new episodes changes nothing (the deltas were cleared)
    sim.fold()
    again = sim.q_dict()
    assert all(np.array_equal(again[k], got[k]) for k in got)


def test_mcq_training_schedule_runs_batched(cuda_device):
    """run_training: the N ramp / beta schedule of run_coverage_pretrain_and_training.py:313-333 with a batch of episodes per
    entry, starting from a coverage-pretrained table: the table stays finite, grows, and exit-adjacent rows carry the exit
    reward's scale."""
    from ffm_b200.mcq_training import coverage_pretrain, run_training
    from oracle import assets
    m = assets.room_map(12, 12)
    sff = assets.sff_norm_min(m, "L1", np.float64)
    params = {"max_steps": 120, "alpha": 0.1, "gamma": 0.99, "step_penalty": 0.01}
    Q0 = coverage_pretrain(m, sff, params, {}, shuffle=False, seed=3)
    Q, mean_steps = run_training(m, sff, params, full_N=40, shared_Q=Q0, num_episodes=12, batch=32, seed=4)
    assert len(mean_steps) == 12 and all(0 < s <= 120 for s in mean_steps)
    assert set(Q0) <= set(Q) and len(Q) > len(Q0)
    rows = np.stack(list(Q.values()))
    assert np.isfinite(rows).all() and rows.max() > 50.0 and rows.max() <= 100.0 + 1e-3


def test_q_table_overflow_is_reported(cuda_device):
    """A hash table too small for the visited states reports it (load factor above 1/2) instead of corrupting rows."""
    from ffm_b200 import McqSim
    from ffm_b200._abi import FfmError
    from helpers import random_positions
    from oracle import assets
    m = assets.room_map(20, 20)
    sff = assets.sff_norm_min(m, "L1", np.float64)
    rng = np.random.RandomState(0)
    B, N = 16, 60
    sim = McqSim(m, sff, B, N, learn="batched", params={"max_steps": 100}, seed=5, q_log2_capacity=10)
    assert sim.q_capacity == 1024
    sim.set_beta(1.0)
    sim.set_positions(*pack_positions([random_positions(m, N, rng) for _ in range(B)], N))
    sim.rollout(100)
    with pytest.raises(FfmError, match="half full"):
        sim.get_q()
    with pytest.raises(Exception):
        McqSim(m, sff, 1, 4, learn="none", seed=1, q_log2_capacity=40)
