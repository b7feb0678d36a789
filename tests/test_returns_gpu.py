"""Rollout buffer of the unified kernel + discounted returns (ffm_rollout_returns) against NumPy."""
import numpy as np
import pytest

from helpers import pack_positions, random_positions
from oracle import assets

pytestmark = pytest.mark.gpu


def np_returns(reward, length, gamma):
    """ffm_learning_core.py:262-278: G = r + gamma * G walking each path backwards (Python floats)."""
    B, T, N = reward.shape
    G = np.zeros((B, T, N), np.float64)
    for b in range(B):
        for n in range(N):
            g = 0.0
            for t in range(min(int(length[b, n]), T) - 1, -1, -1):
                g = float(reward[b, t, n]) + gamma * g
                G[b, t, n] = g
    return G


@pytest.mark.parametrize("shape", [(3, 17, 8), (2, 33, 7), (5, 64, 100)])
def test_returns_bit_exact_on_ragged_paths(cuda_device, shape):
    import torch
    from ffm_b200.sim import rollout_returns
    rng = np.random.RandomState(sum(shape))
    B, T, N = shape
    reward = (rng.randn(B, T, N) * 3).astype(np.float32)
    length = rng.randint(0, T + 1, size=(B, N)).astype(np.int32)
    length[0, 0] = 0; length[-1, -1] = T
    G = rollout_returns(torch.from_numpy(reward).cuda(), torch.from_numpy(length).cuda(), 0.99).cpu().numpy()
    assert np.array_equal(G.view(np.uint64), np_returns(reward, length, 0.99).view(np.uint64))


def test_rollout_buffer_matches_oracle(cuda_device):
    """state / action / reward columns recorded by the unified kernel vs the oracle's per-agent bookkeeping.
    Compaction is stable, so the agents alive at step t, in array order, are the columns with length > t in
    ascending order."""
    import torch
    from ffm_b200 import UnifiedSim
    from ffm_b200.sim import rollout_returns
    from oracle import unified_numpy
    from oracle.inject import PhiloxSource

    m = assets.room_map(12, 12)
    sff = assets.sff_norm_min(m, "L1", np.float32)
    P = dict(k_S=10, k_D=1, gamma=0.99, alpha_v=0.01, exit_reward=100.0, step_penalty=-1.0, collision_penalty=-1.0,
             neighborhood="neumann", block_size=1)
    rng = np.random.RandomState(4)
    N, T = 25, 80
    pos0 = random_positions(m, N, rng)
    sim = UnifiedSim(m, sff, 1, N, mode="critic_only", learn="none", params=P, seed=17)
    sim.set_positions(*pack_positions([pos0], N))
    buf = sim.rollout(T, record=T, record_buffer=True)
    torch.cuda.synchronize()
    length = buf["length"].cpu().numpy()[0]
    reward, state, action = (buf[k].cpu().numpy()[0] for k in ("reward", "state", "action"))
    o = unified_numpy.UnifiedOracle(m, sff, pos0, "critic_only", P, PhiloxSource(17, 0))
    o.step_log = []
    r = o.run(max_steps=T)
    assert r["min_margin"] > 2e-6 and r["steps"] == int(length.max())
    for t, log in enumerate(o.step_log):
        cols = np.flatnonzero(length > t)
        assert len(cols) == len(log["states"]), t
        assert np.array_equal(state[t, cols], log["states"]), t
        assert np.array_equal(action[t, cols], log["slots"]), t
        assert np.array_equal(reward[t, cols], log["rewards"].astype(np.float32)), t
    # every finished path ends with the exit reward (step_penalty + exit_reward [+ collisions])
    last = reward[length - 1, np.arange(N)]
    assert (last >= 99.0 - 8).all()
    G = rollout_returns(buf["reward"], buf["length"], 0.99).cpu().numpy()
    assert np.array_equal(G.view(np.uint64), np_returns(buf["reward"].cpu().numpy(), buf["length"].cpu().numpy(), 0.99).view(np.uint64))
