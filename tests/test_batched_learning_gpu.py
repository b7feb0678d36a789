"""Batched (synchronous) TD learning -- BASELINE config 4 at one GPU -- judged statistically against
the reference's sequential semantics (FFM_LEARN_EXACT), plus the 2N-1 saturated-exit law
(analyze_steps_by_n.py:109-110, output/logs/unified_critic_training/*/summary.txt)."""
import numpy as np
import pytest

from helpers import pack_positions
from oracle import assets

pytestmark = pytest.mark.gpu

PARAMS = dict(k_S=10, k_D=1, k_A=10, alpha_v=0.01, alpha_h=0.1, gamma=0.99, exit_reward=100.0, step_penalty=-1.0,
              collision_penalty=-1.0, neighborhood="neumann", block_size=1)


def _placements(m, N, episodes, radius, rng):
    free = np.argwhere(m == 0)
    near = free[np.abs(free[:, 0] - 0) + np.abs(free[:, 1] - 6) <= radius]
    return [near[rng.choice(len(near), min(N, len(near)), replace=False)] for _ in range(episodes)]


def test_saturated_exit_law_2n_minus_1(cuda_device):
    """critic_only dynamics, radius 15: mean evacuation time ~ 2N-1 + travel (N=50 -> 99.56, N=90 -> 179.10 in
    the reference's logs); bands [2N-1, 2N-1+15] from analyze_steps_by_n.py."""
    from ffm_b200 import UnifiedSim
    m = assets.room_map(12, 12)
    sff = assets.sff_norm_min(m, "L1", np.float32)
    rng = np.random.RandomState(0)
    for N, logged in ((50, 99.56), (90, 179.10)):
        B = 256
        pos = _placements(m, N, B, 15, rng)
        sim = UnifiedSim(m, sff, B, N, mode="critic_only", learn="none", params=PARAMS, seed=N)
        sim.set_positions(*pack_positions(pos, N))
        sim.rollout(400)
        steps, _ = sim.counters()
        assert (sim.get_positions()[1] == 0).all()
        assert abs(steps.mean() - logged) < 1.0, (N, steps.mean())
        assert ((steps >= 2 * N - 1) & (steps <= 2 * N - 1 + 15)).mean() > 0.95


def test_batched_td_tracks_sequential_td(cuda_device):
    """Same episodes, same draws: V learned by synchronous batched TD (64 episodes per sync) vs the reference's
    sequential updates.  The two are different algorithms; on well-visited states they must agree closely."""
    from ffm_b200 import UnifiedSim
    from ffm_b200.sharding import BatchedLearner
    m = assets.room_map(12, 12)
    sff = assets.sff_norm_min(m, "L1", np.float32)
    rng = np.random.RandomState(1)
    N, E, B = 20, 512, 64
    pos = _placements(m, N, E, 7, rng)
    exact = UnifiedSim(m, sff, 1, N, mode="critic_only", learn="exact", params=PARAMS, seed=9)
    for e in range(E):
        exact.set_episode_base(e)
        exact.set_positions(*pack_positions([pos[e]], N))
        exact.rollout(300)
    Ve, se, _, _ = exact.get_tables()
    batched = UnifiedSim(m, sff, B, N, mode="critic_only", learn="batched", params=PARAMS, seed=9)
    learner = BatchedLearner(batched)
    for r in range(E // B):
        batched.set_episode_base(r * B)
        learner.round(*pack_positions(pos[r * B:(r + 1) * B], N), 96, sync_every=4)
    Vb, sb, _, _ = batched.get_tables()
    assert np.array_equal(se, sb)                       # the dynamics do not depend on V: same states visited
    big = se & (np.abs(Ve) > 5.0)
    assert big.sum() > 50
    corr = np.corrcoef(Ve[big], Vb[big])[0, 1]
    assert corr > 0.97, corr
    rel = np.abs(Ve[big] - Vb[big]) / np.abs(Ve[big])
    assert np.median(rel) < 0.2, np.median(rel)


def test_batched_actor_learning_runs_and_keeps_tables_finite(cuda_device):
    from ffm_b200 import UnifiedSim
    from ffm_b200.sharding import BatchedLearner
    m = assets.room_map(12, 12)
    sff = assets.sff_norm_min(m, "L1", np.float32)
    rng = np.random.RandomState(2)
    N, B = 10, 128
    sim = UnifiedSim(m, sff, B, N, mode="both", learn="batched", params={**PARAMS, "epsilon": 0.2}, seed=3)
    learner = BatchedLearner(sim)
    means = []
    for r in range(12):
        sim.set_episode_base(r * B)
        steps, _ = learner.round(*pack_positions(_placements(m, N, B, 5, rng), N), 300)
        means.append(steps.mean())
    V, vs, H, hs = sim.get_tables()
    assert np.isfinite(V).all() and np.isfinite(H).all() and hs.sum() > 100 and np.abs(H[hs]).max() > 0
    assert means[-1] <= means[0] * 1.5               # learning does not blow the policy up


def test_unified_evacuation_time_distribution_matches_oracle(cuda_device):
    """Independent seeds: evacuation-time distribution of the CUDA unified model (critic_only dynamics) vs the NumPy
    oracle of the reference -- two-sample KS bound 0.25 at n = 256 vs 40 (p ~ 0.02), means within 3 %."""
    from scipy import stats
    from ffm_b200 import UnifiedSim
    from oracle import unified_numpy
    from oracle.inject import PhiloxSource
    m = assets.room_map(12, 12)
    sff = assets.sff_norm_min(m, "L1", np.float32)
    N = 30
    rng = np.random.RandomState(7)
    sim = UnifiedSim(m, sff, 256, N, mode="critic_only", learn="none", params=PARAMS, seed=1001)
    sim.set_positions(*pack_positions(_placements(m, N, 256, 15, rng), N))
    sim.rollout(400)
    gpu_steps = sim.counters()[0]
    ref_steps = []
    for e, pos0 in enumerate(_placements(m, N, 40, 15, rng)):
        o = unified_numpy.UnifiedOracle(m, sff, pos0, "critic_only", PARAMS, PhiloxSource(2002, e))
        ref_steps.append(o.run(max_steps=400)["steps"])
    ref_steps = np.array(ref_steps)
    assert stats.ks_2samp(gpu_steps, ref_steps).statistic < 0.25
    assert abs(gpu_steps.mean() - ref_steps.mean()) < 0.03 * ref_steps.mean()
