"""Randomised parity sweeps of the table models against their NumPy oracles (pinned to the reference fixtures by
tests/test_unified_oracle.py / tests/test_mcq_oracle.py): random obstacle maps, crowd sizes, learning modes, block sizes,
reward / learning parameters, epsilon; several episodes sharing the tables.  Trajectories (unless the oracle saw a
knife-edge draw), the V / H / Q tables bit for bit, the DFF."""
import numpy as np
import pytest

from helpers import MARGIN_GUARD, pack_positions
from oracle import assets, c_oracle, mcq_numpy, unified_numpy
from oracle.inject import PhiloxSource

pytestmark = pytest.mark.gpu


def _map(rng, seed):
    H, W = int(rng.randint(12, 26)), int(rng.randint(12, 30))
    m = assets.obstacle_map_c5(H, W, index=100 + seed, fill=float(rng.uniform(0.0, 0.15)), n_exits=4)
    return m, np.argwhere((m == 0) & np.isfinite(c_oracle.geodesic(m, "bfs4")))


@pytest.mark.parametrize("seed", range(10))
def test_unified_random_configuration(cuda_device, seed):
    import torch
    from ffm_b200 import UnifiedSim
    rng = np.random.RandomState(7000 + seed)
    m, reach = _map(rng, seed)
    W = m.shape[1]
    nbh = "neumann" if rng.rand() < 0.6 else "moore"
    sff = c_oracle.geodesic(m, "bfs4" if nbh == "neumann" else "bfs8")
    mode = str(rng.choice(["critic_only", "actor_only", "both"]))
    if mode == "critic_only" and rng.rand() < 0.5:
        sff = sff.astype(np.float64)                       # critic_only keeps the file dtype (ffm_unified.py:70)
    params = dict(k_S=float(rng.uniform(2, 10)), k_D=float(rng.choice([0.5, 1.0])), k_A=float(rng.uniform(2, 10)),
                  alpha_v=float(rng.uniform(0.01, 0.3)), alpha_h=float(rng.uniform(0.05, 0.3)), gamma=float(rng.uniform(0.9, 0.99)),
                  exit_reward=100.0, step_penalty=float(rng.choice([0.0, -1.0])), collision_penalty=-1.0, neighborhood=nbh,
                  block_size=int(rng.randint(1, 6)), epsilon=float(rng.choice([0.0, 0.1])) if mode != "critic_only" else 0.0,
                  diffuse=float(rng.uniform(0.1, 0.3)), decay=float(rng.uniform(0.1, 0.3)))
    N = int(min(len(reach), rng.randint(3, 40)))
    T, key = 45, int(rng.randint(1, 1 << 30))
    o = unified_numpy.UnifiedOracle(m, sff, np.zeros((0, 2)), mode, params)
    sim = UnifiedSim(m, sff, 1, N, mode=mode, learn="exact", params=params, seed=key)
    knife = False
    for ep in range(3):
        pos0 = reach[rng.choice(len(reach), N, replace=False)]
        o.positions, o.t, o.source = pos0.astype(np.int64), 0, PhiloxSource(key, ep)
        o.dff[:] = 0
        r = o.run(max_steps=T)
        knife |= r["min_margin"] < MARGIN_GUARD
        sim.set_episode_base(ep)
        sim.set_positions(*pack_positions([pos0.astype(np.int32)], N))
        cells, cnt = sim.rollout(T, record=T)
        torch.cuda.synchronize()
        if knife:
            pytest.skip("the oracle saw a draw on a CDF knife edge")
        cells, cnt = cells.cpu().numpy()[0], cnt.cpu().numpy()[0]
        assert sim.counters()[0][0] == r["steps"], ep
        for t, want in enumerate(r["traj"]):
            assert cnt[t] == len(want) and np.array_equal(cells[t, :cnt[t]], want[:, 0] * W + want[:, 1]), (ep, t)
    V, vs, H, hs = sim.get_tables()
    assert np.array_equal(vs, o.v_seen) and np.array_equal(V.view(np.uint64), o.V.view(np.uint64)), "V table bits"
    assert np.array_equal(hs, o.h_seen) and np.array_equal(H.view(np.uint64), o.H.view(np.uint64)), "H table bits"
    assert np.array_equal(sim.get_dff()[0].view(np.uint32), o.dff.view(np.uint32))


@pytest.mark.parametrize("seed", range(8))
def test_mcq_random_configuration(cuda_device, seed):
    import torch
    from ffm_b200 import McqSim
    rng = np.random.RandomState(9000 + seed)
    m, reach = _map(rng, seed)
    W = m.shape[1]
    sff = c_oracle.geodesic(m, "bfs4").astype(np.float64 if rng.rand() < 0.5 else np.float32)
    params = {"max_steps": int(rng.randint(20, 50)), "k_S": float(rng.uniform(1, 4)), "k_D": float(rng.choice([0.5, 1.0])),
              "k_Q": float(rng.uniform(0.5, 2.0)), "step_penalty": float(rng.choice([0.0, 0.02])), "stop_penalty": float(rng.choice([0.0, 0.2])),
              "collision_penalty": float(rng.choice([0.0, 0.5])), "timeout_penalty": float(rng.uniform(10, 60)),
              "diffuse": float(rng.uniform(0.1, 0.3)), "decay": float(rng.uniform(0.1, 0.3))}
    alpha, gamma = float(rng.uniform(0.05, 0.3)), float(rng.uniform(0.9, 0.99))
    N = int(min(len(reach), rng.randint(3, 30)))
    key = int(rng.randint(1, 1 << 30))
    o = mcq_numpy.McqOracle(m, sff, np.zeros((0, 2)), params, None, alpha, gamma)
    sim = McqSim(m, sff, 1, N, learn="exact", params=params, seed=key, alpha=alpha, gamma=gamma, q_log2_capacity=16)
    cap = params["max_steps"] + 2
    for ep, beta in enumerate((1.0, float(rng.uniform(0.2, 0.8)), 0.0)):
        pos0 = reach[rng.choice(len(reach), N, replace=False)]
        o.reset(pos0)
        o.source = PhiloxSource(key, ep)
        r = o.run(beta)
        sim.set_episode_base(ep)
        sim.set_beta(beta)
        sim.set_positions(*pack_positions([pos0.astype(np.int32)], N))
        cells, cnt = sim.rollout(cap, record=cap)
        torch.cuda.synchronize()
        if o.min_margin < MARGIN_GUARD:
            pytest.skip("the oracle saw a draw on a CDF knife edge")
        cells, cnt = cells.cpu().numpy()[0], cnt.cpu().numpy()[0]
        assert sim.counters()[0][0] == r["steps"], ep
        for t, want in enumerate(r["traj"]):
            assert cnt[t] == len(want) and np.array_equal(cells[t, :cnt[t]], want[:, 0] * W + want[:, 1]), (ep, t)
    ids, rows = sim.get_q()
    assert np.array_equal(ids, np.flatnonzero(o.q_seen))
    assert np.array_equal(rows.view(np.uint32), o.Q[ids].view(np.uint32)), "Q rows (float32 bits)"
    assert np.array_equal(sim.get_dff()[0].view(np.uint32), o.dff.view(np.uint32))
