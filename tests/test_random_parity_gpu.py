"""Randomised parity sweep of the base model: random obstacle maps (non-square, several exits), random crowd sizes,
both neighbourhoods, random k_S / k_D / diffuse / decay, float32 and float64 SFF, ragged batches -- every kernel variant
against the C restatement of the reference (pinned to the reference fixtures by tests/test_c_oracle.py) under the
recorded-draw protocol: per-step trajectories, pedestrian-step counts, final DFF bits."""
import numpy as np
import pytest

from helpers import pack_positions
from oracle import assets, c_oracle

pytestmark = pytest.mark.gpu


def _case(seed):
    rng = np.random.RandomState(1000 + seed)
    H, W = int(rng.randint(12, 72)), int(rng.randint(12, 90))
    if seed % 3 == 0:
        W = 4 * (W // 4)                                  # widths that take the vectorised DFF stencil
    m = assets.obstacle_map_c5(H, W, index=seed, fill=float(rng.uniform(0.0, 0.25)), n_exits=int(rng.choice([4, 8])))
    nbh = "neumann" if rng.rand() < 0.5 else "moore"
    sff = c_oracle.geodesic(m, "bfs4" if nbh == "neumann" else str(rng.choice(["bfs8", "dijkstra8"])))
    if rng.rand() < 0.4:
        sff = sff.astype(np.float64)
    reach = np.argwhere((m == 0) & np.isfinite(sff))
    B = int(rng.randint(1, 5))
    n_max = int(min(len(reach), rng.randint(1, 400)))
    counts = [int(rng.randint(0, n_max + 1)) for _ in range(B)]
    counts[0] = n_max
    pos0 = [reach[rng.choice(len(reach), k, replace=False)] if k else np.zeros((0, 2), np.int64) for k in counts]
    params = {"neighborhood": nbh, "k_S": float(rng.uniform(0.5, 6.0)), "k_D": float(rng.choice([0.0, 0.5, 1.0, 2.0])),
              "diffuse": float(rng.uniform(0.05, 0.5)), "decay": float(rng.uniform(0.05, 0.5))}
    return m, sff, pos0, n_max, params, int(rng.randint(1, 1 << 30)), int(rng.randint(0, 1000))


@pytest.mark.parametrize("seed", range(24))
def test_random_configuration_matches_the_oracle(cuda_device, core_kernel, seed):
    import torch
    from ffm_b200 import BatchSim
    m, sff, pos0, n_max, params, key, base = _case(seed)
    W = m.shape[1]
    pos, n = pack_positions(pos0, n_max)
    T = 90
    ref = c_oracle.run_core_batch(m, sff, pos, n, params, seed=key, episode_base=base, max_steps=T, threads=4, track_dff=True,
                                  traj_steps=T, want_state=True, guard=1e-5, record_moves=T)
    sim = BatchSim(m, sff, len(pos0), n_max, params, seed=key, episode_base=base, track_dff=True)
    sim.set_positions(pos, n)
    half = T // 3                                          # the state survives a relaunch
    c1, k1 = sim.rollout(half, draws=dict(move=torch.from_numpy(ref["move_draws"]).cuda()), record=half)
    c2, k2 = sim.rollout(T - half, draws=dict(move=torch.from_numpy(ref["move_draws"]).cuda()), record=T - half)
    torch.cuda.synchronize()
    cells = np.concatenate([c1.cpu().numpy(), c2.cpu().numpy()], axis=1)
    cnt = np.concatenate([k1.cpu().numpy(), k2.cpu().numpy()], axis=1)
    steps, ped = sim.counters()
    assert np.array_equal(steps, np.minimum(ref["steps"], T)) and np.array_equal(ped, ref["ped_steps"])
    for e in range(len(pos0)):
        s = int(steps[e])
        assert np.array_equal(cnt[e, :s], ref["traj_n"][e, :s]), e
        mask = np.arange(n_max)[None, :] < cnt[e, :s, None]
        assert np.array_equal(cells[e, :s][mask], ref["traj"][e, :s][mask]), e
    p_gpu, n_gpu = sim.get_positions()
    assert np.array_equal(n_gpu, ref["final_n"])
    for e in range(len(pos0)):
        assert np.array_equal(p_gpu[e, :n_gpu[e], 0] * W + p_gpu[e, :n_gpu[e], 1], ref["final_pos"][e, :n_gpu[e]])
    assert np.array_equal(sim.get_dff().view(np.uint32), ref["final_dff"].view(np.uint32))
