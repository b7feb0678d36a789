"""GPU parity of the base CA rollout (C ABI -> ffm_core_rollout_kernel) against the NumPy oracle
(restatement of model/ffm_core.py, itself pinned to the reference by tests/golden fixtures).

Bar: bit-exact trajectories, step counts and DFF fields under the same keyed draws."""
import numpy as np
import pytest

from helpers import MARGIN_GUARD, oracle_core_episode, pack_positions, random_positions, traj_to_cells
from oracle import assets

pytestmark = pytest.mark.gpu

CASES = [
    # h, w, N, neighbourhood, metric, sff dtype, params
    (12, 12, 50, "neumann", "L1", np.float32, {}),
    (12, 12, 100, "moore", "L1", np.float32, {}),
    (50, 50, 100, "moore", "L2", np.float64, {}),
    (50, 50, 100, "neumann", "L1", np.float64, {}),
    (20, 20, 150, "moore", "Linf", np.float32, {"k_S": 2.5, "k_D": 0.7, "diffuse": 0.3, "decay": 0.1}),
    (16, 24, 60, "moore", "L2", np.float32, {"k_D": 0}),
]


@pytest.mark.parametrize("h,w,N,nbh,metric,dtype,extra", CASES)
def test_rollout_matches_oracle(cuda_device, core_kernel, h, w, N, nbh, metric, dtype, extra):
    import torch
    from ffm_b200 import BatchSim

    B, seed, ep_base = 6, 0xC0FFEE + h * 131 + N, 40
    params = {"neighborhood": nbh, **extra}
    m = assets.room_map(h, w)
    sff = assets.sff_norm_min(m, metric, dtype)
    rng = np.random.RandomState(h * 1000 + N)
    pos0 = [random_positions(m, N - 3 * e, rng) for e in range(B)]   # ragged counts
    ref = [oracle_core_episode(m, sff, pos0[e], params, seed, ep_base + e) for e in range(B)]
    T = max(r["steps"] for r in ref) + 2

    sim = BatchSim(m, sff, B, N, params, seed=seed, episode_base=ep_base, track_dff=True)
    sim.set_positions(*pack_positions(pos0, N))
    cells, cnt = sim.rollout(T, record=T)
    torch.cuda.synchronize()
    steps, ped_steps = sim.counters()
    dff = sim.get_dff()
    cells, cnt = cells.cpu().numpy(), cnt.cpu().numpy()
    _, n_left = sim.get_positions()

    excused = 0
    for e in range(B):
        r = ref[e]
        if r["min_margin"] < MARGIN_GUARD:
            excused += 1
            continue
        assert steps[e] == r["steps"], f"episode {e}: steps {steps[e]} != {r['steps']}"
        assert n_left[e] == 0
        assert ped_steps[e] == len(pos0[e]) + sum(len(p) for p in r["traj"][:-1])
        for t, want in enumerate(traj_to_cells(r["traj"], w)):
            assert cnt[e, t] == len(want), f"episode {e} step {t}: count"
            got = cells[e, t, :cnt[e, t]].astype(np.int64)
            assert np.array_equal(got, want), f"episode {e} step {t}: trajectory differs"
        assert np.array_equal(dff[e].view(np.uint32), r["final_dff"].view(np.uint32)), f"episode {e}: DFF bits differ"
    assert excused <= 1
    sim.close()


def test_stepwise_equals_single_launch(cuda_device, core_kernel):
    """step() x T through separate launches (state round-trips through HBM) == one persistent launch."""
    import torch
    from ffm_b200 import BatchSim

    m = assets.room_map(14, 14)
    sff = assets.sff_norm_min(m, "L2", np.float32)
    rng = np.random.RandomState(5)
    pos0 = [random_positions(m, 40, rng) for _ in range(3)]
    a = BatchSim(m, sff, 3, 40, {}, seed=9)
    b = BatchSim(m, sff, 3, 40, {}, seed=9)
    a.set_positions(*pack_positions(pos0, 40))
    b.set_positions(*pack_positions(pos0, 40))
    a.rollout(25)
    for _ in range(25):
        b.rollout(1)
    torch.cuda.synchronize()
    pa, na = a.get_positions()
    pb, nb = b.get_positions()
    assert np.array_equal(na, nb) and np.array_equal(pa, pb)
    assert np.array_equal(a.get_dff().view(np.uint32), b.get_dff().view(np.uint32))
    assert np.array_equal(a.counters()[1], b.counters()[1])


def test_invalid_inputs_raise(cuda_device):
    from ffm_b200 import BatchSim

    m = assets.room_map(10, 10)
    sff = assets.sff_norm_min(m, "L1", np.float32)
    bad = m.copy(); bad[0, 3] = 0                       # free cell on the border
    with pytest.raises(ValueError):
        BatchSim(bad, sff, 1, 4, {})
    sim = BatchSim(m, sff, 1, 4, {})
    pos = np.array([[[0, 0], [1, 1], [2, 2], [3, 3]]], dtype=np.int32)   # (0,0) is a wall
    sim.set_positions(pos, np.array([4], dtype=np.int32))
    with pytest.raises(ValueError):
        sim.get_positions()
    with pytest.raises(ValueError):
        BatchSim(m, sff, 1, 4, {"k_D": 1}, track_dff=False)
