"""Full-size parity (BASELINE configs[1] geometry: 64x64, 1024 pedestrians, Moore) of the CUDA rollout
against the C oracle (oracle/c, pinned to the reference fixtures by tests/test_c_oracle.py), plus
size-independent properties of the rollout."""
import numpy as np
import pytest

from helpers import MARGIN_GUARD

pytestmark = pytest.mark.gpu


def _c2(nbh="moore"):
    import bench
    m = bench.room_map(64, 64)
    return m, bench.sff_room(m, nbh)


@pytest.mark.parametrize("k_D,track", [(0, False), (1, True)])
def test_c2_full_episodes_match_c_oracle(cuda_device, k_D, track):
    import bench
    import torch
    from ffm_b200 import BatchSim
    from oracle import c_oracle

    m, sff = _c2()
    B, N, seed, base = 48, 1024, 0x5EED0002, 4096 * 3
    params = {"k_S": 3, "k_D": k_D, "neighborhood": "moore"}
    pos = bench.place(m, N, B, base, seed)
    n = np.full((B,), N, np.int32)
    T = 4096
    ref = c_oracle.run_core_batch(m, sff, pos, n, params, seed=seed, episode_base=base, max_steps=T, threads=8,
                                  track_dff=track, traj_steps=T if False else 0, want_state=True)
    sim = BatchSim(m, sff, B, N, params, seed=seed, episode_base=base, track_dff=track)
    sim.set_positions(pos, n)
    rec = 8
    sim.rollout(T)
    torch.cuda.synchronize()
    steps, ped = sim.counters()
    ok = ref["min_margin"] >= MARGIN_GUARD
    assert ok.sum() >= B - 4, f"too many knife-edge episodes: {(~ok).sum()}"
    assert np.array_equal(steps[ok], ref["steps"][ok])
    assert np.array_equal(ped[ok], ref["ped_steps"][ok])
    assert (sim.get_positions()[1] == 0).all()
    if track:
        dff = sim.get_dff()
        for e in np.nonzero(ok)[0]:
            assert np.array_equal(dff[e].view(np.uint32), ref["final_dff"][e].view(np.uint32)), e


def test_c2_trajectories_match_c_oracle(cuda_device):
    import bench
    import torch
    from ffm_b200 import BatchSim
    from oracle import c_oracle

    m, sff = _c2()
    B, N, seed = 4, 1024, 77
    params = {"k_S": 3, "k_D": 1, "neighborhood": "moore"}
    pos = bench.place(m, N, B, 0, seed)
    n = np.full((B,), N, np.int32)
    T = 2400
    ref = c_oracle.run_core_batch(m, sff, pos, n, params, seed=seed, max_steps=T, threads=4, traj_steps=T, want_state=True)
    sim = BatchSim(m, sff, B, N, params, seed=seed)
    sim.set_positions(pos, n)
    cells, cnt = sim.rollout(T, record=T)
    torch.cuda.synchronize()
    cells, cnt = cells.cpu().numpy(), cnt.cpu().numpy()
    for e in range(B):
        if ref["min_margin"][e] < MARGIN_GUARD:
            continue
        s = int(ref["steps"][e])
        assert np.array_equal(cnt[e, :s], ref["traj_n"][e, :s])
        for t in range(s):
            k = cnt[e, t]
            assert np.array_equal(cells[e, t, :k], ref["traj"][e, t, :k]), (e, t)


def test_properties_at_full_size(cuda_device):
    """Size-independent invariants: conservation (nobody appears, exits are monotone), positions stay on
    distinct free cells, sharding/launch-splitting independence, determinism."""
    import bench
    import torch
    from ffm_b200 import BatchSim

    m, sff = _c2()
    B, N, seed = 16, 1024, 5
    params = {"k_S": 3, "k_D": 1, "neighborhood": "moore"}
    pos = bench.place(m, N, B, 100, seed)
    n = np.full((B,), N, np.int32)
    a = BatchSim(m, sff, B, N, params, seed=seed, episode_base=100)
    a.set_positions(pos, n)
    prev = n.copy()
    for chunk in range(6):
        a.rollout(100)
        p, k = a.get_positions()
        assert (k <= prev).all()
        prev = k
        for e in range(B):
            q = p[e, :k[e]]
            assert (m[q[:, 0], q[:, 1]] == 0).all()
            assert len({(int(x), int(y)) for x, y in q}) == k[e]
    # two half-batches with the right episode_base == the full batch (keys are global episode ids)
    b = BatchSim(m, sff, B // 2, N, params, seed=seed, episode_base=100 + B // 2)
    b.set_positions(pos[B // 2:], n[B // 2:])
    b.rollout(600)
    pb, kb = b.get_positions()
    assert np.array_equal(kb, k[B // 2:]) and np.array_equal(pb, p[B // 2:])
    assert np.array_equal(b.get_dff().view(np.uint32), a.get_dff()[B // 2:].view(np.uint32))
