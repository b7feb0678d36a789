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
def test_c2_recorded_draws_bit_exact(cuda_device, k_D, track):
    """North-star protocol at full size: both sides consume the same RECORDED move uniforms (recorder
    re-draws anything within 1e-5 of a CDF boundary); conflicts use the keyed streams.
    Every episode must then agree bit for bit: steps, ped-steps, trajectories, DFF."""
    import bench
    import torch
    from ffm_b200 import BatchSim
    from oracle import c_oracle

    m, sff = _c2()
    B, N, seed, base, T = 6, 1024, 0x5EED0002, 4096 * 3, 2600
    params = {"k_S": 3, "k_D": k_D, "neighborhood": "moore"}
    pos = bench.place(m, N, B, base, seed)
    n = np.full((B,), N, np.int32)
    ref = c_oracle.run_core_batch(m, sff, pos, n, params, seed=seed, episode_base=base, max_steps=T, threads=6,
                                  track_dff=track, traj_steps=T, want_state=True, guard=1e-5, record_moves=T)
    assert (ref["steps"] < T).all() and (ref["min_margin"] >= 1e-5).all()
    sim = BatchSim(m, sff, B, N, params, seed=seed, episode_base=base, track_dff=track)
    sim.set_positions(pos, n)
    cells, cnt = sim.rollout(T, draws=dict(move=torch.from_numpy(ref["move_draws"]).cuda()), record=T)
    torch.cuda.synchronize()
    steps, ped = sim.counters()
    assert np.array_equal(steps, ref["steps"]) and np.array_equal(ped, ref["ped_steps"])
    cells, cnt = cells.cpu().numpy(), cnt.cpu().numpy()
    for e in range(B):
        s = int(ref["steps"][e])
        assert np.array_equal(cnt[e, :s], ref["traj_n"][e, :s])
        mask = np.arange(N)[None, :] < cnt[e, :s, None]
        assert np.array_equal(cells[e, :s][mask], ref["traj"][e, :s][mask]), e
    if track:
        assert np.array_equal(sim.get_dff().view(np.uint32), ref["final_dff"].view(np.uint32))


def test_c2_keyed_streams_agree_up_to_knife_edges(cuda_device):
    """Pure Philox mode at full size (~4.5e5 draws per episode): an episode may differ from the oracle
    only if the oracle saw a draw closer than MARGIN_GUARD to a CDF boundary (exp differs by <= 2 ulp
    between libm and CUDA); the evacuation-time distributions must agree (KS)."""
    import bench
    import torch
    from scipy import stats
    from ffm_b200 import BatchSim
    from oracle import c_oracle

    m, sff = _c2()
    B, N, seed, base = 96, 1024, 0x5EED0002, 0
    params = {"k_S": 3, "k_D": 0, "neighborhood": "moore"}
    pos = bench.place(m, N, B, base, seed)
    n = np.full((B,), N, np.int32)
    ref = c_oracle.run_core_batch(m, sff, pos, n, params, seed=seed, episode_base=base, max_steps=4096, threads=8,
                                  track_dff=False, want_state=True)
    sim = BatchSim(m, sff, B, N, params, seed=seed, episode_base=base, track_dff=False)
    sim.set_positions(pos, n)
    sim.rollout(4096)
    torch.cuda.synchronize()
    steps, ped = sim.counters()
    differ = (steps != ref["steps"]) | (ped != ref["ped_steps"])
    knife = ref["min_margin"] < MARGIN_GUARD
    assert not (differ & ~knife).any(), "an episode without a knife-edge draw differs from the oracle"
    assert differ.mean() < 0.5
    # independent seeds: evacuation-time distribution, KS bound 0.2 at n = 96 vs 96 (p ~ 0.04 level)
    sim2 = BatchSim(m, sff, B, N, params, seed=seed + 1, episode_base=10_000, track_dff=False)
    sim2.set_positions(bench.place(m, N, B, 10_000, seed + 1), n)
    sim2.rollout(4096)
    steps2, _ = sim2.counters()
    assert stats.ks_2samp(steps2, ref["steps"]).statistic < 0.2
    assert abs(steps2.mean() - ref["steps"].mean()) < 0.02 * ref["steps"].mean()


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


@pytest.mark.parametrize("variant,B,T", [("auto", 2, 200), ("ped", 2, 200), ("cl4smem", 2, 200), ("cl8", 2, 200), ("auto", 8, 1000),
                                         ("cl4smem", 8, 1000), ("ped", 8, 1000)])
def test_c3_floor_plan_matches_c_oracle(cuda_device, monkeypatch, variant, B, T):
    """BASELINE configs[2] geometry: 256x256 rooms-and-doors plan, geodesic SFF (generated on the GPU), DFF on,
    10 000 pedestrians.  "auto" = what ffm_create picks by measurement: the cell-centric kernel as a 2-CTA cluster (row
    bands of the owner grid / claim masks in distributed shared memory, score + DFF in L2); "cl4smem" = 4 CTAs with the
    DFF ping-pong on chip as well; "cl8" = 8 CTAs; "ped" = the pedestrian-centric kernel (one CTA, fields in L2).
    Recorded-draw protocol; the long cases cover 8 episodes x 1000 steps."""
    import torch
    from ffm_b200 import BatchSim
    from ffm_b200.sff import generate_sff
    from ffm_b200.workloads import place, rooms_map_c3
    from oracle import c_oracle

    if variant == "ped":
        monkeypatch.setenv("FFM_KERNEL", "ped")
    elif variant == "cl4smem":
        monkeypatch.setenv("FFM_CLUSTER", "4")
        monkeypatch.setenv("FFM_FIELDS_SMEM", "1")
    elif variant.startswith("cl"):
        monkeypatch.setenv("FFM_CLUSTER", variant[2:])
    m = rooms_map_c3()
    sff = generate_sff(m, "bfs8", np.float32)
    assert np.array_equal(sff, c_oracle.geodesic(m, "bfs8"))
    N, seed = 10000, 0x5EED0003
    params = {"k_S": 3, "k_D": 1, "neighborhood": "moore"}
    pos = place(m, N, B, 0, seed)
    n = np.full((B,), N, np.int32)
    ref = c_oracle.run_core_batch(m, sff, pos, n, params, seed=seed, max_steps=T, threads=min(B, 8), traj_steps=T, want_state=True,
                                  guard=1e-5, record_moves=T)
    sim = BatchSim(m, sff, B, N, params, seed=seed)
    info = sim.kernel_info()
    if variant == "ped":
        assert info["name"] == "ffm_core_rollout_kernel" and not info["fields_in_smem"]
    else:
        assert info["name"] == "ffm_cell_rollout_kernel"
        assert info["cluster"] == {"auto": 2, "cl4smem": 4, "cl8": 8}[variant], info
        assert info["fields_in_smem"] == (variant == "cl4smem"), info
    sim.set_positions(pos, n)
    cells, cnt = sim.rollout(T, draws=dict(move=torch.from_numpy(ref["move_draws"]).cuda()), record=T)
    torch.cuda.synchronize()
    steps, ped = sim.counters()
    assert np.array_equal(ped, ref["ped_steps"]) and (steps == T).all()
    cells, cnt = cells.cpu().numpy(), cnt.cpu().numpy()
    assert np.array_equal(cnt, ref["traj_n"])
    for e in range(B):
        for t in (0, 1, 50, T // 2, T - 1):
            k = cnt[e, t]
            assert np.array_equal(cells[e, t, :k], ref["traj"][e, t, :k]), (e, t)
    p_gpu, n_gpu = sim.get_positions()
    assert np.array_equal(n_gpu, ref["final_n"])
    W = 256
    for e in range(B):
        assert np.array_equal(p_gpu[e, :n_gpu[e], 0] * W + p_gpu[e, :n_gpu[e], 1], ref["final_pos"][e, :n_gpu[e]])
    assert np.array_equal(sim.get_dff().view(np.uint32), ref["final_dff"].view(np.uint32))


def test_c3_all_kernel_variants_agree_under_keyed_draws(cuda_device, monkeypatch):
    """C3 geometry, pure Philox mode (no recorded draws), 8 episodes x 400 steps: the pedestrian-centric kernel, the 2-CTA cluster
    with the fields in L2 (twice: run-to-run determinism), the 4-CTA cluster with the fields in distributed shared memory and the
    8-CTA cluster compute the same function -- positions, counts, pedestrian-steps and DFF bits are identical.  (Races in the
    cross-CTA paths would show up here as run-to-run or variant-to-variant differences.)"""
    from ffm_b200 import BatchSim
    from ffm_b200.sff import generate_sff
    from ffm_b200.workloads import place, rooms_map_c3
    m = rooms_map_c3()
    sff = generate_sff(m, "bfs8", np.float32)
    B, N, seed, T = 8, 10000, 0xC0FFEE, 400
    params = {"k_S": 3, "k_D": 1, "neighborhood": "moore"}
    pos = place(m, N, B, 0, seed)
    n = np.full((B,), N, np.int32)
    results = {}
    for tag, env in (("ped", {"FFM_KERNEL": "ped"}), ("cl2", {}), ("cl2_again", {}), ("cl4smem", {"FFM_CLUSTER": "4", "FFM_FIELDS_SMEM": "1"}),
                     ("cl8", {"FFM_CLUSTER": "8"})):
        for k in ("FFM_KERNEL", "FFM_CLUSTER", "FFM_FIELDS_SMEM"):
            monkeypatch.delenv(k, raising=False)
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        sim = BatchSim(m, sff, B, N, params, seed=seed)
        sim.set_positions(pos, n)
        sim.rollout(T // 2)
        sim.rollout(T - T // 2)
        p, k = sim.get_positions()
        results[tag] = (sim.kernel_info()["cluster"], p, k, sim.counters()[1].copy(), sim.get_dff().view(np.uint32).copy())
        sim.close()
    assert [results[t][0] for t in ("ped", "cl2", "cl4smem", "cl8")] == [1, 2, 4, 8]
    ref = results["ped"]
    for tag, got in results.items():
        assert np.array_equal(got[2], ref[2]) and np.array_equal(got[1], ref[1]), tag
        assert np.array_equal(got[3], ref[3]) and np.array_equal(got[4], ref[4]), tag


@pytest.mark.parametrize("k_D,track", [(0, False), (1, True)])
def test_c2_kernel_variants_agree_to_evacuation(cuda_device, monkeypatch, k_D, track):
    """C2 geometry to full evacuation (~1700 steps, ~4.5e5 draws per episode), 96 episodes, pure Philox mode: the cell-centric
    kernel (one CTA), the same as 2- and 4-CTA clusters, and the pedestrian-centric kernel give identical evacuation times,
    pedestrian-step counts and final DFF bits."""
    import bench
    from ffm_b200 import BatchSim
    m, sff = _c2()
    B, N, seed = 96, 1024, 0xD1CE
    params = {"k_S": 3, "k_D": k_D, "neighborhood": "moore"}
    pos = bench.place(m, N, B, 500, seed)
    n = np.full((B,), N, np.int32)
    out = {}
    for tag, env in (("cell", {"FFM_KERNEL": "cell"}), ("ped", {"FFM_KERNEL": "ped"}), ("cl2", {"FFM_KERNEL": "cell", "FFM_CLUSTER": "2", "FFM_FIELDS_SMEM": "1"}),
                     ("cl4g", {"FFM_KERNEL": "cell", "FFM_CLUSTER": "4"})):
        for k in ("FFM_KERNEL", "FFM_CLUSTER", "FFM_FIELDS_SMEM"):
            monkeypatch.delenv(k, raising=False)
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        sim = BatchSim(m, sff, B, N, params, seed=seed, episode_base=500, track_dff=track)
        sim.set_positions(pos, n)
        sim.rollout(4096)
        steps, ped = sim.counters()
        assert (sim.get_positions()[1] == 0).all()
        out[tag] = (steps.copy(), ped.copy(), sim.get_dff().view(np.uint32).copy() if track else None)
        sim.close()
    for tag in ("ped", "cl2", "cl4g"):
        assert np.array_equal(out[tag][0], out["cell"][0]) and np.array_equal(out[tag][1], out["cell"][1]), tag
        if track:
            assert np.array_equal(out[tag][2], out["cell"][2]), tag
