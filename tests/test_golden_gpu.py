"""CUDA path against fixtures produced by the UNMODIFIED reference (tests/golden, oracle/make_golden.py):
bit-exact trajectories, DFF fields and step counts; replay of the stock seeded main.py run."""
import json
import os

import numpy as np
import pytest

from helpers import CORE_FIXTURES, GOLDEN, load_golden, pack_positions

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", CORE_FIXTURES)
def test_reference_fixture(cuda_device, core_kernel, name):
    import torch
    from ffm_b200 import BatchSim

    g = load_golden(name)
    N, T, W = len(g["pos0"]), int(g["steps"]), g["map"].shape[1]
    sim = BatchSim(g["map"], g["sff"], 1, N, g["params"], seed=int(g["seed"]), episode_base=int(g["episode"]),
                   track_dff=True)
    sim.set_positions(*pack_positions([g["pos0"].astype(np.int32)], N))
    every = int(g["dff_every"])
    t = 0
    for k, want_dff in enumerate(g["dff"]):          # DFF snapshots the fixture holds
        upto = k * every + 1
        cells, cnt = sim.rollout(upto - t, record=upto - t)
        torch.cuda.synchronize()
        cells, cnt = cells.cpu().numpy()[0], cnt.cpu().numpy()[0]
        for j in range(upto - t):
            want = g["traj_list"][t + j]
            assert cnt[j] == len(want)
            assert np.array_equal(cells[j, :cnt[j]], want[:, 0] * W + want[:, 1]), f"step {t + j}"
        assert np.array_equal(sim.get_dff()[0].view(np.uint32), want_dff.view(np.uint32)), f"DFF after step {upto}"
        t = upto
    sim.rollout(T + 5 - t)
    steps, _ = sim.counters()
    assert steps[0] == T and sim.get_positions()[1][0] == 0


def test_stock_main_seed42_replay(cuda_device, core_kernel):
    """The reference's stock run (main.py, config/default_config.yaml, seed 42: 272 steps,
    14 079 pedestrian-steps) replayed from its recorded MT19937 draws through the draw buffers."""
    import torch
    from ffm_b200 import BatchSim
    from oracle import assets

    z = np.load(os.path.join(GOLDEN, "stock_main_seed42.npz"))
    params = json.loads(str(z["params"]))
    m = assets.room_map(50, 50)
    sff = assets.sff_norm_min(m, "L1", np.float64)
    T, N = int(z["steps"]), len(z["pos0"])
    move = torch.from_numpy(np.nan_to_num(z["move"], nan=0.0)).reshape(1, T, N).cuda().contiguous()
    cf = np.zeros((1, T, 2500, 2))
    cf[0, z["conflict_t"], z["conflict_cell"]] = z["conflict_u"]
    sim = BatchSim(m, sff, 1, N, params, seed=0)
    sim.set_positions(*pack_positions([z["pos0"].astype(np.int32)], N))
    cells, cnt = sim.rollout(T + 10, draws=dict(move=move, conflict=torch.from_numpy(cf).cuda()), record=T)
    torch.cuda.synchronize()
    steps, ped_steps = sim.counters()
    assert steps[0] == 272 and ped_steps[0] == 14079
    cells, cnt = cells.cpu().numpy()[0], cnt.cpu().numpy()[0]
    offs = np.concatenate([[0], np.cumsum(z["traj_counts"])])
    for t in range(T):
        want = z["traj"][offs[t]:offs[t + 1]].astype(np.int64)
        assert np.array_equal(cells[t, :cnt[t]], want[:, 0] * 50 + want[:, 1]), f"step {t}"
    assert np.array_equal(sim.get_dff()[0].view(np.uint32), z["final_dff"].view(np.uint32))
