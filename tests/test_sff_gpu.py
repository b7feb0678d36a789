"""GPU SFF generation (ffm_sff_generate) against the reference's shipped fields and the oracles."""
import hashlib
import json
import os

import numpy as np
import pytest

from helpers import GOLDEN
from oracle import assets, c_oracle

pytestmark = pytest.mark.gpu


def test_shipped_distance_files_bit_exact(cuda_device):
    """Regenerates data/sff/distance_{L1,L2,Linf}.npy (float64, 195 inf) from the 50x50 room."""
    from ffm_b200.sff import generate_sff
    with open(os.path.join(GOLDEN, "shipped_assets.json")) as f:
        want = json.load(f)
    m = assets.room_map(50, 50)
    for metric in ("L1", "L2", "Linf"):
        got = generate_sff(m, metric, np.float64)
        meta = want[f"data/sff/distance_{metric}.npy"]
        assert got.dtype == np.float64 and int(np.isinf(got).sum()) == meta["n_inf"] == 195
        assert hashlib.sha256(got.tobytes()).hexdigest() == meta["sha256"], metric


def test_12x12_l1_float32(cuda_device):
    from ffm_b200.sff import generate_sff
    m = assets.room_map(12, 12)
    got = generate_sff(m, "L1", np.float32)
    assert got.dtype == np.float32 and np.array_equal(got, assets.sff_norm_min(m, "L1", np.float32))


@pytest.mark.parametrize("metric", ["L1", "L2", "Linf"])
def test_norm_min_multi_exit_batch(cuda_device, metric):
    from ffm_b200.sff import generate_sff
    maps = np.stack([assets.obstacle_map_c5(70, 90, index=i, n_exits=8) for i in range(3)])
    got = generate_sff(maps, metric, np.float64)
    for i in range(3):
        assert np.array_equal(got[i], assets.sff_norm_min_fast(maps[i], metric, np.float64)), i


@pytest.mark.parametrize("mode", ["bfs4", "bfs8", "dijkstra8"])
@pytest.mark.parametrize("shape", [(33, 47), (200, 160)])
def test_geodesic_bit_exact(cuda_device, mode, shape):
    from ffm_b200.sff import generate_sff
    maps = np.stack([assets.obstacle_map_c5(shape[0], shape[1], index=i, n_exits=4) for i in range(2)])
    maps[1, 5:9, 5:9] = 2
    maps[1, 6:8, 6:8] = 0           # enclosed pocket: unreachable -> inf
    got, rounds = generate_sff(maps, mode, np.float32, return_rounds=True)
    assert rounds >= 1
    for i in range(2):
        want = c_oracle.geodesic(maps[i], mode)
        assert np.array_equal(got[i].view(np.uint32), want.view(np.uint32)), (mode, i)
    assert np.isinf(got[1, 6, 6])


def test_geodesic_c3_and_c5_maps(cuda_device):
    """The C3 floor plan (256x256 rooms with doors) and one C5 map (1024x1024, 20% obstacles)."""
    import torch
    from ffm_b200.sff import generate_sff
    m3 = assets.rooms_map_c3()
    got = generate_sff(m3, "bfs8", np.float32)
    assert np.array_equal(got, c_oracle.geodesic(m3, "bfs8"))
    assert np.isfinite(got[(m3 == 0)]).all()
    m5 = assets.obstacle_map_c5(1024, 1024, index=0)
    for mode in ("bfs4", "dijkstra8"):
        got = generate_sff(torch.from_numpy(m5).cuda(), mode, np.float32).cpu().numpy()
        assert np.array_equal(got.view(np.uint32), c_oracle.geodesic(m5, mode).view(np.uint32)), mode


@pytest.mark.parametrize("mode", ["bfs8", "dijkstra8"])
def test_geodesic_work_queue_is_order_independent(cuda_device, mode):
    """The tile work queue relaxes in whatever order the persistent CTAs happen to pop; the fixpoint must not depend on it:
    the same batch, many times over, always equals the heap-Dijkstra oracle bit for bit (odd sizes: ragged edge tiles)."""
    import torch
    from ffm_b200.sff import generate_sff
    maps = np.stack([assets.obstacle_map_c5(257, 301, index=40 + i, fill=0.3, n_exits=8) for i in range(12)])
    want = np.stack([c_oracle.geodesic(mm, mode) for mm in maps])
    dm = torch.from_numpy(maps).cuda()
    for rep in range(12):
        got = generate_sff(dm, mode, np.float32).cpu().numpy()
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), rep


def _serpentine(h, w, gap_every=2, thick=1):
    """Walls on every other row with the gap alternating between the left and the right end: the only path is a serpentine,
    so in-tile path lengths run into the hundreds (the warp kernel's bit planes rebase after 256 levels) and the two sides of
    a tile can be thousands of steps apart (level jumps)."""
    m = np.zeros((h, w), np.uint8)
    m[0, :] = 2; m[-1, :] = 2; m[:, 0] = 2; m[:, -1] = 2
    left = True
    for r in range(2, h - 2, gap_every):
        m[r:r + thick, 1:w - 1] = 2
        if left:
            m[r:r + thick, 1] = 0
        else:
            m[r:r + thick, w - 2] = 0
        left = not left
    m[0, 1] = 3
    m[1, 1] = 0
    return m


@pytest.mark.parametrize("mode", ["bfs4", "bfs8", "dijkstra8"])
def test_geodesic_serpentine_mazes(cuda_device, mode):
    """Long in-tile paths, huge halo spreads, ragged edge tiles, exits in tile corners: BFS levels bit for bit."""
    from ffm_b200.sff import generate_sff
    for shape, every in (((64, 64), 2), ((97, 131), 2), ((130, 70), 3)):
        m = _serpentine(shape[0], shape[1], gap_every=every)
        got = generate_sff(m, mode, np.float32)
        want = c_oracle.geodesic(m, mode)
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), (mode, shape)
        assert np.nanmax(np.where(np.isfinite(want), want, 0)) > 1000          # the path really is long
    # a spiral-free but dense case: random 45 % obstacles, many exits (most cells unreachable pockets)
    rng = np.random.RandomState(3)
    m = (rng.rand(150, 190) < 0.45).astype(np.uint8) * 2
    m[0, :] = 2; m[-1, :] = 2; m[:, 0] = 2; m[:, -1] = 2
    for k in range(6):
        m[0, 10 + 30 * k] = 3
        m[1, 10 + 30 * k] = 0
    got = generate_sff(m, mode, np.float32)
    assert np.array_equal(got.view(np.uint32), c_oracle.geodesic(m, mode).view(np.uint32))


def test_geodesic_unit_cost_kernels_agree(cuda_device, monkeypatch):
    """The warp-per-tile kernel (default for BFS-4 / BFS-8) and the CTA-per-tile kernel (FFM_SFF_KERNEL=tile) give the same field."""
    from ffm_b200.sff import generate_sff
    maps = np.stack([assets.obstacle_map_c5(300, 260, index=70 + i, fill=0.35, n_exits=4) for i in range(6)])
    for mode in ("bfs4", "bfs8"):
        monkeypatch.delenv("FFM_SFF_KERNEL", raising=False)
        a = generate_sff(maps, mode, np.float32)
        monkeypatch.setenv("FFM_SFF_KERNEL", "tile")
        b = generate_sff(maps, mode, np.float32)
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), mode
