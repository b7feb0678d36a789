"""Synthetic maps, fields and placements of the BASELINE configurations (product side; the parity
tests build the same inputs independently on the checker side).

  C1/C2  walled room, one exit in the middle of the top wall (layout of Create_Map.py:9-19)
  C3     256x256 floor plan: border walls, 4 exits, 3x3 rooms separated by walls with 4-cell doors
  C5     1024x1024 maps, ~20 % area of random rectangular obstacles, 8 exits on the border
"""
import numpy as np

from . import philox


def room_map(h, w):
    m = np.zeros((h, w), dtype=np.uint8)
    m[0, :] = 2; m[-1, :] = 2; m[:, 0] = 2; m[:, -1] = 2
    m[0, w // 2] = 3
    return m


def sff_room(m, nbh):
    """Obstacle-blind distance to the exit: Linf for Moore, L1 for von Neumann (Create_SFF.py:24,28);
    float32, inf on non-walkable cells."""
    h, w = m.shape
    er, ec = np.argwhere(m == 3)[0]
    rr, cc = np.meshgrid(np.arange(h), np.arange(w), indexing="ij")
    d = np.maximum(abs(rr - er), abs(cc - ec)) if nbh == "moore" else abs(rr - er) + abs(cc - ec)
    out = np.full((h, w), np.inf, dtype=np.float32)
    walk = (m == 0) | (m == 3)
    out[walk] = d[walk]
    return out


def rooms_map_c3(h=256, w=256, seed=0x5EED0003):
    rng = np.random.RandomState(seed & 0x7FFFFFFF)
    m = np.zeros((h, w), dtype=np.uint8)
    m[0, :] = 2; m[-1, :] = 2; m[:, 0] = 2; m[:, -1] = 2
    rows, cols = [h // 3, 2 * h // 3], [w // 3, 2 * w // 3]
    for r in rows:
        m[r, 1:-1] = 2
    for c in cols:
        m[1:-1, c] = 2
    rb, cb = [0] + rows + [h - 1], [0] + cols + [w - 1]
    for r in rows:
        for k in range(3):
            s = int(rng.randint(cb[k] + 2, cb[k + 1] - 6))
            m[r, s:s + 4] = 0
    for c in cols:
        for k in range(3):
            s = int(rng.randint(rb[k] + 2, rb[k + 1] - 6))
            m[s:s + 4, c] = 0
    for (er, ec) in [(0, w // 4), (0, 3 * w // 4), (h - 1, w // 2), (h // 2, 0)]:
        m[er, ec] = 3
    return m


def obstacle_map_c5(h=1024, w=1024, index=0, fill=0.20, n_exits=8, seed=0x5EED0005):
    rng = np.random.RandomState((seed + index) & 0x7FFFFFFF)
    m = np.zeros((h, w), dtype=np.uint8)
    target, covered = fill * h * w, 0
    while covered < target:
        rh, rw = int(rng.randint(4, max(5, h // 16))), int(rng.randint(4, max(5, w // 16)))
        r0, c0 = int(rng.randint(2, h - rh - 2)), int(rng.randint(2, w - rw - 2))
        blk = m[r0:r0 + rh, c0:c0 + rw]
        covered += int((blk == 0).sum())
        blk[...] = 2
    m[0, :] = 2; m[-1, :] = 2; m[:, 0] = 2; m[:, -1] = 2
    per_side = max(1, n_exits // 4)
    for k in range(per_side):
        f = (k + 1) / (per_side + 1)
        m[0, int(f * w)] = 3; m[h - 1, int(f * w)] = 3
        m[int(f * h), 0] = 3; m[int(f * h), w - 1] = 3
    for r, c in np.argwhere(m == 3):
        m[min(max(r, 1), h - 2), min(max(c, 1), w - 2)] = 0
    return m


def place(m, n, episodes, episode_base, seed):
    """Uniform placement without replacement on free cells, keyed by GLOBAL episode id (Philox stream PLACE):
    the n free cells with the smallest keys, in key order (the batched stand-in for initialize_agents(),
    ffm_core.py:23-26).  int32 [episodes, n, 2]."""
    free = np.argwhere(m == 0).astype(np.int32)
    out = np.empty((episodes, n, 2), dtype=np.int32)
    ords = np.arange(len(free))
    chunk = max(1, (1 << 22) // max(len(free), 1))
    for e0 in range(0, episodes, chunk):
        e1 = min(episodes, e0 + chunk)
        eps = (episode_base + np.arange(e0, e1))[:, None]
        keys, _ = philox.draw2(seed, eps, 0, philox.STREAM_PLACE, ords[None, :])
        sel = np.argsort(keys, axis=1, kind="stable")[:, :n]
        out[e0:e1] = free[sel]
    return out
