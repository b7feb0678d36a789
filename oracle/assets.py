"""Synthetic maps / static floor fields for the BASELINE configs (test infrastructure).

Restates the reference's asset generators:
  * walled room with one exit in the middle of the top wall  -- Create_Map.py:9-19,
    create_12x12_map_and_sff.py:15-25
  * obstacle-blind min-over-exits norm SFF, inf on non-walkable cells -- Create_SFF.py:14-33
    (float64, L1 / L2=np.hypot / Linf), create_12x12_map_and_sff.py:36-50 (float32, L1)
plus the C3 / C5 synthetic floor plans SURVEY.md section 8(d) specifies.
"""
import numpy as np


def room_map(h, w, dtype=np.uint8):
    m = np.zeros((h, w), dtype=dtype)
    m[0, :] = 2
    m[-1, :] = 2
    m[:, 0] = 2
    m[:, -1] = 2
    m[0, w // 2] = 3
    return m


def sff_norm_min(map_array, metric, dtype=np.float64):
    """Create_SFF.py:14-33 literally (triple loop); metric in {"L1","L2","Linf"}."""
    exits = np.argwhere(map_array == 3)
    h, w = map_array.shape
    out = np.full((h, w), np.inf, dtype=dtype)
    for i in range(h):
        for j in range(w):
            if (map_array[i, j] == 0) | (map_array[i, j] == 3):
                for ex, ey in exits:
                    if metric == "L1":
                        d = abs(i - ex) + abs(j - ey)
                    elif metric == "L2":
                        # Create_SFF.py:26 calls np.hypot.  The shipped data/sff/distance_L2.npy equals
                        # the CORRECTLY ROUNDED hypot on every cell (== IEEE sqrt of the exact integer
                        # dx^2+dy^2, == math.hypot); this container's np.hypot is 1 ulp off on 2 of
                        # 2500 cells, so the shipped file -- the reference's own golden vector -- wins.
                        d = np.sqrt(np.float64(int(i - ex) ** 2 + int(j - ey) ** 2))
                    else:
                        d = max(abs(i - ex), abs(j - ey))
                    out[i, j] = min(out[i, j], d)
    return out


def rooms_map_c3(h=256, w=256, seed=0x5EED0003):
    """C3 floor plan: border walls, exits (0,w/4),(0,3w/4),(h-1,w/2),(h/2,0),
    interior walls forming a 3x3 grid of rooms, each wall segment with a 4-cell door gap whose
    offset is drawn from a seeded generator."""
    rng = np.random.RandomState(seed & 0x7FFFFFFF)
    m = np.zeros((h, w), dtype=np.uint8)
    m[0, :] = 2; m[-1, :] = 2; m[:, 0] = 2; m[:, -1] = 2
    rows = [h // 3, 2 * h // 3]
    cols = [w // 3, 2 * w // 3]
    for r in rows:
        m[r, 1:-1] = 2
    for c in cols:
        m[1:-1, c] = 2
    rb = [0] + rows + [h - 1]
    cb = [0] + cols + [w - 1]
    for r in rows:                      # doors in horizontal walls, one per room column
        for k in range(3):
            lo, hi = cb[k] + 2, cb[k + 1] - 6
            s = int(rng.randint(lo, hi))
            m[r, s:s + 4] = 0
    for c in cols:                      # doors in vertical walls, one per room row
        for k in range(3):
            lo, hi = rb[k] + 2, rb[k + 1] - 6
            s = int(rng.randint(lo, hi))
            m[s:s + 4, c] = 0
    for (er, ec) in [(0, w // 4), (0, 3 * w // 4), (h - 1, w // 2), (h // 2, 0)]:
        m[er, ec] = 3
    return m


def obstacle_map_c5(h=1024, w=1024, index=0, fill=0.20, n_exits=8, seed=0x5EED0005):
    """C5 map: border walls, ~``fill`` of the area covered by random axis-aligned
    rectangles, ``n_exits`` exits spread over the border."""
    rng = np.random.RandomState((seed + index) & 0x7FFFFFFF)
    m = np.zeros((h, w), dtype=np.uint8)
    target = fill * h * w
    covered = 0
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
        m[0, int(f * w)] = 3
        m[h - 1, int(f * w)] = 3
        m[int(f * h), 0] = 3
        m[int(f * h), w - 1] = 3
    # keep the cell just inside every exit walkable
    for r, c in np.argwhere(m == 3):
        rr = min(max(r, 1), h - 2); cc = min(max(c, 1), w - 2)
        m[rr, cc] = 0
    return m


def sff_norm_min_fast(map_array, metric, dtype=np.float64):
    """Vectorised equivalent of sff_norm_min (same values; the triple loop is too slow beyond ~100x100)."""
    exits = np.argwhere(map_array == 3)
    h, w = map_array.shape
    rr, cc = np.meshgrid(np.arange(h), np.arange(w), indexing="ij")
    best = np.full((h, w), np.inf, dtype=np.float64)
    for ex, ey in exits:
        dx, dy = np.abs(rr - ex), np.abs(cc - ey)
        d = (dx + dy) if metric == "L1" else (np.maximum(dx, dy) if metric == "Linf" else np.sqrt((dx * dx + dy * dy).astype(np.float64)))
        best = np.minimum(best, d)
    out = np.full((h, w), np.inf, dtype=dtype)
    walk = (map_array == 0) | (map_array == 3)
    out[walk] = best[walk].astype(dtype)
    return out
