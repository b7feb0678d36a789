"""Host-side Philox4x32-10 (NumPy), same keying as csrc/ffm_device.cuh.

Used by the host layer for batched initial placement keys; the in-kernel draws never come from
here.  counter = (entity, step, episode, stream), key = seed; two 53-bit doubles per call.
"""
import numpy as np

STREAM_MOVE, STREAM_CONFLICT, STREAM_EPS, STREAM_PLACE = 0, 1, 2, 3


def _rounds(c, k0, k1):
    m0, m1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
    lo32, sh = np.uint64(0xFFFFFFFF), np.uint64(32)
    c0, c1, c2, c3 = c
    for r in range(10):
        a, b = m0 * c0, m1 * c2
        ka = np.uint64((k0 + r * 0x9E3779B9) & 0xFFFFFFFF)
        kb = np.uint64((k1 + r * 0xBB67AE85) & 0xFFFFFFFF)
        c0, c1, c2, c3 = (b >> sh) ^ c1 ^ ka, b & lo32, (a >> sh) ^ c3 ^ kb, a & lo32
    return c0, c1, c2, c3


def draw2(seed, episode, step, stream, entity):
    seed = int(seed)
    ctr = np.broadcast_arrays(*(np.asarray(x, dtype=np.uint64) & np.uint64(0xFFFFFFFF)
                                for x in (entity, step, episode, stream)))
    o = _rounds(ctr, seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    scale = 1.0 / 9007199254740992.0

    def dbl(a, b):
        return ((a >> np.uint64(5)) * np.uint64(1 << 26) + (b >> np.uint64(6))).astype(np.float64) * scale

    return dbl(o[0], o[1]), dbl(o[2], o[3])
