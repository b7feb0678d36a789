"""Counter-based uniform draws shared by the oracle and the CUDA path (test infrastructure).

Philox4x32-10 (Salmon et al., "Parallel random numbers: as easy as 1, 2, 3", SC'11),
restated from the published algorithm.  The reference itself draws from the process-global
NumPy / ``random`` generators (``model/ffm_core.py:84,95,96``); to make both sides
deterministic *and* order-independent every draw is re-keyed as

    counter = (entity, step, episode, stream)      key = (seed_lo, seed_hi)

and the four output words give two 53-bit doubles built exactly like NumPy's
``random_sample`` (``(a >> 5) * 2**26 + (b >> 6)) / 2**53``):

    u0 = f(out[0], out[1])      u1 = f(out[2], out[3])

Streams:
    STREAM_MOVE     entity = agent index (alive rank)   u0 -> np.random.choice(n, p=probs)
    STREAM_CONFLICT entity = target cell (row * W + col) u0 -> coin np.random.rand() < 0.5
                                                         u1 -> winner agents[int(u1 * k)]
    STREAM_EPS      entity = agent index                 u0 -> random.random() < epsilon
                                                         u1 -> valid[int(u1 * n_valid)]
    STREAM_PLACE    entity = free-cell ordinal, step = 0 u0 -> placement sort key
"""
import numpy as np

STREAM_MOVE = 0
STREAM_CONFLICT = 1
STREAM_EPS = 2
STREAM_PLACE = 3

_M0 = np.uint64(0xD2511F53)
_M1 = np.uint64(0xCD9E8D57)
_W0 = 0x9E3779B9
_W1 = 0xBB67AE85
_MASK = np.uint64(0xFFFFFFFF)
_S32 = np.uint64(32)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised Philox4x32-10.  All inputs broadcastable uint32-valued arrays; returns 4 uint32 arrays."""
    c0, c1, c2, c3 = (np.asarray(c, dtype=np.uint64) & _MASK for c in (c0, c1, c2, c3))
    c0, c1, c2, c3 = np.broadcast_arrays(c0, c1, c2, c3)
    k0 = int(k0) & 0xFFFFFFFF
    k1 = int(k1) & 0xFFFFFFFF
    for _ in range(10):
        p0 = _M0 * c0
        p1 = _M1 * c2
        hi0, lo0 = p0 >> _S32, p0 & _MASK
        hi1, lo1 = p1 >> _S32, p1 & _MASK
        c0, c1, c2, c3 = (hi1 ^ c1 ^ np.uint64(k0)), lo1, (hi0 ^ c3 ^ np.uint64(k1)), lo0
        k0 = (k0 + _W0) & 0xFFFFFFFF
        k1 = (k1 + _W1) & 0xFFFFFFFF
    return tuple(c.astype(np.uint32) for c in (c0, c1, c2, c3))


def _to_double(a, b):
    a = np.asarray(a, dtype=np.uint64)
    b = np.asarray(b, dtype=np.uint64)
    return ((a >> np.uint64(5)) * np.uint64(67108864) + (b >> np.uint64(6))).astype(np.float64) * (1.0 / 9007199254740992.0)


def draw2(seed, episode, step, stream, entity):
    """Two doubles in [0,1) for (seed, episode, step, stream, entity); arrays broadcast."""
    seed = int(seed)
    o0, o1, o2, o3 = philox4x32_10(entity, step, episode, stream, seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    return _to_double(o0, o1), _to_double(o2, o3)


def draw2_scalar(seed, episode, step, stream, entity):
    u0, u1 = draw2(seed, episode, step, stream, entity)
    return float(u0), float(u1)
