"""Known-answer tests of the Philox4x32-10 restatement (Random123 kat_vectors)."""
import numpy as np

from oracle import philox

KATS = [
    ((0, 0, 0, 0), (0, 0), (0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8)),
    ((0xFFFFFFFF,) * 4, (0xFFFFFFFF,) * 2, (0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD)),
    ((0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344), (0xA4093822, 0x299F31D0),
     (0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1)),
]


def test_philox_kats():
    for ctr, key, want in KATS:
        got = philox.philox4x32_10(*ctr, *key)
        assert tuple(int(x) for x in got) == want


def test_draws_are_53_bit_uniforms_and_vectorise():
    ent = np.arange(1000)
    u0, u1 = philox.draw2(123456789012345, 7, 3, philox.STREAM_MOVE, ent)
    assert u0.dtype == np.float64 and (u0 >= 0).all() and (u0 < 1).all() and (u1 < 1).all()
    assert np.all(u0 * 2.0**53 == np.floor(u0 * 2.0**53))
    s0, s1 = philox.draw2_scalar(123456789012345, 7, 3, philox.STREAM_MOVE, 17)
    assert s0 == u0[17] and s1 == u1[17]
    assert abs(u0.mean() - 0.5) < 0.05 and abs(u1.mean() - 0.5) < 0.05


def test_product_side_philox_matches_oracle():
    """ffm_b200/philox.py (host-side placement keys) is an independent restatement; same numbers."""
    from ffm_b200 import philox as prod
    ent = np.arange(257)
    a = philox.draw2(0xDEADBEEFCAFE, 5, 0, philox.STREAM_PLACE, ent)
    b = prod.draw2(0xDEADBEEFCAFE, 5, 0, prod.STREAM_PLACE, ent)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
