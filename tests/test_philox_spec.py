"""Known-answer tests of the noise-stream specification (oracle/philox.py)."""
import numpy as np

from oracle import philox


def _kat(ctr, key):
    return [int(v) for v in philox.philox4x32_10(*[np.uint32(c) for c in ctr], *key)]


def test_random123_known_answers():
    # Random123 kat_vectors, philox4x32 10 rounds
    assert _kat((0, 0, 0, 0), (0, 0)) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert _kat((0xffffffff,) * 4, (0xffffffff, 0xffffffff)) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert _kat((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0)) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_normal_rows_are_standard_normal_and_shard_invariant():
    z = philox.normal_rows(99, 0, 2048, 17)
    assert z.dtype == np.float32 and z.shape == (2048, 256)
    assert abs(float(z.mean())) < 5e-3 and abs(float(z.std()) - 1) < 5e-3
    assert np.isfinite(z).all()
    # rows are keyed by the global sample index: any sharding reproduces the same rows
    a = philox.normal_rows(99, 1000, 8, 17)
    b = np.concatenate([philox.normal_rows(99, 1000, 3, 17), philox.normal_rows(99, 1003, 5, 17)])
    assert np.array_equal(a, b)
    assert not np.array_equal(philox.normal_rows(99, 0, 4, 17), philox.normal_rows(99, 0, 4, 18))
    assert not np.array_equal(philox.normal_rows(99, 0, 4, 17), philox.normal_rows(100, 0, 4, 17))
    # 64-bit sample indices reach the fourth counter word
    assert not np.array_equal(philox.normal_rows(1, 5, 1, 0), philox.normal_rows(1, 5 + (1 << 32), 1, 0))
