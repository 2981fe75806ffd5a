"""CPU: the vertex stage (vs_snap, vkvg_b200/csrc/pipeline.h) replaces `p * 2 / W` by `p * 2 * (1 / W)` when the surface size W is a power of
two, with 1 / W built from W's exponent bits (0x7F000000 - bits(W)).  Both are the correctly rounded value of the same real number, so they
must agree bit for bit for every float - checked here on random bit patterns (subnormals, infinities and NaN included) for every power of
two a surface can be; and the bit trick must give exactly 1 / W."""
import numpy as np


def test_reciprocal_of_a_power_of_two_from_its_exponent_bits():
    for k in range(0, 31):
        w = np.float32(2.0 ** k)
        inv = (np.uint32(0x7F000000) - w.view(np.uint32)).view(np.float32)
        assert inv == np.float32(2.0 ** -k)
        assert (w.view(np.uint32) & np.uint32(0x007FFFFF)) == 0          # what vs_snap tests: an all-zero mantissa
    for w in (np.float32(3.0), np.float32(1000.0), np.float32(4097.0), np.float32(1920.0)):
        assert (w.view(np.uint32) & np.uint32(0x007FFFFF)) != 0          # everything else keeps the division


def test_multiplying_by_the_reciprocal_equals_dividing_bit_for_bit():
    rng = np.random.default_rng(3)
    bits = rng.integers(0, 2 ** 32, 2_000_000, dtype=np.uint64).astype(np.uint32)
    coords = np.concatenate([bits.view(np.float32),                                             # every kind of float
                             (rng.random(1_000_000) * 40000 - 20000).astype(np.float32),       # device coordinates as scenes have them
                             np.array([0.0, -0.0, np.inf, -np.inf, 1e-45, -1e-45, 1.17549435e-38, 3.4028235e38], np.float32)])
    with np.errstate(all="ignore"):
        p2 = coords * np.float32(2.0)
        for k in (0, 1, 4, 9, 10, 12, 13, 14, 20, 30):
            w = np.float32(2.0 ** k)
            inv = (np.uint32(0x7F000000) - w.view(np.uint32)).view(np.float32)
            a = (p2 / w) - np.float32(1.0)
            b = (p2 * inv) - np.float32(1.0)
            same = (a.view(np.uint32) == b.view(np.uint32)) | (np.isnan(a) & np.isnan(b))
            assert same.all(), (k, coords[~same][:4])
