"""The packed-integer OVER of the warp-per-tile fine kernel (raster.cu: blend_int) against the fp32 chain of the
reference's blend state as the oracle and blend_over evaluate it (src/vkvg_device_internal.c:203-209, UNORM8 store with
round-half-up): every (source channel S, source alpha A, destination channel D) with S <= A, which is what the kernel
sends down the integer path (solid colour, opacity 1, OVER, no channel above alpha).  CPU only, numpy."""
import numpy as np


def blend_float(S, A, D):
    f32 = np.float32
    lut = (np.arange(256, dtype=np.float32) / f32(255.0)).astype(f32)   # lut[i] == (float)i / 255.0f
    ia = (f32(1.0) - lut[A]).astype(f32)
    t = (lut[D] * ia).astype(f32)
    r = (lut[S] + t).astype(f32)
    q = ((r * f32(255.0)).astype(f32) + f32(0.5)).astype(f32)          # unorm8: cvt.rzi.sat.u8.f32(v * 255 + 0.5)
    return np.clip(np.trunc(q), 0, 255).astype(np.int64)


def blend_int(S, A, D):
    v = D * (255 - A) + 128
    return S + ((v + (v >> 8)) >> 8)


def test_integer_blend_equals_fp32_blend_for_every_premultiplied_case():
    D = np.arange(256, dtype=np.int64)[None, None, :]
    S = np.arange(256, dtype=np.int64)[None, :, None]
    for a0 in range(0, 256, 32):
        A = np.arange(a0, a0 + 32, dtype=np.int64)[:, None, None]
        f, i = blend_float(S, A, D), blend_int(S, A, D)
        ok = (S <= A) & (D <= 255)
        assert np.array_equal(f[np.broadcast_to(ok, f.shape)], i[np.broadcast_to(ok, i.shape)])
        assert int(i[np.broadcast_to(ok, i.shape)].max()) <= 255          # no saturation needed when S <= A


def test_packed_form_matches_scalar_form():
    rng = np.random.default_rng(7)
    dst = rng.integers(0, 1 << 32, 200000, dtype=np.uint64)
    A = int(rng.integers(0, 256))
    col = [int(rng.integers(0, A + 1)) for _ in range(3)] + [A]
    s_lo, s_hi, IA = col[0] | col[2] << 16, col[1] | col[3] << 16, 255 - A
    lo, hi = dst & 0x00FF00FF, (dst >> 8) & 0x00FF00FF
    x, y = lo * IA + 0x00800080, hi * IA + 0x00800080
    assert int(x.max()) < 1 << 32 and int(y.max()) < 1 << 32
    x, y = x + ((x >> 8) & 0x00FF00FF), y + ((y >> 8) & 0x00FF00FF)
    x, y = ((x >> 8) & 0x00FF00FF) + s_lo, ((y >> 8) & 0x00FF00FF) + s_hi
    out = (x & 0xFF) | ((y & 0xFF) << 8) | (((x >> 16) & 0xFF) << 16) | (((y >> 16) & 0xFF) << 24)
    for k in range(4):
        d = ((dst >> (8 * k)) & 0xFF).astype(np.int64)
        assert np.array_equal(((out >> (8 * k)) & 0xFF).astype(np.int64), blend_int(col[k], A, d))
