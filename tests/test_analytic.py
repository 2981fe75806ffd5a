"""Analytic-coverage mode (BASELINE.json north_star: "with an analytic-coverage mode alongside").

The reference has no such mode, so there is nothing of vkvg's to pin it against; what is checked:
  CPU  the oracle's definition (oracle/vkvg_oracle.c: ovk_area_brute) against closed forms — shoelace area,
       half-covered pixels, convergence of the 16-sample MSAA winding average towards it;
  GPU  the tile rasteriser's area integral (backdrop + V + H decomposition, float) against the oracle's edge-by-edge
       double-precision evaluation, and whole scenes' pixels within 1/255 (north_star's pixel tolerance).
"""
import numpy as np
import pytest

from tests import scenes
from tests.golden import make_golden as mg
from tests.test_gpu_parity import random_edges

AREA_TOL = 2e-4   # |A_gpu - A_oracle| per pixel for edges within a few hundred pixels of the tile (float vs double)


def shoelace(p):
    x, y = p[:, 0] / 256.0, p[:, 1] / 256.0
    return 0.5 * float(np.sum(x * np.roll(y, -1) - np.roll(x, -1) * y))


def poly_edges(p):
    p = np.asarray(p, np.int64)
    return np.concatenate([p, np.roll(p, -1, 0)], 1).astype(np.int32)


# ---------------------------------------------------------------------------------------------------------------
# CPU: the oracle's definition
# ---------------------------------------------------------------------------------------------------------------
def test_oracle_area_sums_to_polygon_area(oracle_lib):
    rng = np.random.default_rng(5)
    for _ in range(30):
        k = int(rng.integers(3, 9))
        p = rng.integers(2 * 256, 62 * 256, (k, 2))
        A = oracle_lib.area_brute(poly_edges(p), 64, 64)
        # W's sign convention: a polygon with positive shoelace area in y-down window space winds -1
        assert abs(A.sum() + shoelace(p)) < 1e-9 * max(1.0, abs(shoelace(p)))


def test_oracle_area_known_pixels(oracle_lib):
    # axis-aligned rectangle [1.5, 5.25] x [2, 4.5]: interior 1, edges fractional, outside 0
    p = np.array([[384, 512], [1344, 512], [1344, 1152], [384, 1152]])
    A = np.abs(oracle_lib.area_brute(poly_edges(p), 8, 8))
    assert A[3, 3] == 1.0 and A[2, 1] == 0.5 and A[2, 5] == 0.25 and A[4, 3] == 0.5 and A[4, 5] == 0.125
    assert A[0].sum() == 0 and A[:, 0].sum() == 0 and A[:, 6:].sum() == 0
    # a diagonal through a pixel: the triangle (0,0) (1,0) (0,1) covers half of pixel (0,0)
    t = np.array([[0, 0], [256, 0], [0, 256]])
    assert abs(abs(oracle_lib.area_brute(poly_edges(t), 4, 4)[0, 0]) - 0.5) < 1e-12


def test_oracle_area_is_the_limit_of_msaa(oracle_lib):
    polys, _ = scenes.polygons_c2(40, 128, 11)
    e = np.concatenate([poly_edges(np.round(p * 256)) for p in polys])
    A = oracle_lib.area_brute(e, 128, 128)
    w16 = oracle_lib.winding_brute(e, 128, 128, 16).mean(axis=2)
    assert np.abs(w16 - A).mean() < 0.02 and abs(w16.sum() - A.sum()) < 0.01 * np.abs(A).sum()


def test_oracle_analytic_scene_close_to_msaa16(oracle_lib):
    """whole-scene sanity: analytic pixels stay near the 16-sample MSAA pixels of the same calls."""
    a, m = oracle_lib.Oracle(96, 96, analytic=True), oracle_lib.Oracle(96, 96, 16)
    for g in (a, m):
        mg.pixel_scene(g, "nz_convex", 0, size=96)
    d = np.abs(a.pixels().astype(int) - m.pixels().astype(int))
    assert np.percentile(d, 99) <= 24 and d.mean() < 1.0


# ---------------------------------------------------------------------------------------------------------------
# GPU
# ---------------------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def deva():
    import vkvg_b200 as v
    d = v.Device(4, analytic=True)
    yield d
    d.close()


@pytest.mark.gpu
@pytest.mark.parametrize("w,h", [(64, 64), (100, 70), (17, 133)])
def test_area_vs_oracle(deva, oracle_lib, w, h):
    rng = np.random.default_rng(w)
    for kind, n, tol in (("short", 100, AREA_TOL), ("axis", 40, AREA_TOL), ("grid", 25, 2e-3), ("long", 12, 2e-3)):
        e = random_edges(rng, n, w, h, kind)
        got = deva.area(e, w, h).astype(np.float64)
        ref = oracle_lib.area_brute(e, w, h)
        assert np.abs(got - ref).max() < tol, (kind, float(np.abs(got - ref).max()))


@pytest.mark.gpu
def test_area_sums_to_polygon_area_full_size(deva):
    """size-independent property at 4096^2 (C2's surface): sum of A over the surface = -(signed area) of every polygon."""
    polys, _ = scenes.polygons_c2(20000, 4096, 2)
    ps = [np.round(np.clip(p, 1, 4095) * 256) for p in polys]
    e = np.concatenate([poly_edges(p) for p in ps])
    A = deva.area(e, 4096, 4096).astype(np.float64)
    want = -sum(shoelace(p) for p in ps)
    assert abs(A.sum() - want) < 2e-4 * sum(abs(shoelace(p)) for p in ps)


@pytest.mark.gpu
@pytest.mark.parametrize("name", mg.PIXEL_SCENES)
@pytest.mark.parametrize("seed", [0, 1, 2])
def test_pixels_vs_oracle(deva, oracle_lib, name, seed):
    import vkvg_b200 as v
    s = v.Surface(deva, 128, 128)
    c = v.Context(s)
    o = oracle_lib.Oracle(128, 128, analytic=True)
    for g in (c, o):
        mg.pixel_scene(g, name, seed)
    c.flush()
    d = np.abs(s.pixels().astype(int) - o.pixels().astype(int))
    assert np.percentile(d, 99.9) <= 1 and d.max() <= 2, (int(d.max()), float((d > 0).mean()))


@pytest.mark.gpu
def test_capture_and_multi_flush(deva, oracle_lib):
    """A of the last draw through the ordinary flush path; a second flush blends over the stored pixels."""
    import vkvg_b200 as v
    s = v.Surface(deva, 96, 80)
    c = v.Context(s)
    o = oracle_lib.Oracle(96, 80, analytic=True)
    for g in (c, o):
        g.set_source_rgba(0.1, 0.5, 0.9, 0.6)
        g.set_fill_rule(0)
        scenes.random_path(g, 77, size=80)
        g.fill()
    A = c.flush_capture_winding()
    assert np.abs(A.astype(np.float64) - o.last_area()).max() < AREA_TOL
    for g in (c, o):
        g.set_source_rgba(0.9, 0.2, 0.1, 0.5)
        g.set_line_width(5)
        g.move_to(5, 5)
        g.line_to(90, 70)
        g.line_to(10, 60)
        g.stroke()
    c.flush()
    d = np.abs(s.pixels().astype(int) - o.pixels().astype(int))
    assert d.max() <= 1


@pytest.mark.gpu
def test_analytic_c2_close_to_msaa16():
    """analytic and 16-sample MSAA renderings of C2-style polygons agree to within sampling error."""
    import vkvg_b200 as v
    polys, cols = scenes.polygons_c2(400, 512, 5)
    out = []
    for dev in (v.Device(4, analytic=True), v.Device(16)):
        s = v.Surface(dev, 512, 512)
        c = v.Context(s)
        for p, col in zip(polys, cols):
            c.set_source_rgba(*[float(x) for x in col])
            c.move_to(float(p[0, 0]), float(p[0, 1]))
            for q in p[1:]:
                c.line_to(float(q[0]), float(q[1]))
            c.close_path()
            c.fill()
        c.flush()
        out.append(s.pixels().astype(int))
        dev.close()
    d = np.abs(out[0] - out[1])
    assert d.mean() < 1.0 and np.percentile(d, 99) <= 24
