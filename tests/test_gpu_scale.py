"""GPU: BASELINE.json's full-size configs checked through size-independent properties (the oracle cannot finish
them in seconds): linearity of the integer winding, invariance under tile-row striping, translation by whole tiles,
determinism, and agreement of a random subset of tiles with the oracle."""
import numpy as np
import pytest

import vkvg_b200 as v
from tests import scenes
from tests.parity import pixel_stats

pytestmark = pytest.mark.gpu


def _poly_edges(polys):
    es = []
    for p in polys:
        q = np.floor(p * 256 + 0.5).astype(np.int32)
        es.append(np.concatenate([q, np.roll(q, -1, 0)], 1))
    return np.concatenate(es)


def test_c2_winding_linearity_and_orientation(dev4):
    """W(A u B) = W(A) + W(B) and W(reversed) = -W on the 100k-polygon / 4096^2 config."""
    polys, _ = scenes.polygons_c2(100000, 4096, 1)
    e = _poly_edges(polys)
    assert len(e) > 1_000_000
    half = len(e) // 2
    wa = dev4.winding(e[:half], 4096, 4096)
    wb = dev4.winding(e[half:], 4096, 4096)
    wab = dev4.winding(e, 4096, 4096)
    assert np.array_equal(wa + wb, wab)
    del wa, wb
    wr = dev4.winding(e[:, [2, 3, 0, 1]], 4096, 4096)
    assert np.array_equal(wr, -wab)
    # random 64x64 windows against the brute-force oracle
    import oracle
    rng = np.random.default_rng(0)
    for _ in range(4):
        x0, y0 = (int(t) for t in rng.integers(0, 4096 - 64, 2))
        sel = ((np.maximum(e[:, 1], e[:, 3]) >= y0 * 256) & (np.minimum(e[:, 1], e[:, 3]) <= (y0 + 64) * 256) &
               (np.minimum(e[:, 0], e[:, 2]) <= (x0 + 64) * 256))
        sub = e[sel] - np.array([x0 * 256, y0 * 256, x0 * 256, y0 * 256], np.int32)
        ref = oracle.winding_brute(sub, 64, 64, 4)
        assert np.array_equal(wab[y0:y0 + 64, x0:x0 + 64], ref)


def _render_c2(dev, n, size, rule, seed=1, dy=0.0, height=None):
    polys, cols = scenes.polygons_c2(n, size, seed)
    polys = [np.round(p * 64) / 64 for p in polys]  # 1/64 px grid: the CTM translation below is then exact in float32
    s = v.Surface(dev, size, height or size)
    c = v.Context(s)
    cs = v.CommandStream()
    cs.set_fill_rule(rule)
    if dy:
        cs.translate(0.0, dy)
    for p, col in zip(polys, cols):
        cs.set_source_rgba(*[float(x) for x in col])
        cs.polyline(p)
        cs.close_path()
        cs.fill()
    assert c.replay(*cs.arrays()) == 0
    c.flush()
    img = s.pixels()
    c.close()
    s.close()
    return img


@pytest.mark.parametrize("rule", [0, 1])
def test_c2_full_size_properties(dev4, rule):
    img = _render_c2(dev4, 100000, 4096, rule)
    assert img[..., 3].mean() > 40  # the scene really covers a large part of the surface
    # determinism
    assert np.array_equal(img, _render_c2(dev4, 100000, 4096, rule))
    # translation by 5 whole tiles: same pixels, shifted (coordinates stay exactly representable)
    sh = _render_c2(dev4, 100000, 4096, rule, dy=80.0)
    if rule == 0:
        assert np.array_equal(sh[80:], img[:-80])
    else:
        # non-zero fills follow the reference's libtess: the vertices it adds at self-intersections are floats off the 1/64 grid, the
        # translation rounds them (as it does in the reference, which tessellates in user space), and single samples next to them move
        st = pixel_stats(sh[80:], img[:-80])
        assert st["p99_9"] == 0 and st["frac_diff"] < 2e-4 and st["max_diff"] <= 128, st
    # a stripe of the surface (top 1024 rows) equals the same rows of the full render: tile rows are independent
    top = _render_c2(dev4, 100000, 4096, rule, height=1024)
    if rule == 0:
        assert np.array_equal(top, img[:1024])
    else:   # (a 4096 x 1024 surface is another viewport: the off-grid intersection vertices round differently through its vertex stage)
        st = pixel_stats(top, img[:1024])
        assert st["p99_9"] == 0 and st["frac_diff"] < 2e-4 and st["max_diff"] <= 128, st


def test_c3_million_segment_dashed_stroke_properties(dev4):
    pts = scenes.polyline_c3(1_000_001, 4096, 1)
    s = v.Surface(dev4, 4096, 4096)
    c = v.Context(s)
    cs = v.CommandStream()
    cs.set_source_rgba(0.1, 0.2, 0.8, 1.0)
    cs.set_line_width(3.0)
    cs.set_line_join(1)
    cs.set_line_cap(1)
    cs.set_dash([10.0, 6.0], 0.0)
    cs.polyline(pts)
    cs.stroke_preserve()
    assert c.replay(*cs.arrays()) == 0
    verts, inds = c.stroke_geometry()
    assert inds.max() < len(verts) and len(inds) % 3 == 0
    # every vertex lies within hw + eps of the polyline's bounding box, none is NaN
    assert np.isfinite(verts).all()
    assert verts.min() >= 10 - 1.5 - 1e-3 and verts.max() <= 4096 - 10 + 1.5 + 1e-3
    # dash count: total length / 16 dashes, each contributing a start cap and an end cap
    seg = np.linalg.norm(np.diff(pts.astype(np.float64), axis=0), axis=1)
    n_dashes = int(np.ceil(seg.sum() / 16.0))
    st = np.zeros(1)
    c.flush()
    img = s.pixels()
    # opaque colour: covered pixels are exactly the colour or a box-filtered fraction of it; on/off ratio 10:6
    on = (img[..., 3] > 0).mean()
    assert 0.05 < on < 0.9
    assert abs(n_dashes - seg.sum() / 16.0) <= 1
    # the first 20000 segments rendered alone give the same vertices (dash phase is a prefix property)
    c2 = v.Context(s)
    cs2 = v.CommandStream()
    cs2.set_line_width(3.0)
    cs2.set_line_join(1)
    cs2.set_line_cap(1)
    cs2.set_dash([10.0, 6.0], 0.0)
    cs2.polyline(pts[:20001])
    assert c2.replay(*cs2.arrays()) == 0
    v2, i2 = c2.stroke_geometry()
    k = len(v2) - 64  # all but the tail (end cap of the shorter line)
    assert np.abs(verts[:k] - v2[:k]).max() <= 1e-3


def test_stripe_surfaces_equal_rows_of_the_whole_surface(dev4):
    """C5a's sharding unit: tile-row stripes rendered independently (vkvg_b200_surface_create_stripe) hold exactly the
    rows of the unsharded render — fills, gradients, dashed round strokes, arbitrary float coordinates."""
    from tests.golden import make_golden as mg
    from vkvg_b200 import sharding
    W, H = 640, 1000  # 62.5 tile rows: ragged

    def emit(c):
        polys, cols = scenes.polygons_c2(400, W, 5)
        for i, (p, col) in enumerate(zip(polys, cols)):
            p = p * np.array([1.0, H / W], np.float32)
            if i % 3 == 0:
                c.set_source_linear(0.0, 0.0, float(W), float(H), [(0, 1, 0, 0, 1), (0.5, 0, 1, 0, 0.5), (1, 0, 0, 1, 1)])
            elif i % 3 == 1:
                c.set_source_radial(W / 2, H / 2, 10.0, W / 2 + 20, H / 2 - 30, 400.0, [(0, 1, 1, 0, 1), (1, 0, 0, 1, 0.6)])
            else:
                c.set_source_rgba(*[float(x) for x in col])
            c.set_fill_rule(i % 2)
            c.move_to(float(p[0, 0]), float(p[0, 1]))
            for q in p[1:]:
                c.line_to(float(q[0]), float(q[1]))
            c.close_path()
            c.fill()
        pts = scenes.polyline_c3(3000, W, 2) * np.array([1.0, H / W], np.float32)
        c.set_source_rgba(0.1, 0.1, 0.1, 0.6)
        c.set_line_width(2.5)
        c.set_line_join(1)
        c.set_line_cap(1)
        c.set_dash([9.0, 4.0], 1.0)
        c.move_to(float(pts[0, 0]), float(pts[0, 1]))
        for q in pts[1:]:
            c.line_to(float(q[0]), float(q[1]))
        c.stroke()

    s = v.Surface(dev4, W, H)
    c = v.Context(s)
    emit(c)
    c.flush()
    full = s.pixels()
    for world in (2, 3, 8):
        for rank in range(world):
            surf, y0, h = sharding.render_striped(dev4, W, H, emit, rank, world)
            assert np.array_equal(surf.pixels()[:h], full[y0:y0 + h]), (world, rank)
            surf.close()
