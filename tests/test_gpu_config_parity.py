"""GPU: parity AT THE BASELINE CONFIGS (BASELINE.json configs[1..4]), not at reduced sizes.

The checker is the reference's own object code where it applies (oracle/_ref: its tessellation - flattening, stroker with dashes,
libtess - compiled unmodified from /root/reference and shipped to the GPU box as a built library) with the oracle's scalar restatement
of the Vulkan rasteriser behind it, and the oracle's own restatement elsewhere.  Surfaces the scalar rasteriser cannot hold (8192^2,
16384^2 at 4 samples) are checked through WINDOWS of the logical surface (Oracle(window=...), tests/test_oracle_golden.py proves a
window equals the crop of the whole render).

Bars (BASELINE.json north_star): winding / coverage bit-exact, vertices within 1e-3 px, pixels within 1/255 at the 99.9th percentile.
"""
import numpy as np
import pytest

import vkvg_b200 as v
from tests import scenes
from tests.parity import align_vertices, pixel_stats

pytestmark = pytest.mark.gpu


def _c2_stream(polys, cols, rule):
    cs = v.CommandStream()
    cs.set_fill_rule(rule)
    for p, col in zip(polys, cols):
        cs.set_source_rgba(*[float(x) for x in col])
        cs.polyline(p)
        cs.close_path()
        cs.fill()
    return cs


def _emit_c2(g, polys, cols, rule):
    g.set_fill_rule(rule)
    for p, col in zip(polys, cols):
        g.set_source_rgba(*[float(x) for x in col])
        g.polyline(p)
        g.close_path()
        g.fill()


@pytest.mark.parametrize("rule", [1, 0])
def test_c2_whole_scene_pixels_vs_reference(dev4, oracle_lib, rule):
    """C2 as BASELINE states it: 100k self-intersecting polygons, 4096^2, 4 samples, both fill rules - every pixel of the frame against
    the reference's tessellation (even-odd: triangle fans + stencil; non-zero: libtess triangles) rasterised by the oracle.
    Even-odd must be identical.  Non-zero differs where libtess inserts new float vertices at self-intersections:
    inside the 1/255 @ p99.9 bar, and identical to the oracle's own winding != 0."""
    if not oracle_lib.ref_available():
        pytest.skip("oracle/_ref was not built")
    polys, cols = scenes.polygons_c2(100000, 4096, 1)
    s = v.Surface(dev4, 4096, 4096)
    c = v.Context(s)
    assert c.replay(*_c2_stream(polys, cols, rule).arrays()) == 0
    c.flush()
    img = s.pixels()
    c.close()
    s.close()
    r = oracle_lib.Ref(4096, 4096, 4)
    o = oracle_lib.Oracle(4096, 4096, 4)
    _emit_c2(r, polys, cols, rule)
    r.render_with(o)
    ref = o.pixels()
    r.close()
    o.close()
    st = pixel_stats(img, ref)
    print("C2 rule=%d vs reference:" % rule, st)
    if rule == 0:
        assert st["n_diff"] == 0, st
    else:
        assert st["p99_9"] <= 1 and st["frac_diff"] < 1e-3, st
        # the oracle's own non-zero (winding != 0 on the original edges) on a window of the same scene: identical
        x0, y0, w, h = 1024, 2048, 512, 256
        ow = oracle_lib.Oracle(4096, 4096, 4, window=(x0, y0, w, h))
        keep = [i for i, p in enumerate(polys) if p[:, 0].max() >= x0 - 1 and p[:, 0].min() <= x0 + w + 1 and p[:, 1].max() >= y0 - 1 and p[:, 1].min() <= y0 + h + 1]
        _emit_c2(ow, [polys[i] for i in keep], cols[keep], rule)
        assert np.array_equal(ow.pixels(), img[y0:y0 + h, x0:x0 + w])


def test_c3_million_segment_dashed_stroke_vs_reference(dev4, oracle_lib):
    """The north star's named target: the 1M-segment polyline, width 3, round joins and caps, dash {10, 6}, 4096^2.
    Geometry: the vertex stream of the CUDA stroker (dash phase from a float64 scan of float32 segment lengths) against the
    reference's (float32 phase carried segment to segment with fmodf, src/vkvg_context_internal.c:1252-1262): same vertices within
    1e-3 px in the same order, except at the few sites where a dash boundary falls within float32 rounding of a joint and the two
    phases put it on different sides (the boundary cap is then built from the neighbouring segment's normal, and a join appears or
    disappears).  Pixels: the whole 4096^2 frame against the reference's draw list through the oracle."""
    if not oracle_lib.ref_available():
        pytest.skip("oracle/_ref was not built")
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
    c.flush()
    img = s.pixels()
    c.close()
    s.close()

    r = oracle_lib.Ref(4096, 4096, 4)
    r.set_source_rgba(0.1, 0.2, 0.8, 1.0)
    r.set_line_width(3.0)
    r.set_line_join(1)
    r.set_line_cap(1)
    r.set_dash([10.0, 6.0], 0.0)
    r.polyline(pts)
    r.stroke()
    rverts, rinds = r.cached_vertices(), r.cached_indices()
    al = align_vertices(verts, rverts)
    # Two kinds of sites.  (1) ONE extra vertex on either side: a round join / cap sizes its arc with `while (a < a1)` in float32
    # (src/vkvg_context_internal.c:1111-1121, :1178-1237); where a1 - a comes within an ulp of a whole number of steps, CUDA's acosf / atan2f
    # (<= 2 ulp from glibc's) decide the last step differently and the arc gains or loses a vertex that coincides with its neighbour
    # (a zero-area triangle: no sample changes).  (2) everything else: dash boundaries that the two phases put on different segments.
    single = [t for t in al["sites"] if t[2] + t[3] == 1]
    other = [t for t in al["sites"] if t[2] + t[3] != 1]
    print("C3 geometry: %d / %d vertices (CUDA / reference), %d matched in order, max drift %.2e px; %d arc-length knife edges (one coincident vertex more or "
          "less), %d other sites %s; %d / %d indices" % (len(verts), len(rverts), al["matched"], al["max_drift"], len(single), len(other), other[:12], len(inds), len(rinds)))
    assert al["ok"], al["stopped_at"]
    assert al["max_drift"] <= 1e-3
    assert len(other) <= 64 and len(single) <= 0.005 * len(rverts)
    assert al["matched"] >= 0.995 * len(rverts)
    # the extra vertex of a type-(1) site lies within 1e-3 px of a neighbouring vertex of the same stream
    for (i, j, da, db) in single[:2000]:
        arr, k = (verts, i) if da else (rverts, j)
        near = min(np.abs(arr[k] - arr[k - 1]).max() if k > 0 else 9.0, np.abs(arr[k] - arr[k + 1]).max() if k + 1 < len(arr) else 9.0)
        assert near <= 2e-3, (i, j, da, db, float(near))

    o = oracle_lib.Oracle(4096, 4096, 4)
    r.render_with(o)
    ref = o.pixels()
    r.close()
    o.close()
    st = pixel_stats(img, ref)
    print("C3 pixels vs reference:", st)
    assert st["p99_9"] <= 1 and st["frac_diff"] < 1e-4, st
    # (differences are single samples: next to the sites above, and where a vertex that drifted by < 1e-3 px snaps to the neighbouring
    #  1/256 grid point)
    assert st["max_diff"] <= 64 * 2, st


def _emit_c4(g, paths):
    g.set_fill_rule(1)
    for pts, kind, (cx, cy), stops in paths:
        if kind == 0:
            g.set_source_linear(cx - 64, cy - 64, cx + 64, cy + 64, stops)
        else:
            g.set_source_radial(cx, cy, 4.0, cx + 8, cy - 8, 96.0, stops)
        g.move_to(float(pts[-1, 2, 0]), float(pts[-1, 2, 1]))
        for sg in pts:
            g.curve_to(*[float(x) for x in sg.ravel()])
        g.close_path()
        g.fill()


def test_c4_windows_of_the_real_scene(dev4, oracle_lib):
    """C4 as BASELINE states it (50k closed cubic paths, 3-stop linear / radial gradients, alpha 1 / 0.5, non-zero, 8192^2): the whole
    scene on the GPU, three 384^2 windows of it against (a) the oracle's restatement - identical - and (b) the reference's own
    flattening + libtess triangles through the oracle - inside the 1/255 @ p99.9 bar."""
    size = 8192
    paths = scenes.curves_c4(50000, size, 1)
    s = v.Surface(dev4, size, size)
    c = v.Context(s)
    _emit_c4(c, paths)
    c.flush()
    img = s.pixels()
    c.close()
    s.close()
    assert img[..., 3].mean() > 20
    for (x0, y0) in ((3000, 5000), (0, 0), (7700, 4100)):
        w = h = 384
        keep = [p for p in paths if p[0][..., 0].max() >= x0 - 2 and p[0][..., 0].min() <= x0 + w + 2 and p[0][..., 1].max() >= y0 - 2 and p[0][..., 1].min() <= y0 + h + 2]
        assert len(keep) > 50
        o = oracle_lib.Oracle(size, size, 4, window=(x0, y0, w, h))
        _emit_c4(o, keep)
        crop = img[y0:y0 + h, x0:x0 + w]
        st = pixel_stats(crop, o.pixels())
        print("C4 window (%d, %d): %d paths, vs oracle %s" % (x0, y0, len(keep), st))
        assert st["n_diff"] == 0, st
        o.close()
        if oracle_lib.ref_available():
            r = oracle_lib.Ref(size, size, 4)
            o2 = oracle_lib.Oracle(size, size, 4, window=(x0, y0, w, h))
            _emit_c4(r, keep)
            r.render_with(o2)
            st = pixel_stats(crop, o2.pixels())
            print("C4 window (%d, %d) vs reference tessellation: %s" % (x0, y0, st))
            assert st["p99_9"] <= 1 and st["frac_diff"] < 5e-3, st
            r.close()
            o2.close()


def test_c5a_stripe_of_the_real_scene(dev4, oracle_lib):
    """C5a's sharding unit on the real scene (16384^2, 238k C2-style polygons + 25 dashed 100k-segment polylines, bench.py's scene): the
    tile-row stripe a rank of an 8-GPU run renders (rows 6144..8191, vkvg_b200_surface_create_stripe) against the oracle on a
    full-width window of the logical surface inside it.  Polygons (pre-filtered by their bounding boxes, order kept): identical.
    Then the polylines on top - the dash phase depends on the whole prefix, so the oracle strokes each one completely: within the
    pixel bar (round joins size their arcs with float trigonometry, CUDA's and glibc's differ in the last place: single samples)."""
    size = 16384
    polys, cols = scenes.polygons_c2(238000, size, 1)
    lines = [scenes.polyline_c3(100001, size, 100 + i) for i in range(25)]
    y0, h = 6144, 2048
    wy, wh = y0 + 512, 96
    s = v.Surface(dev4, size, h, full_height=size, origin_y=y0)
    c = v.Context(s)
    assert c.replay(*_c2_stream(polys, cols, 1).arrays()) == 0
    c.flush()
    img = s.pixels()[wy - y0:wy - y0 + wh].copy()
    o = oracle_lib.Oracle(size, size, 4, window=(0, wy, size, wh))
    keep = [i for i, p in enumerate(polys) if p[:, 1].max() >= wy - 2 and p[:, 1].min() <= wy + wh + 2]
    assert len(keep) > 1000
    _emit_c2(o, [polys[i] for i in keep], cols[keep], 1)
    st = pixel_stats(img, o.pixels())
    print("C5a stripe rows %d..%d, %d of %d polygons reach the window, vs oracle: %s" % (wy, wy + wh, len(keep), len(polys), st))
    assert st["n_diff"] == 0, st

    def strokes(g, poly):
        g.set_line_width(3.0)
        g.set_line_join(1)
        g.set_line_cap(1)
        g.set_dash([10.0, 6.0], 0.0)
        for i, pts in enumerate(lines):
            g.set_source_rgba(0.1 + 0.03 * i, 0.2, 0.8 - 0.02 * i, 1.0)
            poly(g, pts)
            g.stroke()

    cs = v.CommandStream()
    strokes(cs, lambda g, pts: g.polyline(pts))
    assert c.replay(*cs.arrays()) == 0
    c.flush()
    img = s.pixels()[wy - y0:wy - y0 + wh].copy()
    strokes(o, lambda g, pts: g.polyline(pts))
    st = pixel_stats(img, o.pixels())
    print("C5a stripe with the 25 dashed polylines on top, vs oracle: %s" % st)
    assert st["p99_9"] <= 1 and st["frac_diff"] < 1e-4, st
    c.close()
    s.close()
    o.close()
