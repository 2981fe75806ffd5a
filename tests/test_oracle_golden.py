"""CPU: the oracle (oracle/vkvg_oracle.c) against the committed golden vectors that the reference's own object
code produced (tests/golden/make_golden.py).  Geometry is bit-exact; pixels are bit-exact."""
import os

import numpy as np
import pytest

from tests import scenes
from tests.golden import make_golden as mg

GOLD = os.path.dirname(os.path.abspath(mg.__file__))


@pytest.fixture(scope="module")
def geo():
    return np.load(os.path.join(GOLD, "geometry.npz"))


@pytest.fixture(scope="module")
def pix():
    return np.load(os.path.join(GOLD, "pixels.npz"))


@pytest.mark.parametrize("seed", mg.GEOMETRY_SEEDS)
def test_oracle_geometry_matches_reference_golden(oracle_lib, geo, seed):
    o = oracle_lib.Oracle(256, 256, 4)
    mg.geometry_scene(o, seed)
    assert np.array_equal(o.path_points(), geo["pts_%d" % seed])
    assert np.array_equal(o.path_table(), geo["tab_%d" % seed])
    o.stroke_preserve()
    assert np.array_equal(o.last_vertices(), geo["verts_%d" % seed])
    assert np.array_equal(o.last_indices(), geo["inds_%d" % seed])


@pytest.mark.parametrize("name", mg.PIXEL_SCENES)
def test_oracle_pixels_match_reference_drawlist_golden(oracle_lib, pix, name):
    for seed in range(3):
        o = oracle_lib.Oracle(128, 128, 4)
        mg.pixel_scene(o, name, seed)
        assert np.array_equal(o.pixels(), pix["%s_%d" % (name, seed)]), (name, seed)


def test_oracle_tiger_matches_golden(oracle_lib, pix):
    w, h, shapes = scenes.load_nsvg(os.path.join(GOLD, "tiger.nsvg.bin"))
    assert (w, h, len(shapes)) == (900.0, 900.0, 239)
    o = oracle_lib.Oracle(1024, 1024, 4)
    scenes.render_nsvg(o, shapes)
    assert np.array_equal(o.pixels(), pix["tiger_1024"])


def test_winding_brute_definition(oracle_lib):
    """unit square (24.8 fixed point) covers exactly the samples inside it, both orientations, all sample counts."""
    sq = np.array([[256, 256, 768, 256], [768, 256, 768, 768], [768, 768, 256, 768], [256, 768, 256, 256]], np.int32)
    for s in (1, 2, 4, 8, 16):
        w = oracle_lib.winding_brute(sq, 4, 4, s)
        assert (np.abs(w[1:3, 1:3]) == 1).all() and len(np.unique(w[1:3, 1:3])) == 1
        w[1:3, 1:3] = 0
        assert not w.any()
        wr = oracle_lib.winding_brute(sq[::-1, [2, 3, 0, 1]], 4, 4, s)
        assert (wr[1:3, 1:3] == -oracle_lib.winding_brute(sq, 4, 4, s)[1:3, 1:3]).all()


def test_winding_brute_empty_and_degenerate(oracle_lib):
    assert not oracle_lib.winding_brute(np.zeros((0, 4), np.int32), 8, 8, 4).any()
    assert not oracle_lib.winding_brute(np.array([[100, 100, 100, 100], [0, 300, 900, 300]], np.int32), 8, 8, 4).any()


def test_oracle_fill_rule_and_small_subpaths(oracle_lib):
    """sub-paths with <= 2 points are ignored by fills (reference internal.c:1617,:1759)."""
    o = oracle_lib.Oracle(32, 32, 4)
    o.set_source_rgba(1, 0, 0, 1)
    o.move_to(2, 2)
    o.line_to(30, 30)
    o.fill()
    assert not o.pixels().any()
    o.move_to(4, 4)
    o.line_to(28, 4)
    o.line_to(28, 28)
    o.fill()
    assert o.pixels()[..., 3].any()


def test_oracle_invalid_dash_status(oracle_lib):
    o = oracle_lib.Oracle(32, 32, 4)
    o.set_dash([0.0, 0.0])
    o.move_to(2, 2)
    o.line_to(30, 30)
    o.stroke()
    assert o.status() == 13  # VKVG_STATUS_INVALID_DASH (include/vkvg.h:125-150)


def test_windowed_oracle_equals_the_crop_of_the_whole_surface():
    """Oracle(..., window=...) (test infrastructure for the 8192^2 / 16384^2 configs) stores and rasterises only a window of the logical
    surface; it must hold exactly the pixels of that region of the whole-surface render - fills of both rules, gradients, strokes,
    clip / save / restore, and the reference's recorded draw list."""
    import oracle
    from oracle import Oracle
    from tests.golden import make_golden2 as mg2
    for name in mg.PIXEL_SCENES:
        full = Oracle(128, 128, 4)
        mg.pixel_scene(full, name, 1)
        a = full.pixels()
        for (x0, y0, w, h) in ((16, 32, 64, 48), (0, 0, 128, 16), (100, 90, 28, 38)):
            o = Oracle(128, 128, 4, window=(x0, y0, w, h))
            mg.pixel_scene(o, name, 1)
            assert np.array_equal(o.pixels(), a[y0:y0 + h, x0:x0 + w]), (name, x0, y0)
    full, o = Oracle(128, 128, 4), Oracle(128, 128, 4, window=(30, 20, 70, 90))
    for g in (full, o):
        mg2.clip_scene(g, "save_restore", 0)
    assert np.array_equal(o.pixels(), full.pixels()[20:110, 30:100])
    if oracle.ref_available():
        imgs = []
        for win in (None, (40, 8, 60, 100)):
            r = oracle.Ref(128, 128, 4)
            mg.pixel_scene(r, "mixed", 1)
            t = Oracle(128, 128, 4, window=win)
            r.render_with(t)
            imgs.append(t.pixels())
            r.close()
        assert np.array_equal(imgs[1], imgs[0][8:108, 40:100])
