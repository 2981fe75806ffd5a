"""vkvg-svg.h: the repository's own SVG parser (vkvg_b200/csrc/svg.cpp) against the reference's nanoSVG.

CPU part: every document under tests/golden/svg/ must parse into exactly the shape list nanoSVG produces
(tests/golden/svg/*.nsvg.bin, written by oracle/nsvg_dump.c through tests/golden/make_svg_golden.py): same shapes in the
same order, same paint types and colours, bit-identical control points, stroke widths and opacities.
GPU part: vkvg_svg_render of tiger.svg gives the pixels of the committed golden frame (tests/golden/pixels.npz)."""
import glob
import os

import numpy as np
import pytest

import vkvg_b200 as v
from tests import scenes

HERE = os.path.dirname(os.path.abspath(__file__))
SVG_DIR = os.path.join(HERE, "golden", "svg")
DOCS = sorted(os.path.basename(f)[:-4] for f in glob.glob(os.path.join(SVG_DIR, "*.svg")))


def _shapes(blob, tmp_path):
    p = tmp_path / "d.bin"
    p.write_bytes(blob)
    return scenes.load_nsvg(str(p))


def test_fixture_set():
    assert {"tiger", "rect", "path", "vkvg", "checkbox", "test", "extra_shapes", "extra_gradients"} <= set(DOCS)


@pytest.mark.parametrize("name", DOCS)
def test_parser_matches_nanosvg_dump(name, tmp_path):
    svg = v.Svg(os.path.join(SVG_DIR, name + ".svg"))
    mine = svg.serialize()
    gold = open(os.path.join(SVG_DIR, name + ".nsvg.bin"), "rb").read()
    if mine != gold:   # explain before failing
        wm, hm, sm = _shapes(mine, tmp_path)
        wg, hg, sg = _shapes(gold, tmp_path)
        assert (wm, hm) == (wg, hg)
        assert len(sm) == len(sg), (len(sm), len(sg))
        for i, (a, b) in enumerate(zip(sm, sg)):
            for k in a:
                if k != "paths":
                    assert a[k] == b[k], (i, k, a[k], b[k])
            assert len(a["paths"]) == len(b["paths"]), i
            for j, ((pa, ca), (pb, cb)) in enumerate(zip(a["paths"], b["paths"])):
                assert ca == cb and pa.shape == pb.shape, (i, j)
                assert np.array_equal(pa.view(np.uint32), pb.view(np.uint32)), (i, j, np.abs(pa - pb).max())
    assert mine == gold
    w, h = svg.dimensions()
    wg, hg, _ = _shapes(gold, tmp_path)
    assert (w, h) == (int(wg), int(hg))


def test_fragment_and_missing_file():
    L = v.lib()
    assert not L.vkvg_svg_load(b"/nonexistent/file.svg")
    frag = open(os.path.join(SVG_DIR, "rect.svg")).read()
    a, b = v.Svg(fragment=frag), v.Svg(os.path.join(SVG_DIR, "rect.svg"))
    assert a.serialize() == b.serialize()
    assert a.dimensions() == (400, 110)


def test_pathological_documents_terminate():
    for frag in ("", "<svg", "<svg><path d='M0 0 L'/></svg>", "<svg><g transform='translate(1,2,3) scale(2)'><rect width='4' height='4'/></g></svg>",
                 "<svg><path d='M 1e400 0 L 5 5 5 0 z A'/><polygon points='1'/><rect width='-5' height='5' rx='1'/></svg>",
                 "<svg viewBox='0 0'><circle r='5'/></svg>"):
        s = v.Svg(fragment=frag)
        assert s.serialize()[:4] == b"NSVG"


@pytest.mark.gpu
def test_tiger_svg_render_matches_golden_frame():
    gold = np.load(os.path.join(HERE, "golden", "pixels.npz"))["tiger_1024"]
    dev = v.Device(4)
    surf = v.Surface(dev, 1024, 1024)
    ctx = v.Context(surf)
    svg = v.Svg(os.path.join(SVG_DIR, "tiger.svg"))
    ctx.render_svg(svg)
    ctx.flush()
    got = surf.pixels()
    assert np.array_equal(got, gold), int((got != gold).any(axis=2).sum())
    # by id: a single shape draws less than the whole document, and a unknown id draws nothing
    surf2 = v.Surface(dev, 1024, 1024)
    c2 = v.Context(surf2)
    c2.render_svg(svg, "no-such-id")
    c2.flush()
    assert not surf2.pixels().any()
    c2.render_svg(svg, "path8")
    c2.flush()
    n = int((surf2.pixels()[..., 3] > 0).sum())
    assert 0 < n < int((gold[..., 3] > 0).sum())


@pytest.mark.gpu
def test_surface_create_from_svg():
    L = v.lib()
    dev = v.Device(4)
    h = L.vkvg_surface_create_from_svg(dev.h, 0, 0, os.path.join(SVG_DIR, "path.svg").encode())
    assert h and L.vkvg_surface_status(h) == 0
    assert (L.vkvg_surface_get_width(h), L.vkvg_surface_get_height(h)) == (400, 400)
    out = np.zeros((400, 400, 4), np.uint8)
    assert L.vkvg_b200_surface_read_premultiplied(h, out.ctypes.data) == 0
    assert out[200, 200, 3] > 0 and out[10, 10, 3] == 0   # inside the green square / outside
    L.vkvg_surface_destroy(h)
