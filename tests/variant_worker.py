"""Worker of tests/test_gpu_variants.py: renders fixed scenes on cuda:0 with whatever kernel variants the environment selects
(VKVG_B200_STROKE, VKVG_B200_FLATTEN are read once per process by the library) and prints one JSON line of SHA-256 digests -
flattened points, stroke vertices and indices, stroke edges (sorted: the order in which warps reserve room is not fixed) and pixels."""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import vkvg_b200 as v  # noqa: E402
from tests import scenes  # noqa: E402


def digest(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def sorted_rows(e):
    return e[np.lexsort(e.T[::-1])] if len(e) else e


def polyline(c, pts, close=False):
    c.move_to(float(pts[0, 0]), float(pts[0, 1]))
    for p in pts[1:]:
        c.line_to(float(p[0]), float(p[1]))
    if close:
        c.close_path()


def stroke_cases():
    """(name, points, closed, width, join, cap, dashes): more than three vertices per item (staged blocks), miter joins of two vertices
    (blocks that write straight to global memory), a stroke so wide that a block's output exceeds the staging area, closed paths."""
    walk = scenes.polyline_c3(3001, 512, 5)
    ring = np.stack([256 + 200 * np.cos(np.linspace(0, 2 * np.pi, 700, endpoint=False)), 256 + 200 * np.sin(np.linspace(0, 2 * np.pi, 700, endpoint=False))], 1).astype(np.float32)
    zig = scenes.polyline_c3(300, 512, 9)
    return [
        ("dashed_round", walk, False, 3.0, 1, 1, [10.0, 6.0]),
        ("miter", walk[:1500], False, 2.0, 0, 0, []),
        ("wide_round", zig, False, 400.0, 1, 1, []),   # arc step pi / 80: about 16 vertices per join
        ("closed_round", ring, True, 5.0, 1, 0, []),
        ("closed_dashed_bevel", ring, True, 4.0, 2, 2, [7.0, 3.0, 2.0]),
    ]


def main():
    dev = v.Device(4)
    out = {}
    # ---- flatten: cubics and arcs of every length (the warp-per-curve counting pass serves batches that keep the 64-point cache) ----
    s = v.Surface(dev, 512, 512)
    c = v.Context(s)
    for seed in range(40):
        scenes.random_path(c, 100 + seed, 512)
    c.move_to(10.0, 10.0)
    c.curve_to(2000.0, -1500.0, -1800.0, 2400.0, 500.0, 500.0)     # a long curve: hundreds of points
    c.curve_to(500.0, 500.0, 500.0, 500.0, 500.0, 500.0)            # a degenerate one
    c.curve_to(501.0, 500.5, 502.0, 501.0, 503.0, 501.5)            # a nearly straight one (leaf at depth 1)
    pts = c.path_points()
    first, cnt, cur = c.path_subpaths()
    out["flatten_points"] = digest(pts)
    out["flatten_subpaths"] = digest(np.concatenate([first, cnt])) + digest(cur)
    out["n_points"] = int(len(pts))
    c.set_source_rgba(0.2, 0.4, 0.9, 0.7)
    c.set_fill_rule(1)
    c.fill()
    c.flush()
    out["flatten_pixels"] = digest(s.pixels())
    # ---- strokes ----
    for name, p, closed, width, join, cap, dashes in stroke_cases():
        s = v.Surface(dev, 512, 512)
        c = v.Context(s)
        c.set_line_width(width)
        c.set_line_join(join)
        c.set_line_cap(cap)
        if dashes:
            c.set_dash(dashes, 1.5)
        polyline(c, p, closed)
        verts, inds = c.stroke_geometry()
        edges = c.path_edges(stroke=True)
        out[name + "_verts"] = digest(verts)
        out[name + "_inds"] = digest(inds)
        out[name + "_edges"] = digest(sorted_rows(edges))
        out[name + "_n"] = [int(len(verts)), int(len(inds)), int(len(edges))]
        c.set_source_rgba(0.9, 0.3, 0.1, 1.0 if name == "wide_round" else 0.6)   # (opaque where every pixel is covered hundreds of times)
        c.stroke()
        c.flush()
        out[name + "_pixels"] = digest(s.pixels())
    print(json.dumps(out))


if __name__ == "__main__":
    main()
