"""Generates the committed golden vectors under tests/golden/ from the REFERENCE's own object code.

Run in the build container (needs /root/reference so that `make -C oracle ref` can compile the unmodified
reference tessellation sources into oracle/_ref/libvkvg_ref.so):

    python tests/golden/make_golden.py

What is pinned, and by what:
  geometry.npz   flattened points, `pathes` tables, stroke vertices and indices for seeded scenes, produced by the
                 reference's vkvg_* entry points (src/vkvg_context.c, src/vkvg_context_internal.c) — no oracle code involved.
  pixels.npz     resolved premultiplied RGBA8 images of seeded scenes and of tiger.svg at 1024x1024: the reference's
                 recorded draw list (its exact vertex/index/uniform bytes and pipeline state) rasterised by the
                 Vulkan restatement in oracle/vkvg_oracle.c (ovk_raster_ref_drawlist).  The reference has no pixels
                 of its own without a Vulkan ICD, so the rasterisation rules are this repo's definition.
The GPU box has no /root/reference; tests there compare against these files.
"""
import os
import sys
import zlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import oracle  # noqa: E402
from oracle import Oracle, Ref  # noqa: E402
from tests import scenes  # noqa: E402

GEOMETRY_SEEDS = list(range(24))
PIXEL_SCENES = ["eo", "nz_convex", "stroke_opaque", "stroke_alpha", "stroke_dash", "grad_linear", "grad_radial", "paint", "mixed"]


def geometry_scene(g, seed):
    """path + stroke state for one seed (shared by the generator and the tests)."""
    scenes.random_path(g, seed)
    g.set_line_width(1 + seed % 7)
    g.set_line_join(seed % 3)
    g.set_line_cap((seed // 3) % 3)
    if seed % 2:
        g.set_dash([10, 6], 3.0)


def pixel_scene(g, name, seed, size=128):
    """one of PIXEL_SCENES drawn on g (any of Oracle / Ref / vkvg_b200.Context)."""
    r = scenes.SplitMix64(7000 + seed)
    lin = [(0, 1, 0, 0, 1), (0.5, 0, 1, 0, 0.5), (1, 0, 0, 1, 1)]
    rad = [(0, 1, 1, 0, 1), (0.6, 0, 1, 1, 0.7), (1, 1, 0, 1, 1)]
    if name == "nz_convex":
        g.set_fill_rule(1)
        g.set_source_rgba(0.2, 0.7, 0.3, 0.8)
        g.arc(size * 0.5, size * 0.5, size * 0.3, 0.0, 6.2831855)
        g.fill()
        g.set_source_rgba(0.9, 0.1, 0.3, 0.5)
        g.rectangle(size * 0.1, size * 0.2, size * 0.5, size * 0.4)
        g.fill()
        return
    if name == "paint":
        g.set_source_rgba(0.3, 0.6, 0.9, 0.5)
        g.paint()
        g.set_source_linear(0, 0, size, size, lin)
        g.paint()
        return
    n_paths = 3 if name == "mixed" else 1
    for k in range(n_paths):
        g.new_path()
        scenes.random_path(g, seed * 16 + k, size=size)
        kind = name if name != "mixed" else ["eo", "grad_radial", "stroke_dash"][k]
        if kind in ("eo", "grad_linear", "grad_radial"):
            g.set_fill_rule(0)
            if kind == "eo":
                g.set_source_rgba(r.u(), r.u(), r.u(), 0.3 + 0.7 * r.u())
            elif kind == "grad_linear":
                g.set_source_linear(size * 0.1, size * 0.15, size * 0.8, size * 0.9, lin)
            else:
                g.set_source_radial(size * 0.5, size * 0.5, size * 0.05, size * 0.55, size * 0.45, size * 0.4, rad)
            g.fill()
        else:
            g.set_source_rgba(r.u(), r.u(), r.u(), 0.5 if kind == "stroke_alpha" else 1.0)
            g.set_line_width(1 + (seed + k) % 6)
            g.set_line_join((seed + k) % 3)
            g.set_line_cap((seed + k + 1) % 3)
            g.set_dash([8, 5], 1.0) if kind == "stroke_dash" else g.set_dash([])
            g.stroke()


def main():
    oracle.build(ref=True)
    geo = {}
    for seed in GEOMETRY_SEEDS:
        r = Ref(256, 256, 4)
        geometry_scene(r, seed)
        geo["pts_%d" % seed] = r.path_points()
        geo["tab_%d" % seed] = r.path_table()
        r.stroke_preserve()
        geo["verts_%d" % seed] = r.cached_vertices()
        geo["inds_%d" % seed] = r.cached_indices()
        r.close()
    np.savez_compressed(os.path.join(HERE, "geometry.npz"), **geo)

    pix = {}
    for name in PIXEL_SCENES:
        for seed in range(3):
            r, o = Ref(128, 128, 4), Oracle(128, 128, 4)
            pixel_scene(r, name, seed)
            r.render_with(o)
            pix["%s_%d" % (name, seed)] = o.pixels()
            r.close()
            o.close()
    w, h, shapes = scenes.load_nsvg(os.path.join(HERE, "tiger.nsvg.bin"))
    r, o = Ref(1024, 1024, 4), Oracle(1024, 1024, 4)
    scenes.render_nsvg(r, shapes)
    r.render_with(o)
    pix["tiger_1024"] = o.pixels()
    np.savez_compressed(os.path.join(HERE, "pixels.npz"), **pix)
    for f in ("geometry.npz", "pixels.npz"):
        print(f, os.path.getsize(os.path.join(HERE, f)), "bytes")


if __name__ == "__main__":
    main()
