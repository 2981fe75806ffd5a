"""Second golden set, generated like tests/golden/make_golden.py from the REFERENCE's own object code
(oracle/_ref/libvkvg_ref.so = unmodified /root/reference sources behind the recording shim):

    python -m tests.golden.make_golden2

  elliptic.npz   flattened points + `vkvg_path_extents` of seeded paths built with vkvg_elliptic_arc_to,
                 vkvg_rel_elliptic_arc_to, vkvg_rounded_rectangle2 and vkvg_ellipse (src/vkvg_context.c:665-696, :1579-1639,
                 src/vkvg_context_internal.c:1473-1580, :1879-1917)
  clip.npz       resolved pixels of seeded scenes using vkvg_clip / clip_preserve / reset_clip / save / restore: the
                 reference's recorded stencil + colour draws rasterised by the oracle's Vulkan restatement."""
import ctypes as C
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import Oracle, Ref  # noqa: E402
from tests import scenes  # noqa: E402

ELLIPTIC_SEEDS = list(range(16))


def elliptic_scene(g, seed):
    """one path per seed mixing lines with SVG-style elliptical arcs (all flag combinations, rotated, radii too small
    for the chord, zero radius), rounded_rectangle2 and ellipse."""
    r = scenes.SplitMix64(9100 + seed)
    if seed % 4 == 3:
        g.scale(1.0 + r.uniform(0, 2), 1.0 + r.uniform(0, 2))
    g.move_to(r.uniform(20, 200), r.uniform(20, 200))
    for k in range(3 + seed % 4):
        large, sweep = bool((seed + k) & 1), bool((seed + k) & 2)
        rx, ry = r.uniform(5, 90), r.uniform(5, 90)
        if k == 2 and seed % 5 == 0:
            rx = 0.0
        phi = r.uniform(-3.2, 3.2) if seed % 3 else 0.0
        if k % 2:
            g.rel_elliptic_arc_to(r.uniform(-60, 60), r.uniform(-60, 60), large, sweep, rx, ry, phi)
        else:
            g.elliptic_arc_to(r.uniform(10, 240), r.uniform(10, 240), large, sweep, rx, ry, phi)
        if k == 1:
            g.line_to(r.uniform(10, 240), r.uniform(10, 240))
    if seed % 2:
        g.close_path()
    g.rounded_rectangle2(r.uniform(10, 100), r.uniform(10, 100), r.uniform(60, 120), r.uniform(60, 120), r.uniform(2, 25), r.uniform(2, 25))
    if seed % 3 == 1:
        g.new_sub_path()
        g.elliptic_arc_to(r.uniform(10, 240), r.uniform(10, 240), True, False, 30.0, 12.0, 0.5)   # arc that opens a path
    g.ellipse(r.uniform(10, 40), r.uniform(10, 40), r.uniform(60, 200), r.uniform(60, 200), r.uniform(0, 3))


def ref_path_extents(r):
    f = C.c_float
    x1, y1, x2, y2 = f(), f(), f(), f()
    fn = r._lib.vkvg_path_extents
    fn.argtypes = [C.c_void_p] + [C.POINTER(f)] * 4
    fn(r._ctx, C.byref(x1), C.byref(y1), C.byref(x2), C.byref(y2))
    return np.array([x1.value, y1.value, x2.value, y2.value], np.float32)


def main_elliptic():
    out = {}
    for seed in ELLIPTIC_SEEDS:
        r = Ref(256, 256, 4, record=False)
        elliptic_scene(r, seed)
        out["ext_%d" % seed] = ref_path_extents(r)
        out["pts_%d" % seed] = r.path_points()
        out["tab_%d" % seed] = r.path_table()
        r.close()
    r = Ref(64, 64, 4, record=False)
    out["ext_empty"] = ref_path_extents(r)
    r.close()
    np.savez_compressed(os.path.join(HERE, "elliptic.npz"), **out)
    print("elliptic.npz", os.path.getsize(os.path.join(HERE, "elliptic.npz")), "bytes")


if __name__ == "__main__":
    main_elliptic()
    if "main_clip" in globals():
        globals()["main_clip"]()
