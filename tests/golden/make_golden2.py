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


CLIP_SCENES = ["rect_eo", "circle_nz", "nested", "save_restore", "reset", "preserve_stroke", "deep_stack", "restore_quirk", "clear_inside",
               "op_clear", "op_difference", "op_mixed"]


def _blob(g, r, n, size):
    """a few overlapping translucent shapes covering most of the surface"""
    for k in range(n):
        g.set_source_rgba(r.u(), r.u(), r.u(), 0.35 + 0.6 * r.u())
        if k % 3 == 0:
            g.rectangle(r.uniform(-10, size * 0.6), r.uniform(-10, size * 0.6), r.uniform(20, size), r.uniform(20, size))
            g.fill()
        elif k % 3 == 1:
            g.arc(r.uniform(0, size), r.uniform(0, size), r.uniform(10, size * 0.5), 0.0, 6.2831855)
            g.fill()
        else:
            g.set_line_width(r.uniform(2, 9))
            g.move_to(r.uniform(0, size), r.uniform(0, size))
            g.line_to(r.uniform(0, size), r.uniform(0, size))
            g.line_to(r.uniform(0, size), r.uniform(0, size))
            g.stroke()


def _star(g, cx, cy, ro, ri, n=5):
    import math
    for i in range(2 * n):
        a = math.pi * i / n
        rr = ro if i % 2 == 0 else ri
        (g.move_to if i == 0 else g.line_to)(cx + rr * math.sin(a), cy - rr * math.cos(a))
    g.close_path()


def clip_scene(g, name, seed, size=128):
    """scenes exercising vkvg_clip / clip_preserve / reset_clip and the clip part of save / restore, including the
    reference's bookkeeping quirks (src/vkvg_context.c:698-795, :1251-1512); drawn on any of Ref / Oracle / Context."""
    r = scenes.SplitMix64(8800 + seed)
    s = float(size)
    if name == "rect_eo":
        g.set_fill_rule(0)
        g.rectangle(s * 0.2 + seed, s * 0.15, s * 0.55, s * 0.6)
        g.rectangle(s * 0.35, s * 0.3 + seed, s * 0.2, s * 0.2)     # hole under even-odd
        g.clip()
        _blob(g, r, 5, size)
    elif name == "circle_nz":
        g.set_fill_rule(1)
        g.arc(s * 0.5, s * 0.5, s * 0.33 + seed, 0.0, 6.2831855)
        g.clip()
        g.set_source_linear(0, 0, s, s, [(0, 1, 0, 0, 1), (0.5, 0, 1, 0, 0.6), (1, 0, 0, 1, 1)])
        g.paint()
        _blob(g, r, 3, size)
    elif name == "nested":
        g.set_fill_rule(1)
        _star(g, s * 0.5, s * 0.5, s * 0.48, s * 0.2)
        g.clip()
        _blob(g, r, 2, size)
        g.set_fill_rule(0)
        g.rectangle(s * 0.1, s * 0.4, s * 0.8, s * 0.35)
        g.clip()                                               # intersection of star and rectangle
        g.set_source_rgba(0.1, 0.1, 0.9, 0.8)
        g.paint()
    elif name == "save_restore":
        g.rectangle(s * 0.1, s * 0.1, s * 0.8, s * 0.8)
        g.clip()
        g.save()
        g.arc(s * 0.5, s * 0.5, s * 0.25, 0.0, 6.2831855)
        g.clip()
        g.set_source_rgba(0.9, 0.2, 0.1, 0.7)
        g.paint()
        g.restore()                                            # back to the rectangle
        g.set_source_rgba(0.1, 0.8, 0.2, 0.4)
        g.paint()
        _blob(g, r, 2, size)
    elif name == "reset":
        g.rectangle(s * 0.3, s * 0.3, s * 0.3, s * 0.3)
        g.clip()
        _blob(g, r, 2, size)
        g.reset_clip()
        g.set_source_rgba(0.2, 0.2, 0.8, 0.3)
        g.paint()
        g.reset_clip()                                         # second reset is a no-op
        _blob(g, r, 2, size)
    elif name == "preserve_stroke":
        g.set_fill_rule(0)
        _star(g, s * 0.5, s * 0.5, s * 0.45, s * 0.18, 7)
        g.clip_preserve()
        g.set_source_rgba(0.8, 0.7, 0.1, 1.0)
        g.set_line_width(9.0)
        g.stroke()                                             # only the inner half of the stroke survives
        g.set_source_rgba(0.1, 0.3, 0.7, 0.5)
        g.rectangle(0, s * 0.45, s, s * 0.2)
        g.fill()
    elif name == "deep_stack":
        for k in range(5):                                     # five nested clip saves: save bits 2..6
            g.rectangle(s * 0.04 * (k + 1), s * 0.03 * (k + 1), s * (0.92 - 0.08 * k), s * (0.94 - 0.06 * k))
            g.clip()
            g.save()
        g.arc(s * 0.5, s * 0.5, s * 0.2, 0.0, 6.2831855)
        g.clip()
        g.set_source_rgba(1, 0, 0, 0.6)
        g.paint()
        for k in range(5):
            g.restore()
            g.set_source_rgba(0.2 * k, 1 - 0.2 * k, 0.5, 0.35)
            g.paint()
    elif name == "restore_quirk":
        # after a restore the context believes no clip is active: a later save / clip / restore pair wipes the stencil
        g.rectangle(s * 0.2, s * 0.2, s * 0.6, s * 0.6)
        g.clip()
        g.save()
        g.restore()
        g.save()
        g.arc(s * 0.5, s * 0.5, s * 0.2, 0.0, 6.2831855)
        g.clip()
        g.set_source_rgba(0.9, 0.1, 0.1, 0.8)
        g.paint()
        g.restore()
        g.set_source_rgba(0.1, 0.1, 0.9, 0.4)
        g.paint()                                              # unclipped in the reference
    elif name == "clear_inside":
        g.rectangle(s * 0.25, s * 0.25, s * 0.5, s * 0.5)
        g.clip()
        _blob(g, r, 2, size)
        g.clear()                                              # wipes colour and the clip
        g.set_source_rgba(0.3, 0.9, 0.3, 0.5)
        g.arc(s * 0.5, s * 0.5, s * 0.45, 0.0, 6.2831855)
        g.fill()
    elif name in ("op_clear", "op_difference", "op_mixed"):
        # vkvg_set_operator: CLEAR (0) and DIFFERENCE (3) have pipelines of their own, SOURCE (1) draws like OVER (2)
        g.set_source_rgba(0.2, 0.5, 0.8, 1.0)
        g.paint()
        _blob(g, r, 3, size)
        ops = {"op_clear": [0], "op_difference": [3], "op_mixed": [3, 0, 1, 2]}[name]
        for i, op in enumerate(ops):
            g.set_operator(op)
            g.set_source_rgba(0.9 - 0.2 * i, 0.3 + 0.1 * i, 0.2, 0.55 + 0.15 * (i % 2))
            g.arc(s * (0.3 + 0.15 * i), s * (0.35 + 0.1 * i), s * 0.22, 0.0, 6.2831855)
            g.fill()
            g.set_line_width(6.0)
            g.move_to(s * 0.1, s * (0.2 + 0.2 * i))
            g.line_to(s * 0.9, s * (0.3 + 0.15 * i))
            g.stroke()
            if seed:
                g.set_fill_rule(0)
                _star(g, s * 0.6, s * 0.6, s * 0.3, s * 0.12)
                g.fill()
                g.set_fill_rule(1)
        g.set_operator(2)
    else:
        raise KeyError(name)


def main_clip():
    pix = {}
    for name in CLIP_SCENES:
        for seed in range(2):
            r, o = Ref(128, 128, 4), Oracle(128, 128, 4)
            clip_scene(r, name, seed)
            r.render_with(o)
            pix["%s_%d" % (name, seed)] = o.pixels()
            r.close()
            o.close()
    np.savez_compressed(os.path.join(HERE, "clip.npz"), **pix)
    print("clip.npz", os.path.getsize(os.path.join(HERE, "clip.npz")), "bytes")


SURF_SEQS = list(range(8))


def checker(w, h, seed=0):
    """premultiplied RGBA8 test image: coloured checker with a translucent diagonal band"""
    y, x = np.mgrid[0:h, 0:w]
    a = np.where((x + y + seed) % 7 < 2, 128, 255).astype(np.float32)
    rgb = np.stack([(x * 37 + seed * 11) % 256, (y * 53) % 256, ((x // 3 + y // 3) % 2) * 200 + 30], -1).astype(np.float32)
    img = np.concatenate([np.floor(rgb * a[..., None] / 255.0), a[..., None]], -1)
    return np.ascontiguousarray(img.astype(np.uint8))


def surface_sequence(g, seed, set_source):
    """a sequence of CTM changes, set_source_surface and pattern-matrix calls; set_source(g, **kw) applies the source.  Returns
    after every step what the caller wants to record (the test reads the push state through its own accessor)."""
    r = scenes.SplitMix64(6600 + seed)
    steps = []
    g.translate(r.uniform(-5, 20), r.uniform(-5, 20))
    steps.append("t")
    if seed % 2:
        g.rotate(r.uniform(-1, 1))
    set_source(g, x=r.uniform(-10, 30), y=r.uniform(-10, 30))
    steps.append("s")
    g.scale(r.uniform(0.5, 2.5), r.uniform(0.5, 2.5))
    steps.append("c")
    set_source(g, x=r.uniform(-10, 30), y=r.uniform(-10, 30), extend=seed % 4, filter=4 if seed % 3 == 0 else 3,
               matrix=[r.uniform(0.5, 2), r.uniform(-0.3, 0.3), r.uniform(-0.3, 0.3), r.uniform(0.5, 2), r.uniform(-8, 8), r.uniform(-8, 8)])
    steps.append("p")
    if seed % 3 == 1:
        g.save()
        g.translate(3.0, -2.0)
        g.restore()
    else:
        g.translate(r.uniform(-3, 3), 1.0)   # a CTM change drops the pattern matrix from matInv (reference behaviour)
    steps.append("e")
    return steps


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
