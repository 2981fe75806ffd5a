"""first-contact GPU probe (not a pytest file): prints what breaks."""
import sys, os, traceback
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import oracle
from oracle import Oracle
import vkvg_b200 as v
from tests import scenes

def step(name, fn):
    try:
        r = fn()
        print("OK  ", name, r if r is not None else "")
    except Exception as e:
        print("FAIL", name, repr(e))
        traceback.print_exc()

dev = v.Device(4)
def winding_small():
    e = np.array([[10*256, 10*256, 50*256, 12*256], [50*256, 12*256, 30*256, 55*256], [30*256, 55*256, 10*256, 10*256]], np.int32)
    w = dev.winding(e, 64, 64)
    ref = oracle.winding_brute(e, 64, 64, 4)
    return int((w != ref).sum()), int((ref != 0).sum())
step("winding_small", winding_small)

def flatten():
    s = v.Surface(dev, 256, 256); c = v.Context(s); o = Oracle(256, 256, 4)
    for g in (c, o):
        g.move_to(10, 10); g.line_to(100, 20); g.curve_to(150, 50, 120, 150, 50, 100); g.close_path()
    a, b = c.path_points(), o.path_points()
    return a.shape, b.shape, float(np.abs(a - b).max()) if a.shape == b.shape else None
step("flatten", flatten)

def fill(rule):
    def f():
        s = v.Surface(dev, 256, 256); c = v.Context(s); o = Oracle(256, 256, 4)
        for g in (c, o):
            g.set_source_rgba(1, 0, 0, 0.5); g.set_fill_rule(rule)
            g.move_to(10, 10); g.line_to(100, 20); g.curve_to(150, 50, 120, 150, 50, 100); g.close_path(); g.fill()
        c.flush()
        a, b = s.pixels(), o.pixels()
        d = np.abs(a.astype(int) - b.astype(int)).max(axis=2)
        return int((d > 0).sum()), int(d.max()), int((b[..., 3] > 0).sum())
    return f
step("fill_eo", fill(0)); step("fill_nz", fill(1))

def stroke(dash):
    def f():
        s = v.Surface(dev, 256, 256); c = v.Context(s); o = Oracle(256, 256, 4)
        for g in (c, o):
            g.set_source_rgba(0, 0, 1, 0.7); g.set_line_width(7); g.set_line_join(1); g.set_line_cap(1)
            if dash: g.set_dash([10, 6])
            g.move_to(10, 10); g.line_to(100, 20); g.curve_to(150, 50, 120, 150, 50, 100); g.close_path()
        va, ia = c.stroke_geometry()
        o.stroke_preserve(); vb, ib = o.last_vertices(), o.last_indices()
        geo = (va.shape, vb.shape, ia.shape, ib.shape, float(np.abs(va - vb).max()) if va.shape == vb.shape else None,
               bool(np.array_equal(ia, ib)) if ia.shape == ib.shape else None)
        o.clear(); c.stroke(); o.stroke() if False else None
        return geo
    return f
step("stroke", stroke(False)); step("stroke_dash", stroke(True))

def stroke_pixels():
    s = v.Surface(dev, 256, 256); c = v.Context(s); o = Oracle(256, 256, 4)
    for g in (c, o):
        g.set_source_rgba(0, 0, 1, 0.7); g.set_line_width(7); g.set_line_join(1); g.set_line_cap(1); g.set_dash([10, 6])
        g.move_to(10, 10); g.line_to(100, 20); g.curve_to(150, 50, 120, 150, 50, 100); g.close_path(); g.stroke()
    c.flush()
    a, b = s.pixels(), o.pixels()
    d = np.abs(a.astype(int) - b.astype(int)).max(axis=2)
    return int((d > 0).sum()), int(d.max()), int((b[..., 3] > 0).sum())
step("stroke_pixels", stroke_pixels)

def tiger():
    w, h, shapes = scenes.load_nsvg(os.path.join(os.path.dirname(__file__), "golden", "tiger.nsvg.bin"))
    s = v.Surface(dev, 1024, 1024); c = v.Context(s)
    dev.set_profiling(True)
    scenes.render_nsvg(c, shapes); c.flush()
    st = dev.last_stats()
    img = s.write_to_memory()
    s.write_to_png("gpurun_out/tiger_gpu.png")
    return st, int((img[..., 3] > 0).sum())
step("tiger", tiger)
print("launches", v.lib().vkvg_b200_launch_count())
