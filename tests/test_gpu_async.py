"""GPU: the execution model around the kernels — device-side counts with capacity overflow + replay, asynchronous flushes,
and CUDA graph replay of structurally identical frames — never changes a pixel."""
import numpy as np
import pytest

import vkvg_b200 as v
from tests import scenes
from tests.golden import make_golden as mg

pytestmark = pytest.mark.gpu


def _frame(g, k, size):
    """same structure for every k (same number of elements, sub-paths and draws), different coordinates and colours"""
    r = scenes.SplitMix64(4242)
    for i in range(40):
        g.set_source_rgba(r.u(), r.u(), r.u(), 0.4 + 0.5 * r.u())
        x, y = r.uniform(5, size - 40), r.uniform(5, size - 40)
        g.move_to(x + k, y)
        g.curve_to(x + 30, y - 10 + k, x + 35 + k, y + 30, x + 5, y + 35)
        g.line_to(x - 8, y + 12 + 0.5 * k)
        g.close_path()
        if i % 3:
            g.fill()
        else:
            g.set_line_width(2.0 + i % 4)
            g.stroke()


def test_graph_replay_of_repeated_frames_is_bit_exact(oracle_lib):
    dev = v.Device(4)
    surf = v.Surface(dev, 160, 160)
    ctx = v.Context(surf)
    r0 = dev.graph_replays()
    for k in range(7):
        ctx.clear()
        _frame(ctx, k, 160)
        ctx.flush()
        o = oracle_lib.Oracle(160, 160, 4)
        _frame(o, k, 160)
        got, ref = surf.pixels(), o.pixels()
        assert np.array_equal(got, ref), (k, int((got != ref).any(axis=2).sum()))
        o.close()
    assert dev.graph_replays() - r0 >= 3   # frames 0-1 launch directly, frame 2 captures, later ones replay
    # a frame of a different structure falls back to plain launches and is still right
    ctx.clear()
    mg.pixel_scene(ctx, "mixed", 1, 160)
    ctx.flush()
    o = oracle_lib.Oracle(160, 160, 4)
    mg.pixel_scene(o, "mixed", 1, 160)
    assert np.array_equal(surf.pixels(), o.pixels())
    dev.close()


def test_graphs_off_gives_the_same_pixels(oracle_lib):
    dev = v.Device(4)
    dev.set_graphs(False)
    surf = v.Surface(dev, 160, 160)
    ctx = v.Context(surf)
    for k in range(4):
        ctx.clear()
        _frame(ctx, k, 160)
        ctx.flush()
    o = oracle_lib.Oracle(160, 160, 4)
    _frame(o, 3, 160)
    assert np.array_equal(surf.pixels(), o.pixels())
    assert dev.graph_replays() == 0
    dev.close()


def test_capacity_overflow_replays_the_batch(oracle_lib):
    """a fresh device sizes its intermediates from guesses; a scene that needs far more (long dashed stroke: ~40 vertices per
    input point; thousands of path-tiles) overflows them, nothing reaches the surface, and the replay with room is exact"""
    dev = v.Device(4)
    surf = v.Surface(dev, 256, 256)
    ctx = v.Context(surf)
    o = oracle_lib.Oracle(256, 256, 4)
    pts = scenes.polyline_c3(400, 256, 5)
    for g in (ctx, o):
        g.set_source_rgba(0.2, 0.4, 0.9, 0.6)
        g.paint()
        g.set_source_rgba(0.9, 0.3, 0.1, 0.8)
        g.set_line_width(3.0)
        g.set_line_join(1)
        g.set_line_cap(1)
        g.set_dash([2.0, 1.5], 0.0)
        g.move_to(float(pts[0, 0]), float(pts[0, 1]))
        for p in pts[1:]:
            g.line_to(float(p[0]), float(p[1]))
        g.stroke()
        for i in range(300):
            g.set_source_rgba(0.1, 0.8, 0.3, 0.3)
            g.rectangle(3.0 + (i % 20) * 12, 2.0 + (i // 20) * 16, 30.0, 40.0)
            g.fill()
    ctx.flush()
    got, ref = surf.pixels(), o.pixels()
    assert np.array_equal(got, ref), int((got != ref).any(axis=2).sum())
    dev.close()


def test_flush_is_asynchronous_but_ordered(oracle_lib):
    """several flushes queued back to back without reading anything in between, then one read"""
    dev = v.Device(4)
    surf = v.Surface(dev, 128, 128)
    ctx = v.Context(surf)
    o = oracle_lib.Oracle(128, 128, 4)
    for name in ("eo", "stroke_alpha", "grad_linear", "stroke_dash", "grad_radial"):
        mg.pixel_scene(ctx, name, 2)
        ctx.flush()
        mg.pixel_scene(o, name, 2)
    assert np.array_equal(surf.pixels(), o.pixels())
    dev.close()


def test_batch_of_canvases_equals_separate_surfaces(oracle_lib):
    """C5b building block: n canvases in one surface, one flush; every band equals the canvas rendered on its own"""
    from tests.golden import make_golden2 as mg2
    dev = v.Device(4)
    names = ["mixed", "grad_radial", "stroke_dash", "paint", "eo", "grad_linear"]
    surf = v.Surface(dev, 128, 128, batch=len(names) + 2)
    ctx = v.Context(surf)
    refs = []
    for i, name in enumerate(names):
        ctx.set_canvas(i)
        ctx.identity_matrix()
        ctx.translate(0.25 * i, -0.5 * i)      # per-canvas jitter as in the C5b configuration
        mg.pixel_scene(ctx, name, 1)
        o = oracle_lib.Oracle(128, 128, 4)
        o.translate(0.25 * i, -0.5 * i)
        mg.pixel_scene(o, name, 1)
        refs.append(o.pixels())
        o.close()
    # canvas 6: clipping stays inside its canvas; canvas 7 stays empty
    ctx.set_canvas(6)
    ctx.identity_matrix()
    mg2.clip_scene(ctx, "nested", 0)
    o = oracle_lib.Oracle(128, 128, 4)
    mg2.clip_scene(o, "nested", 0)
    refs.append(o.pixels())
    refs.append(np.zeros((128, 128, 4), np.uint8))
    ctx.flush()
    got = surf.pixels()
    assert got.shape == (128 * 8, 128, 4)
    for i, ref in enumerate(refs):
        band = got[128 * i:128 * (i + 1)]
        assert np.array_equal(band, ref), (i, int((band != ref).any(axis=2).sum()))
    with pytest.raises(v.VkvgError):
        ctx.set_canvas(8)
    dev.close()


def test_recording_replays_to_the_same_pixels(oracle_lib):
    """vkvg_start_recording / stop_recording / replay (reference include/vkvg.h:1961-1970): calls are stored, not executed, and a
    replay on another context draws what the direct calls draw"""
    import ctypes as C
    from tests.golden import make_golden2 as mg2
    L = v.lib()
    dev = v.Device(4)
    a, b = v.Surface(dev, 128, 128), v.Surface(dev, 128, 128)
    ca, cb = v.Context(a), v.Context(b)
    ca.start_recording()
    for scene in (lambda g: mg.pixel_scene(g, "mixed", 1), lambda g: mg2.clip_scene(g, "save_restore", 0), lambda g: mg2.clip_scene(g, "op_mixed", 1)):
        scene(ca)
    ca.set_dash([4.0, 2.0, 1.0, 2.0], 1.5)
    ca.set_line_width(3.0)
    ca.move_to(10.0, 120.0)
    ca.rel_line_to(100.0, -8.0)
    ca.elliptic_arc_to(60.0, 90.0, True, False, 30.0, 14.0, 0.4)
    ca.stroke()
    rec = ca.stop_recording()
    assert rec
    ca.flush()
    assert not a.pixels().any()                       # nothing was drawn while recording
    n = L.vkvg_recording_get_count(rec)
    assert n > 50
    cmd, off = C.c_uint32(), C.c_void_p()
    codes = []
    for i in range(n):
        L.vkvg_recording_get_command(rec, i, C.byref(cmd), C.byref(off))
        codes.append(cmd.value)
    # the reference's command codes (src/recording/vkvg_record_internal.h): fill, stroke, clip, save, restore, set_operator, set_dash ...
    assert {0x0202, 0x0203, 0x0204, 0x0001, 0x0002, 0x1105, 0x1107, 0x0104, 0x0505, 0x010C, 0x0802} <= set(codes)
    L.vkvg_recording_get_command(rec, n - 1, C.byref(cmd), C.byref(off))
    assert cmd.value == 0x0203                        # VKVG_CMD_STROKE
    L.vkvg_recording_get_command(rec, n, C.byref(cmd), C.byref(off))
    assert cmd.value == 0 and not off.value
    ca.replay_recording(rec)
    ca.flush()
    # the same calls issued directly
    for scene in (lambda g: mg.pixel_scene(g, "mixed", 1), lambda g: mg2.clip_scene(g, "save_restore", 0), lambda g: mg2.clip_scene(g, "op_mixed", 1)):
        scene(cb)
    cb.set_dash([4.0, 2.0, 1.0, 2.0], 1.5)
    cb.set_line_width(3.0)
    cb.move_to(10.0, 120.0)
    cb.rel_line_to(100.0, -8.0)
    cb.elliptic_arc_to(60.0, 90.0, True, False, 30.0, 14.0, 0.4)
    cb.stroke()
    cb.flush()
    pa, pb = a.pixels(), b.pixels()
    assert np.array_equal(pa, pb) and pa.any()
    L.vkvg_recording_destroy(rec)
    ca.start_recording()
    assert ca.stop_recording() is None               # an empty recording is dropped
    dev.close()


def test_contexts_on_different_threads(oracle_lib):
    """the reference's tests/multithreading/multithreaded.c idiom: one device, every thread its own surface and context; then the
    thread results are composited onto one surface as surface paints"""
    import threading
    names = ["mixed", "stroke_alpha", "grad_radial", "eo", "stroke_dash", "grad_linear"]
    dev = v.Device(4)
    surfs = [v.Surface(dev, 128, 128) for _ in names]
    errors = []

    def work(i):
        try:
            c = v.Context(surfs[i])
            for rep in range(3):             # several flushes per thread, interleaved with the other threads' on the one device stream
                if rep:
                    c.clear()
                mg.pixel_scene(c, names[i], 1)
                c.flush()
            c.close()
        except Exception as e:               # pragma: no cover
            errors.append(e)

    threads = [threading.Thread(target=work, args=(i,)) for i in range(len(names))]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors
    refs = []
    for i, name in enumerate(names):
        o = oracle_lib.Oracle(128, 128, 4)
        mg.pixel_scene(o, name, 1)
        refs.append(o.pixels())
        assert np.array_equal(surfs[i].pixels(), refs[i]), name
        o.close()
    # composite: 3 x 2 grid of the thread surfaces
    dst = v.Surface(dev, 384, 256)
    dc = v.Context(dst)
    od = oracle_lib.Oracle(384, 256, 4)
    for g, srcs in ((dc, surfs), (od, refs)):
        for i, src in enumerate(srcs):
            g.set_source_surface(src, 128.0 * (i % 3), 128.0 * (i // 3))
            g.rectangle(128.0 * (i % 3), 128.0 * (i // 3), 128.0, 128.0)
            g.fill()
    dc.flush()
    assert np.array_equal(dst.pixels(), od.pixels())
    dev.close()


def test_recording_right_after_a_flush_does_not_disturb_it():
    """vkvg_flush returns while the GPU still works; the recorder's arrays (copied to the device straight from pinned memory) are
    overwritten by the next frame's calls at once.  A large upload (24 MB) followed immediately by a different frame must give the
    same pixels as the same two frames with a device synchronisation in between."""
    n = 1_000_000
    rng = np.random.default_rng(11)
    a = (rng.random((n, 2)) * 500 + 6).astype(np.float32)
    b = (rng.random((n, 2)) * 500 + 6).astype(np.float32)
    imgs = []
    for wait in (True, False):
        dev = v.Device(4)
        surf = v.Surface(dev, 512, 512)
        ctx = v.Context(surf)
        for k, pts in enumerate((a, b, a[::-1].copy())):
            cs = v.CommandStream()
            cs.set_source_rgba(0.2 + 0.3 * k, 0.9 - 0.3 * k, 0.5, 0.5)
            cs.set_line_width(1.0)
            cs.polyline(pts)
            cs.stroke()
            assert ctx.replay(*cs.arrays()) == 0
            ctx.flush()
            if wait:
                dev.synchronize()
        imgs.append(surf.pixels())
        ctx.close()
        surf.close()
        dev.close()
    assert np.array_equal(imgs[0], imgs[1])
