"""GPU: vkvg_b200_submit - the packed command stream decoded by kernels (vkvg_b200/csrc/decode.cu) - against the same stream decoded on
the host call by call: identical pixels, identical context state afterwards; streams outside what the device decodes (ops outside the
subset, points the reference would drop, close_path calls it would ignore) fall back to the host decoder and give the same result too."""
import numpy as np
import pytest

import bench
import vkvg_b200 as v
from tests import scenes

pytestmark = pytest.mark.gpu


def _render(dev, size, cs, mode, batch=None, pre=None, post=None):
    L = v.lib()
    L.vkvg_b200_set_submit_decoder(mode)
    try:
        s = v.Surface(dev, size, size, batch=batch)
        c = v.Context(s)
        if pre:
            pre(c)
        d0, h0 = v.submit_counts()
        assert c.submit(*cs.arrays2()) == 0
        d1, h1 = v.submit_counts()
        if post:
            post(c)
            c.flush()
        img = s.pixels()
        c.close()
        s.close()
    finally:
        L.vkvg_b200_set_submit_decoder(0)
    return img, (d1 - d0, h1 - h0)


@pytest.mark.parametrize("workload,n_limit,rule", [("c2", 3000, "nz"), ("c2", 3000, "eo"), ("c3", 30000, "nz"), ("c4", 400, "nz"), ("c1", None, "eo"), ("c5a", 1500, "nz")])
def test_device_decode_equals_host_decode(dev4, workload, n_limit, rule):
    emit, _, _ = bench.build_scene(workload, 1, rule, n_limit=n_limit)
    cs = v.CommandStream()
    emit(cs)
    size = 1024 if workload in ("c1", "c2") else 2048
    a, where_a = _render(dev4, size, cs, 0)
    b, where_b = _render(dev4, size, cs, 1)
    assert where_a == (1, 0), "the device decoder declined a regular stream: %r" % (where_a,)
    assert where_b == (0, 1)
    assert a[..., 3].any()
    assert np.array_equal(a, b)


def test_batch_of_canvases_through_submit(dev4):
    emit, _, _ = bench.build_scene("c5b", 1, "eo")
    cs = v.CommandStream()
    emit(cs)
    a, where = _render(dev4, 1024, cs, 0, batch=bench.C5B_BATCH)
    b, _ = _render(dev4, 1024, cs, 1, batch=bench.C5B_BATCH)
    assert where == (1, 0)
    assert np.array_equal(a, b)


def test_state_before_and_after_the_stream(dev4):
    """the stream starts from the context's state (colour, line state, dashes, CTM set by ordinary calls) and leaves its own behind"""
    def pre(c):
        c.set_source_rgba(0.2, 0.9, 0.3, 0.7)
        c.set_line_width(6.0)
        c.set_line_join(1)
        c.set_dash([7.0, 3.0], 2.0)
        c.translate(30.0, 20.0)

    def post(c):   # drawn with whatever the stream left: source, width, fill rule, CTM
        c.move_to(10.0, 200.0)
        c.line_to(300.0, 180.0)
        c.line_to(180.0, 300.0)
        c.stroke()
        c.rectangle(50.0, 50.0, 80.0, 60.0)
        c.fill()

    cs = v.CommandStream()
    cs.move_to(20.0, 20.0); cs.line_to(200.0, 40.0); cs.line_to(120.0, 220.0); cs.stroke()          # with the state set by pre()
    cs.set_source_rgba(0.9, 0.1, 0.1, 0.5); cs.set_line_width(3.0); cs.set_fill_rule(0)
    cs.translate(100.0, 5.0)
    cs.move_to(0.0, 0.0); cs.line_to(150.0, 10.0); cs.line_to(20.0, 140.0); cs.line_to(140.0, 150.0); cs.close_path(); cs.fill_preserve(); cs.stroke()
    cs.set_source_linear(0.0, 0.0, 200.0, 200.0, [(0, 1, 0, 0, 1), (1, 0, 0, 1, 0.5)])
    cs.polyline(np.array([[30, 160], [220, 170], [60, 260], [200, 280]], np.float32)); cs.close_path(); cs.fill()
    cs.set_source_rgba(0.1, 0.1, 0.8, 0.9)
    a, where = _render(dev4, 512, cs, 0, pre=pre, post=post)
    b, _ = _render(dev4, 512, cs, 1, pre=pre, post=post)
    assert where == (1, 0)
    assert np.array_equal(a, b)


@pytest.mark.parametrize("case", ["arc", "nan", "duplicate", "short_close", "open_end", "lone_move", "clip"])
def test_irregular_streams_fall_back_to_the_host(dev4, case):
    cs = v.CommandStream()
    cs.set_source_rgba(0.3, 0.5, 0.9, 0.8)
    cs.move_to(20.0, 20.0); cs.line_to(200.0, 30.0); cs.line_to(100.0, 200.0); cs.close_path(); cs.fill()
    if case == "arc":
        cs.arc(120.0, 120.0, 60.0, 0.0, 4.0); cs.fill()
    elif case == "nan":
        cs.move_to(30.0, 30.0); cs.line_to(float("nan"), 50.0); cs.line_to(180.0, 90.0); cs.line_to(60.0, 170.0); cs.fill()
    elif case == "duplicate":
        cs.move_to(30.0, 30.0); cs.line_to(150.0, 50.0); cs.line_to(150.0, 50.0); cs.line_to(60.0, 170.0); cs.fill()
    elif case == "short_close":
        cs.move_to(30.0, 30.0); cs.line_to(150.0, 50.0); cs.close_path(); cs.line_to(60.0, 170.0); cs.line_to(10.0, 100.0); cs.fill()
    elif case == "open_end":
        cs.move_to(30.0, 30.0); cs.line_to(150.0, 50.0); cs.line_to(60.0, 170.0)
    elif case == "lone_move":
        cs.move_to(30.0, 30.0); cs.move_to(40.0, 40.0); cs.line_to(150.0, 50.0); cs.line_to(60.0, 170.0); cs.fill()
    elif case == "clip":
        cs.move_to(0.0, 0.0); cs.line_to(128.0, 0.0); cs.line_to(128.0, 256.0); cs.line_to(0.0, 256.0); cs.clip()
        cs.move_to(30.0, 30.0); cs.line_to(250.0, 50.0); cs.line_to(60.0, 170.0); cs.fill()

    def post(c):   # (an open path left by the stream is filled by the next call)
        c.fill()
    a, where = _render(dev4, 256, cs, 0, post=post)
    b, _ = _render(dev4, 256, cs, 1, post=post)
    assert where == (0, 1), (case, where)
    assert np.array_equal(a, b)


@pytest.mark.parametrize("size", [512, 2048])
def test_read_back_target_gets_the_same_image(dev4, size):
    """vkvg_b200_surface_set_readback: the fine pass runs in bands of tile rows, each copied out while the next renders; the registered
    host buffer then holds exactly what an ordinary read-back returns - over several frames (graph replays included) and a clear."""
    import ctypes as C
    emit, _, _ = bench.build_scene("c2", 3, "nz", n_limit=2500)
    cs = v.CommandStream()
    emit(cs)
    cmds, args = cs.arrays2()
    L = v.lib()
    s = v.Surface(dev4, size, size)
    c = v.Context(s)
    assert c.submit(cmds, args) == 0
    want = s.pixels()
    target = np.zeros((size, size, 4), np.uint8)
    assert L.vkvg_b200_surface_set_readback(s.h, target.ctypes.data) == 0
    for frame in range(5):
        target[:] = 7
        c.clear()
        assert c.submit(cmds, args) == 0
        assert L.vkvg_b200_surface_read_premultiplied(s.h, target.ctypes.data) == 0
        assert np.array_equal(target, want), frame
    # drawing on top without a clear: the per-sample plane of the banded frames is what the next flush continues from
    c.set_source_rgba(0.1, 0.8, 0.2, 0.5)
    c.rectangle(10.0, 10.0, size - 60.0, size / 2.0)
    c.fill()
    c.flush()
    assert L.vkvg_b200_surface_read_premultiplied(s.h, target.ctypes.data) == 0
    assert L.vkvg_b200_surface_set_readback(s.h, None) == 0
    s2 = v.Surface(dev4, size, size)
    c2 = v.Context(s2)
    assert c2.submit(cmds, args) == 0
    c2.set_source_rgba(0.1, 0.8, 0.2, 0.5)
    c2.rectangle(10.0, 10.0, size - 60.0, size / 2.0)
    c2.fill()
    c2.flush()
    assert np.array_equal(target, s2.pixels())
    assert np.array_equal(s.pixels(), s2.pixels())


def test_gradient_source_in_force_when_the_stream_starts(dev4):
    """a stream may begin with draws that use the source the previous stream (or ordinary calls) left behind - a gradient included"""
    first = v.CommandStream()
    first.set_source_radial(120.0, 110.0, 10.0, 130.0, 100.0, 150.0, [(0, 1, 1, 0, 1), (0.6, 0, 1, 1, 0.7), (1, 0, 0, 1, 0.4)])
    first.move_to(10.0, 10.0); first.line_to(200.0, 30.0); first.line_to(60.0, 180.0); first.close_path(); first.fill()
    second = v.CommandStream()
    second.move_to(60.0, 60.0); second.line_to(250.0, 90.0); second.line_to(120.0, 240.0); second.close_path(); second.fill()   # with the radial gradient
    second.set_source_rgba(0.9, 0.2, 0.1, 0.5)
    second.move_to(5.0, 200.0); second.line_to(100.0, 210.0); second.line_to(40.0, 250.0); second.close_path(); second.fill()
    imgs = []
    for mode in (0, 1):
        v.lib().vkvg_b200_set_submit_decoder(mode)
        try:
            s = v.Surface(dev4, 256, 256)
            c = v.Context(s)
            d0 = v.submit_counts()
            assert c.submit(*first.arrays2()) == 0 and c.submit(*second.arrays2()) == 0
            d1 = v.submit_counts()
            assert (d1[0] - d0[0], d1[1] - d0[1]) == ((2, 0) if mode == 0 else (0, 2))
            imgs.append(s.pixels())
            c.close()
            s.close()
        finally:
            v.lib().vkvg_b200_set_submit_decoder(0)
    assert imgs[0][..., 3].any() and np.array_equal(imgs[0], imgs[1])
