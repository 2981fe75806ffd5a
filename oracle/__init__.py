"""TEST INFRASTRUCTURE ONLY — ctypes bindings for the CPU oracle (oracle/liboracle.so) and, when it has
been built, the reference's own tessellation code (oracle/_ref/libvkvg_ref.so).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this
package.  The product package (vkvg_b200) never does.

`Oracle` and `Ref` expose the same drawing vocabulary (move_to, line_to, curve_to, arc, close_path,
set_line_width/cap/join/dash, set_fill_rule, set_source_*, fill, stroke, paint ...) so a scene written
once as a Python function can be replayed on the oracle, on the reference and on the CUDA library.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_f, _u, _i, _p = C.c_float, C.c_uint32, C.c_int, C.c_void_p


def build(ref=True):
    """(Re)build liboracle.so and, if /root/reference is present, _ref/libvkvg_ref.so."""
    subprocess.run(["make", "-s", "-C", _HERE], check=True)
    if ref and os.path.isdir(os.environ.get("VKVG_REF", "/root/reference")):
        subprocess.run(["make", "-s", "-C", _HERE, "ref", "nsvg"], check=True)


def _load(path):
    if not os.path.exists(path):
        raise FileNotFoundError(path + " (run `make -C oracle` / `make -C oracle ref`)")
    return C.CDLL(path)


_PATH_API = {
    "new_path": [], "new_sub_path": [], "close_path": [],
    "move_to": [_f] * 2, "line_to": [_f] * 2, "rel_move_to": [_f] * 2, "rel_line_to": [_f] * 2,
    "curve_to": [_f] * 6, "rel_curve_to": [_f] * 6, "quadratic_to": [_f] * 4,
    "arc": [_f] * 5, "arc_negative": [_f] * 5, "rectangle": [_f] * 4,
    "set_line_width": [_f], "set_miter_limit": [_f], "set_line_cap": [_i], "set_line_join": [_i],
    "set_fill_rule": [_i], "set_opacity": [_f], "set_operator": [_i], "set_source_rgba": [_f] * 4, "set_source_color": [_u],
    "translate": [_f] * 2, "scale": [_f] * 2, "rotate": [_f], "identity_matrix": [],
    "fill": [], "fill_preserve": [], "stroke": [], "stroke_preserve": [], "paint": [],
    "ellipse": [_f] * 5, "rounded_rectangle": [_f] * 5, "rounded_rectangle2": [_f] * 6, "rel_quadratic_to": [_f] * 4,
    "elliptic_arc_to": [_f, _f, C.c_bool, C.c_bool, _f, _f, _f], "rel_elliptic_arc_to": [_f, _f, C.c_bool, C.c_bool, _f, _f, _f],
    "clip": [], "clip_preserve": [], "reset_clip": [], "save": [], "restore": [],
}


class _Base:
    """Common call surface; subclasses set self._lib, self._ctx and self._prefix."""

    def __getattr__(self, name):
        if name in _PATH_API:
            fn = getattr(self._lib, self._prefix + name)
            fn.argtypes = [_p] + _PATH_API[name]
            fn.restype = None
            ctx = self._ctx
            return lambda *a: fn(ctx, *a)
        raise AttributeError(name)

    def polyline(self, xy):
        """move_to(xy[0]) then line_to for the rest (the two entry points are bound once: a million-point line is a million calls)"""
        xy = np.asarray(xy, np.float32).reshape(-1, 2).tolist()
        mv, ln = self.move_to, self.line_to
        mv(*xy[0])
        for q in xy[1:]:
            ln(q[0], q[1])

    def set_dash(self, dashes, offset=0.0):
        arr = (C.c_float * len(dashes))(*dashes)
        fn = getattr(self._lib, self._prefix + "set_dash")
        fn.argtypes = [_p, C.POINTER(C.c_float), _u, _f]
        fn(self._ctx, arr, len(dashes), offset)


class Oracle(_Base):
    _prefix = "ovk_"
    _libh = None

    @classmethod
    def lib(cls):
        if cls._libh is None:
            L = _load(os.path.join(_HERE, "liboracle.so"))
            L.ovk_create.restype = _p
            L.ovk_create.argtypes = [_u, _u, _u]
            L.ovk_create_window.restype = _p
            L.ovk_create_window.argtypes = [_u] * 7
            L.ovk_pixels.restype = _p
            L.ovk_sample_pixels.restype = _p
            L.ovk_last_coverage.restype = _p
            for n in ("ovk_pixels", "ovk_sample_pixels", "ovk_last_coverage", "ovk_destroy", "ovk_clear", "ovk_status"):
                getattr(L, n).argtypes = [_p]
            L.ovk_flatten_cubic.restype = _u
            L.ovk_flatten_cubic.argtypes = [_f] * 9 + [_p, _u]
            L.ovk_winding_brute.argtypes = [_p, C.c_uint64, _u, _u, _u, _p]
            L.ovk_transform_snap.argtypes = [_p, _u, _u, _p, C.c_uint64, _p]
            L.ovk_raster_ref_drawlist.argtypes = [_p, _p, _u, _p]
            L.ovk_set_source_linear.argtypes = [_p] + [_f] * 4 + [_p, _u]
            L.ovk_set_source_radial.argtypes = [_p] + [_f] * 6 + [_p, _u]
            L.ovk_set_matrix.argtypes = [_p, _p]
            L.ovk_get_matrix.argtypes = [_p, _p]
            L.ovk_write_to_memory.argtypes = [_p, _p]
            L.ovk_set_capture_coverage.argtypes = [_p, _i]
            L.ovk_set_coverage_mode.argtypes = [_p, _i]
            L.ovk_last_area.restype = _p
            L.ovk_last_area.argtypes = [_p]
            L.ovk_area_brute.argtypes = [_p, C.c_uint64, _u, _u, _p]
            for n in ("ovk_path_points", "ovk_path_table", "ovk_last_vertices", "ovk_last_indices"):
                getattr(L, n).restype = _u
                getattr(L, n).argtypes = [_p, C.POINTER(_p)]
            cls._libh = L
        return cls._libh

    def __init__(self, width, height, samples=4, analytic=False, window=None):
        """window=(x0, y0, w, h): only that part of the logical width x height surface is stored and rasterised (it holds
        exactly the pixels the same region of the whole surface would); pixels() etc. then return window-sized arrays."""
        self._lib = self.lib()
        if analytic:
            samples = 1   # analytic-coverage mode keeps one colour per pixel
        self.full_width, self.full_height = width, height
        if window is not None:
            assert not analytic, "the analytic-coverage restatement needs the whole surface"
            x0, y0, w, h = (int(t) for t in window)
            self.window = (x0, y0, w, h)
            self._ctx = self._lib.ovk_create_window(width, height, samples, x0, y0, w, h)
            width, height = w, h
        else:
            self.window = (0, 0, width, height)
            self._ctx = self._lib.ovk_create(width, height, samples)
        self.width, self.height, self.samples = width, height, samples   # of the stored pixels
        if not self._ctx:
            raise ValueError("unsupported sample count %r or window %r" % (samples, window))
        if analytic:
            self._lib.ovk_set_coverage_mode(self._ctx, 1)

    def close(self):
        if self._ctx:
            self._lib.ovk_destroy(self._ctx)
            self._ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def clear(self):
        self._lib.ovk_clear_ctx.argtypes = [_p]
        self._lib.ovk_clear_ctx(self._ctx)   # vkvg_clear: surface + stencil wiped, clip-state bookkeeping updated

    def status(self):
        return self._lib.ovk_status(self._ctx)

    def set_matrix(self, m):
        a = np.asarray(m, np.float32)
        self._lib.ovk_set_matrix(self._ctx, a.ctypes.data)

    def get_matrix(self):
        a = np.zeros(6, np.float32)
        self._lib.ovk_get_matrix(self._ctx, a.ctypes.data)
        return a

    def set_source_linear(self, x0, y0, x1, y1, stops):
        s = np.asarray(stops, np.float32).reshape(-1, 5)
        self._lib.ovk_set_source_linear(self._ctx, x0, y0, x1, y1, s.ctypes.data, len(s))

    def set_source_radial(self, cx0, cy0, r0, cx1, cy1, r1, stops):
        s = np.asarray(stops, np.float32).reshape(-1, 5)
        self._lib.ovk_set_source_radial(self._ctx, cx0, cy0, r0, cx1, cy1, r1, s.ctypes.data, len(s))

    def set_source_surface(self, src, x=0.0, y=0.0, extend=None, filter=None, matrix=None):
        """src: (H, W, 4) uint8 premultiplied pixels or another Oracle; same call shapes as vkvg_b200.Context.set_source_surface"""
        img = src.pixels() if isinstance(src, Oracle) else np.ascontiguousarray(src, np.uint8)
        self._src_keep = img   # the oracle samples the caller's buffer
        L = self._lib
        L.ovk_set_source_surface.argtypes = [_p, _p, _u, _u, _f, _f, _i, _i, _p, _i]
        h, w = img.shape[:2]
        if extend is None and filter is None and matrix is None:
            L.ovk_set_source_surface(self._ctx, img.ctypes.data, w, h, x, y, 0, 0, None, 0)
            return
        L.ovk_set_source_surface(self._ctx, img.ctypes.data, w, h, x, y, 0, 0, None, 0)
        m = None if matrix is None else np.asarray(matrix, np.float32)
        self._mat_keep = m
        linear = 1 if filter in (2, 4) else 0     # VKVG_FILTER_BEST, VKVG_FILTER_BILINEAR
        L.ovk_set_source_surface(self._ctx, img.ctypes.data, w, h, 0.0, 0.0, int(extend or 0), linear, None if m is None else m.ctypes.data, 1)

    def source_push(self):
        out = np.zeros(10, np.float32)
        self._lib.ovk_get_source_push.argtypes = [_p, _p]
        self._lib.ovk_get_source_push(self._ctx, out.ctypes.data)
        return out

    def _arr(self, fn, dtype, per):
        ptr = _p()
        n = fn(self._ctx, C.byref(ptr))
        if n == 0:
            return np.zeros((0, per) if per > 1 else (0,), dtype)
        buf = (C.c_byte * (n * per * np.dtype(dtype).itemsize)).from_address(ptr.value)
        a = np.frombuffer(buf, dtype=dtype).copy()
        return a.reshape(n, per) if per > 1 else a

    def path_points(self):
        return self._arr(self._lib.ovk_path_points, np.float32, 2)

    def path_table(self):
        return self._arr(self._lib.ovk_path_table, np.uint32, 1)

    def last_vertices(self):
        return self._arr(self._lib.ovk_last_vertices, np.float32, 2)

    def last_indices(self):
        return self._arr(self._lib.ovk_last_indices, np.uint32, 1)

    def pixels(self):
        ptr = self._lib.ovk_pixels(self._ctx)
        buf = (C.c_ubyte * (self.width * self.height * 4)).from_address(ptr)
        return np.frombuffer(buf, np.uint8).reshape(self.height, self.width, 4).copy()

    def sample_pixels(self):
        ptr = self._lib.ovk_sample_pixels(self._ctx)
        buf = (C.c_ubyte * (self.width * self.height * self.samples * 4)).from_address(ptr)
        return np.frombuffer(buf, np.uint8).reshape(self.height, self.width, self.samples, 4).copy()

    def write_to_memory(self):
        out = np.zeros((self.height, self.width, 4), np.uint8)
        self._lib.ovk_write_to_memory(self._ctx, out.ctypes.data)
        return out

    def capture_coverage(self, on=True):
        self._lib.ovk_set_capture_coverage(self._ctx, int(on))

    def last_coverage(self):
        ptr = self._lib.ovk_last_coverage(self._ctx)
        buf = (C.c_int32 * (self.width * self.height * self.samples)).from_address(ptr)
        return np.frombuffer(buf, np.int32).reshape(self.height, self.width, self.samples).copy()

    def last_area(self):
        """analytic mode: integral of the winding number over each pixel for the last draw, (H, W) float64."""
        ptr = self._lib.ovk_last_area(self._ctx)
        buf = (C.c_double * (self.width * self.height)).from_address(ptr)
        return np.frombuffer(buf, np.float64).reshape(self.height, self.width).copy()

    def raster_ref_drawlist(self, draws_ptr, n, blob_ptr):
        self._lib.ovk_raster_ref_drawlist(self._ctx, draws_ptr, n, blob_ptr)


def flatten_cubic(p, tol):
    """p: 8 floats (x0,y0..x3,y3). Returns (n,2) float32 points (excluding p0, including the end point)."""
    L = Oracle.lib()
    args = [float(v) for v in p] + [float(tol)]
    n = L.ovk_flatten_cubic(*args, None, 0)
    out = np.zeros((n, 2), np.float32)
    L.ovk_flatten_cubic(*args, out.ctypes.data, n)
    return out


def area_brute(edges, width, height):
    """analytic mode: integral of the winding number of directed 24.8 edges over each pixel square, (H, W) float64."""
    e = np.ascontiguousarray(edges, np.int32).reshape(-1, 4)
    out = np.zeros((height, width), np.float64)
    Oracle.lib().ovk_area_brute(e.ctypes.data, len(e), width, height, out.ctypes.data)
    return out


def winding_brute(edges, width, height, samples):
    e = np.ascontiguousarray(edges, np.int32).reshape(-1, 4)
    out = np.zeros((height, width, samples), np.int32)
    Oracle.lib().ovk_winding_brute(e.ctypes.data, len(e), width, height, samples, out.ctypes.data)
    return out


def transform_snap(m, width, height, xy):
    m = np.asarray(m, np.float32)
    xy = np.ascontiguousarray(xy, np.float32).reshape(-1, 2)
    out = np.zeros((len(xy), 2), np.int32)
    Oracle.lib().ovk_transform_snap(m.ctypes.data, width, height, xy.ctypes.data, len(xy), out.ctypes.data)
    return out


def sample_positions(samples):
    a = np.zeros((samples, 2), np.int32)
    L = Oracle.lib()
    L.ovk_sample_positions.argtypes = [_u, _p]
    if not L.ovk_sample_positions(samples, a.ctypes.data):
        raise ValueError(samples)
    return a


# --------------------------------------------------------------------------------------------------
class _RefDraw(C.Structure):
    _fields_ = [("kind", C.c_int32), ("pipeline", _u), ("cmpMask", _u), ("ref", _u), ("writeMask", _u),
                ("sc_x", C.c_int32), ("sc_y", C.c_int32), ("sc_w", _u), ("sc_h", _u), ("count", _u), ("first", _u),
                ("vertexOffset", C.c_int32), ("vbo", _u), ("ibo", _u), ("ubo", _u), ("push", C.c_ubyte * 80)]


class _RefList(C.Structure):
    _fields_ = [("draws", C.POINTER(_RefDraw)), ("n_draws", _u), ("cap_draws", _u), ("blob", _p),
                ("n_blob", C.c_uint64), ("cap_blob", C.c_uint64)]


def ref_available():
    return os.path.exists(os.path.join(_HERE, "_ref", "libvkvg_ref.so"))


class Ref(_Base):
    """The reference's own vkvg_* entry points (unmodified object code) on a fake device that records draws."""
    _prefix = "vkvg_"
    _libh = None

    @classmethod
    def lib(cls):
        if cls._libh is None:
            L = _load(os.path.join(_HERE, "_ref", "libvkvg_ref.so"))
            L.ref_device_create.restype = _p
            L.ref_device_create.argtypes = [_u]
            L.ref_surface_create.restype = _p
            L.ref_surface_create.argtypes = [_p, _u, _u]
            L.vkvg_create.restype = _p
            L.vkvg_create.argtypes = [_p]
            for n in ("vkvg_destroy", "vkvg_flush", "vkvg_clear", "vkvg_status", "vkvg_surface_destroy", "ref_device_destroy",
                      "ref_ctx_finish_path", "vkvg_save", "vkvg_restore"):
                getattr(L, n).argtypes = [_p]
            L.ref_drawlist.restype = C.POINTER(_RefList)
            L.ref_set_recording.argtypes = [_i]
            for n in ("ref_ctx_points", "ref_ctx_pathes", "ref_ctx_vertices", "ref_ctx_indices"):
                getattr(L, n).restype = _u
                getattr(L, n).argtypes = [_p, C.POINTER(_p)]
            L.vkvg_pattern_create_linear.restype = _p
            L.vkvg_pattern_create_linear.argtypes = [_f] * 4
            L.vkvg_pattern_create_radial.restype = _p
            L.vkvg_pattern_create_radial.argtypes = [_f] * 6
            L.vkvg_pattern_add_color_stop.argtypes = [_p] + [_f] * 5
            L.vkvg_pattern_destroy.argtypes = [_p]
            L.vkvg_set_source.argtypes = [_p, _p]
            L.vkvg_set_matrix.argtypes = [_p, _p]
            cls._libh = L
        return cls._libh

    def __init__(self, width, height, samples=4, record=True):
        self._lib = self.lib()
        self.width, self.height, self.samples = width, height, samples
        self._lib.ref_set_recording(int(record))
        self._lib.ref_drawlist_reset()
        self._dev = self._lib.ref_device_create(samples)
        self._surf = self._lib.ref_surface_create(self._dev, width, height)
        self._ctx = self._lib.vkvg_create(self._surf)

    def close(self):
        if self._ctx:
            self._lib.vkvg_destroy(self._ctx)
            self._lib.vkvg_surface_destroy(self._surf)
            self._lib.ref_device_destroy(self._dev)
            self._ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def clear(self):
        self._lib.vkvg_clear(self._ctx)

    def flush(self):
        self._lib.vkvg_flush(self._ctx)

    def save(self):
        self._lib.vkvg_save(self._ctx)

    def restore(self):
        self._lib.vkvg_restore(self._ctx)

    def status(self):
        return self._lib.vkvg_status(self._ctx)

    def set_matrix(self, m):
        a = np.asarray(m, np.float32)
        self._lib.vkvg_set_matrix(self._ctx, a.ctypes.data)

    def _grad(self, pat, stops):
        for s in np.asarray(stops, np.float32).reshape(-1, 5):
            self._lib.vkvg_pattern_add_color_stop(pat, *[float(v) for v in s])
        self._lib.vkvg_set_source(self._ctx, pat)
        self._lib.vkvg_pattern_destroy(pat)

    def set_source_linear(self, x0, y0, x1, y1, stops):
        self._grad(self._lib.vkvg_pattern_create_linear(x0, y0, x1, y1), stops)

    def set_source_radial(self, cx0, cy0, r0, cx1, cy1, r1, stops):
        self._grad(self._lib.vkvg_pattern_create_radial(cx0, cy0, r0, cx1, cy1, r1), stops)

    def _arr(self, fn, dtype, per, stride=None):
        ptr = _p()
        n = fn(self._ctx, C.byref(ptr))
        item = np.dtype(dtype).itemsize
        stride = stride or per * item
        if n == 0:
            return np.zeros((0, per) if per > 1 else (0,), dtype)
        buf = (C.c_byte * (n * stride)).from_address(ptr.value)
        raw = np.frombuffer(buf, np.uint8).reshape(n, stride)[:, :per * item].copy()
        a = raw.view(dtype)
        return a.reshape(n, per) if per > 1 else a.reshape(n)

    def path_points(self):
        self._lib.ref_ctx_finish_path(self._ctx)
        return self._arr(self._lib.ref_ctx_points, np.float32, 2)

    def path_table(self):
        self._lib.ref_ctx_finish_path(self._ctx)
        return self._arr(self._lib.ref_ctx_pathes, np.uint32, 1)

    def cached_vertices(self):
        """positions of the vertices currently in the context's vertex cache (24-byte Vertex stride)."""
        return self._arr(self._lib.ref_ctx_vertices, np.float32, 2, stride=24)

    def cached_indices(self):
        return self._arr(self._lib.ref_ctx_indices, np.uint32, 1)

    def drawlist(self):
        return self._lib.ref_drawlist().contents

    def render_with(self, oracle):
        """flush, then rasterise everything recorded so far with the oracle's Vulkan restatement."""
        self.flush()
        dl = self.drawlist()
        oracle.raster_ref_drawlist(C.cast(dl.draws, _p), dl.n_draws, dl.blob)
        self._lib.ref_drawlist_reset()
