"""vkvg_b200 — B200-native implementation of vkvg's path-rendering hot path behind the vkvg.h C API.

The product is the shared library ``vkvg_b200/libvkvg_b200.so`` (hand-written sm_100a CUDA + a C host that
exports the reference's ``vkvg_*`` entry points, see ``include/vkvg.h``).  This module is only a thin ctypes
binding of that C ABI for tests and benchmarks written in Python; it contains no rendering logic and has no
CPU fallback: if the library or a CUDA device is missing, construction raises.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("VKVG_B200_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "libvkvg_b200.so")
_f, _u, _i, _p = C.c_float, C.c_uint32, C.c_int, C.c_void_p

# vkvg.h enum values
CAP_BUTT, CAP_ROUND, CAP_SQUARE = 0, 1, 2
JOIN_MITER, JOIN_ROUND, JOIN_BEVEL = 0, 1, 2
FILL_EVEN_ODD, FILL_NON_ZERO = 0, 1
STATUS_SUCCESS = 0

# vkvg_b200.h op codes for the packed command stream
OPS = {name: i + 1 for i, name in enumerate([
    "MOVE_TO", "LINE_TO", "CURVE_TO", "CLOSE_PATH", "NEW_PATH", "ARC", "ARC_NEGATIVE", "RECTANGLE", "FILL", "FILL_PRESERVE",
    "STROKE", "STROKE_PRESERVE", "PAINT", "SET_SOURCE_RGBA", "SET_LINE_WIDTH", "SET_LINE_CAP", "SET_LINE_JOIN",
    "SET_MITER_LIMIT", "SET_FILL_RULE", "SET_DASH", "SET_SOURCE_LINEAR", "SET_SOURCE_RADIAL", "TRANSLATE", "SCALE", "ROTATE",
    "IDENTITY_MATRIX", "SAVE", "RESTORE", "CLEAR", "SET_OPACITY", "POLYLINE", "FLUSH", "SET_CANVAS", "CLIP", "CLIP_PRESERVE", "RESET_CLIP"])}


_FIXED_ARGS = {OPS[k]: n for k, n in dict(
    MOVE_TO=2, LINE_TO=2, CURVE_TO=6, CLOSE_PATH=0, NEW_PATH=0, ARC=5, ARC_NEGATIVE=5, RECTANGLE=4, FILL=0, FILL_PRESERVE=0, STROKE=0, STROKE_PRESERVE=0,
    PAINT=0, SET_SOURCE_RGBA=4, SET_LINE_WIDTH=1, SET_LINE_CAP=1, SET_LINE_JOIN=1, SET_MITER_LIMIT=1, SET_FILL_RULE=1, TRANSLATE=2, SCALE=2, ROTATE=1,
    IDENTITY_MATRIX=0, SAVE=0, RESTORE=0, CLEAR=0, SET_OPACITY=1, FLUSH=0, SET_CANVAS=1, CLIP=0, CLIP_PRESERVE=0, RESET_CLIP=0).items()}


def submit_counts():
    """(streams decoded on the device, streams decoded on the host) by vkvg_b200_submit so far"""
    a, b = C.c_uint64(0), C.c_uint64(0)
    lib().vkvg_b200_submit_counts(C.byref(a), C.byref(b))
    return int(a.value), int(b.value)


class DeviceCreateInfo(C.Structure):
    """vkvg_device_create_info_t (include/vkvg.h)."""
    _fields_ = [("samples", _u), ("deferredResolve", C.c_bool), ("inst", _p), ("phy", _p), ("vkdev", _p), ("qFamIdx", _u),
                ("qIndex", _u), ("threadAware", C.c_bool)]


class Stats(C.Structure):
    """vkvg_b200_stats_t (include/vkvg_b200.h)."""
    _fields_ = [(n, C.c_uint64) for n in ("n_elems", "n_points", "n_fill_edges", "n_stroke_items", "n_verts", "n_inds", "n_edges",
                                          "n_path_tiles", "n_nonempty", "n_tile_edges")] + [("ms_total", _f), ("ms_fine", _f),
                                                                                           ("h2d_bytes", C.c_uint64),
                                                                                           ("ms_stage", _f * 5), ("ms_host_upload", _f)]
    STAGES = ("flatten", "stroke", "edges", "binning", "fine")

    def as_dict(self):
        d = {n: getattr(self, n) for n, _ in self._fields_ if n != "ms_stage"}
        d["ms_stage"] = dict(zip(self.STAGES, [float(x) for x in self.ms_stage]))
        return d


_SIGS = {
    # name: (restype, argtypes)
    "vkvg_device_create": (_p, [_p]), "vkvg_device_destroy": (None, [_p]), "vkvg_device_status": (_i, [_p]),
    "vkvg_device_reference": (_p, [_p]), "vkvg_device_get_reference_count": (_u, [_p]),
    "vkvg_surface_create": (_p, [_p, _u, _u]), "vkvg_surface_destroy": (None, [_p]), "vkvg_surface_status": (_i, [_p]),
    "vkvg_surface_reference": (_p, [_p]), "vkvg_surface_get_reference_count": (_u, [_p]), "vkvg_surface_clear": (None, [_p]),
    "vkvg_surface_get_width": (_u, [_p]), "vkvg_surface_get_height": (_u, [_p]), "vkvg_surface_get_vk_format": (_i, [_p]),
    "vkvg_surface_write_to_png": (_i, [_p, C.c_char_p]), "vkvg_surface_write_to_memory": (_i, [_p, _p]),
    "vkvg_create": (_p, [_p]), "vkvg_destroy": (None, [_p]), "vkvg_status": (_i, [_p]), "vkvg_reference": (_p, [_p]),
    "vkvg_get_reference_count": (_u, [_p]), "vkvg_flush": (None, [_p]), "vkvg_status_to_string": (C.c_char_p, [_i]),
    "vkvg_new_path": (None, [_p]), "vkvg_close_path": (None, [_p]), "vkvg_new_sub_path": (None, [_p]),
    "vkvg_get_current_point": (None, [_p, C.POINTER(_f), C.POINTER(_f)]), "vkvg_has_current_point": (C.c_bool, [_p]),
    "vkvg_line_to": (None, [_p, _f, _f]), "vkvg_rel_line_to": (None, [_p, _f, _f]), "vkvg_move_to": (None, [_p, _f, _f]),
    "vkvg_rel_move_to": (None, [_p, _f, _f]), "vkvg_arc": (None, [_p] + [_f] * 5), "vkvg_arc_negative": (None, [_p] + [_f] * 5),
    "vkvg_curve_to": (None, [_p] + [_f] * 6), "vkvg_rel_curve_to": (None, [_p] + [_f] * 6),
    "vkvg_quadratic_to": (None, [_p] + [_f] * 4), "vkvg_rel_quadratic_to": (None, [_p] + [_f] * 4),
    "vkvg_rectangle": (_i, [_p] + [_f] * 4), "vkvg_rounded_rectangle": (_i, [_p] + [_f] * 5), "vkvg_ellipse": (None, [_p] + [_f] * 5),
    "vkvg_rounded_rectangle2": (None, [_p] + [_f] * 6),
    "vkvg_elliptic_arc_to": (None, [_p, _f, _f, C.c_bool, C.c_bool, _f, _f, _f]),
    "vkvg_rel_elliptic_arc_to": (None, [_p, _f, _f, C.c_bool, C.c_bool, _f, _f, _f]),
    "vkvg_path_extents": (None, [_p] + [C.POINTER(_f)] * 4),
    "vkvg_stroke": (None, [_p]), "vkvg_stroke_preserve": (None, [_p]), "vkvg_fill": (None, [_p]), "vkvg_fill_preserve": (None, [_p]),
    "vkvg_paint": (None, [_p]), "vkvg_clear": (None, [_p]),
    "vkvg_clip": (None, [_p]), "vkvg_clip_preserve": (None, [_p]), "vkvg_reset_clip": (None, [_p]),
    "vkvg_set_opacity": (None, [_p, _f]), "vkvg_get_opacity": (_f, [_p]), "vkvg_set_source_color": (None, [_p, _u]),
    "vkvg_set_source_rgba": (None, [_p] + [_f] * 4), "vkvg_set_source_rgb": (None, [_p] + [_f] * 3), "vkvg_set_source": (None, [_p, _p]),
    "vkvg_set_line_width": (None, [_p, _f]), "vkvg_set_miter_limit": (None, [_p, _f]), "vkvg_get_miter_limit": (_f, [_p]),
    "vkvg_set_line_cap": (None, [_p, _i]), "vkvg_set_line_join": (None, [_p, _i]), "vkvg_set_operator": (None, [_p, _i]),
    "vkvg_set_fill_rule": (None, [_p, _i]), "vkvg_set_dash": (None, [_p, C.POINTER(_f), _u, _f]),
    "vkvg_get_dash": (None, [_p, C.POINTER(_f), C.POINTER(_u), C.POINTER(_f)]),
    "vkvg_get_line_width": (_f, [_p]), "vkvg_get_line_cap": (_i, [_p]), "vkvg_get_line_join": (_i, [_p]),
    "vkvg_get_operator": (_i, [_p]), "vkvg_get_fill_rule": (_i, [_p]), "vkvg_get_source": (_p, [_p]), "vkvg_get_target": (_p, [_p]),
    "vkvg_save": (None, [_p]), "vkvg_restore": (None, [_p]), "vkvg_translate": (None, [_p, _f, _f]), "vkvg_scale": (None, [_p, _f, _f]),
    "vkvg_rotate": (None, [_p, _f]), "vkvg_transform": (None, [_p, _p]), "vkvg_set_matrix": (None, [_p, _p]),
    "vkvg_get_matrix": (None, [_p, _p]), "vkvg_identity_matrix": (None, [_p]),
    "vkvg_matrix_init_identity": (None, [_p]), "vkvg_matrix_init": (None, [_p] + [_f] * 6),
    "vkvg_matrix_init_translate": (None, [_p, _f, _f]), "vkvg_matrix_init_scale": (None, [_p, _f, _f]),
    "vkvg_matrix_init_rotate": (None, [_p, _f]), "vkvg_matrix_translate": (None, [_p, _f, _f]), "vkvg_matrix_scale": (None, [_p, _f, _f]),
    "vkvg_matrix_rotate": (None, [_p, _f]), "vkvg_matrix_multiply": (None, [_p, _p, _p]),
    "vkvg_matrix_transform_distance": (None, [_p, C.POINTER(_f), C.POINTER(_f)]),
    "vkvg_matrix_transform_point": (None, [_p, C.POINTER(_f), C.POINTER(_f)]), "vkvg_matrix_invert": (_i, [_p]),
    "vkvg_matrix_get_scale": (None, [_p, C.POINTER(_f), C.POINTER(_f)]),
    "vkvg_pattern_status": (_i, [_p]), "vkvg_pattern_reference": (_p, [_p]), "vkvg_pattern_get_reference_count": (_u, [_p]),
    "vkvg_pattern_create_linear": (_p, [_f] * 4), "vkvg_pattern_create_radial": (_p, [_f] * 6), "vkvg_pattern_destroy": (None, [_p]),
    "vkvg_pattern_add_color_stop": (_i, [_p] + [_f] * 5), "vkvg_pattern_get_color_stop_count": (_i, [_p, C.POINTER(_u)]),
    "vkvg_pattern_get_type": (_i, [_p]), "vkvg_pattern_set_matrix": (None, [_p, _p]), "vkvg_pattern_get_matrix": (None, [_p, _p]),
    "vkvg_pattern_set_extend": (None, [_p, _i]), "vkvg_pattern_get_extend": (_i, [_p]),
    "vkvg_pattern_set_filter": (None, [_p, _i]), "vkvg_pattern_get_filter": (_i, [_p]),
    "vkvg_pattern_create_for_surface": (_p, [_p]), "vkvg_set_source_surface": (None, [_p, _p, _f, _f]),
    "vkvg_surface_create_from_image": (_p, [_p, C.c_char_p]), "vkvg_surface_create_from_bitmap": (_p, [_p, _p, _u, _u]),
    "vkvg_start_recording": (None, [_p]), "vkvg_stop_recording": (_p, [_p]), "vkvg_replay": (None, [_p, _p]),
    "vkvg_replay_command": (None, [_p, _p, _u]), "vkvg_recording_get_count": (_u, [_p]), "vkvg_recording_get_data": (_p, [_p]),
    "vkvg_recording_get_command": (None, [_p, _u, C.POINTER(_u), C.POINTER(_p)]), "vkvg_recording_destroy": (None, [_p]),
    # vkvg-svg.h
    "vkvg_svg_load": (_p, [C.c_char_p]), "vkvg_svg_load_fragment": (_p, [C.c_char_p]), "vkvg_svg_destroy": (None, [_p]),
    "vkvg_svg_get_dimensions": (None, [_p, C.POINTER(_u), C.POINTER(_u)]), "vkvg_svg_render": (None, [_p, _p, C.c_char_p]),
    "vkvg_surface_create_from_svg": (_p, [_p, _u, _u, C.c_char_p]), "vkvg_surface_create_from_svg_fragment": (_p, [_p, _u, _u, C.c_char_p]),
    "vkvg_b200_svg_serialize": (C.c_uint64, [_p, _p, C.c_uint64]),
    # vkvg_b200.h
    "vkvg_b200_flatten_path": (_u, [_p, _p, _p, _u, _p, _p, _u, C.POINTER(_u)]),
    "vkvg_b200_stroke_geometry": (None, [_p, _p, _u, C.POINTER(_u), _p, _u, C.POINTER(_u)]),
    "vkvg_b200_path_edges": (C.c_uint64, [_p, _i, _p, C.c_uint64]),
    "vkvg_b200_flush_capture_winding": (None, [_p, _p]),
    "vkvg_b200_winding": (_i, [_p, _p, C.c_uint64, _u, _u, _p]),
    "vkvg_b200_surface_read_premultiplied": (_i, [_p, _p]), "vkvg_b200_surface_set_readback": (_i, [_p, _p]),
    "vkvg_b200_surface_ipc_export": (_i, [_p, _p]), "vkvg_b200_ipc_open": (_p, [_p, _p]), "vkvg_b200_ipc_close": (_i, [_p, _p]),
    "vkvg_b200_launch_count": (C.c_uint64, []), "vkvg_b200_set_profiling": (None, [_p, _i]),
    "vkvg_b200_last_stats": (None, [_p, C.POINTER(Stats)]), "vkvg_b200_device_synchronize": (None, [_p]),
    "vkvg_b200_device_ordinal": (_i, [_p]), "vkvg_b200_surface_device_pointer": (_p, [_p]),
    "vkvg_b200_flush_keep": (None, [_p]), "vkvg_b200_replay_resident": (None, [_p, _p, _i]),
    "vkvg_b200_replay": (_i, [_p, _p, C.c_uint64, _p, C.c_uint64]),
    "vkvg_b200_submit": (_i, [_p, _p, C.c_uint64, _p, C.c_uint64]), "vkvg_b200_set_submit_decoder": (None, [_i]),
    "vkvg_b200_submit_counts": (None, [_p, _p]),
    "vkvg_b200_time_resident": (_i, [_p, _p, _u, _i, _i, C.POINTER(Stats)]),
    "vkvg_b200_device_set_graphs": (None, [_p, _i]), "vkvg_b200_device_set_stage_timing": (None, [_p, _i]),
    "vkvg_b200_device_graph_replays": (C.c_uint64, [_p]),
    "vkvg_b200_set_fine_kernel": (None, [_i]), "vkvg_b200_get_fine_kernel": (_i, []),
    "vkvg_b200_device_set_coverage_mode": (_i, [_p, _i]), "vkvg_b200_device_get_coverage_mode": (_i, [_p]),
    "vkvg_b200_get_source_push": (None, [_p, _p]),
    "vkvg_b200_surface_create_batch": (_p, [_p, _u, _u, _u]), "vkvg_b200_set_canvas": (_i, [_p, _u]),
    "vkvg_b200_surface_create_stripe": (_p, [_p, _u, _u, _u, _u]), "vkvg_b200_surface_copy_to_device": (_i, [_p, _p]),
}

_lib = None


def lib():
    """Load libvkvg_b200.so (building it first if it is missing) and attach prototypes."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            from . import build as _build
            _build.build()
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGS.items():
            fn = getattr(L, name)  # AttributeError here means include/*.h and the library disagree
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def exported_symbols():
    return sorted(_SIGS)


class VkvgError(RuntimeError):
    pass


COVERAGE_MSAA, COVERAGE_ANALYTIC = 0, 1


class Device:
    def __init__(self, samples=4, analytic=False):
        """analytic=True selects the analytic-coverage mode (include/vkvg_b200.h §0) instead of MSAA."""
        L = lib()
        info = DeviceCreateInfo(samples=samples)
        self.h = L.vkvg_device_create(C.byref(info))
        st = L.vkvg_device_status(self.h)
        if st != STATUS_SUCCESS:
            self.h = None
            raise VkvgError("vkvg_device_create failed: %s (no CUDA device? this library has no CPU path)" %
                            L.vkvg_status_to_string(st).decode())
        self.samples = samples
        self.analytic = bool(analytic)
        if analytic and L.vkvg_b200_device_set_coverage_mode(self.h, COVERAGE_ANALYTIC):
            raise VkvgError("vkvg_b200_device_set_coverage_mode failed")

    def close(self):
        if self.h:
            lib().vkvg_device_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_profiling(self, on=True):
        lib().vkvg_b200_set_profiling(self.h, int(on))

    def last_stats(self):
        s = Stats()
        lib().vkvg_b200_last_stats(self.h, C.byref(s))
        return s.as_dict()

    def set_graphs(self, on=True):
        lib().vkvg_b200_device_set_graphs(self.h, int(on))

    def set_stage_timing(self, on=True):
        lib().vkvg_b200_device_set_stage_timing(self.h, int(on))

    def graph_replays(self):
        return int(lib().vkvg_b200_device_graph_replays(self.h))

    def synchronize(self):
        lib().vkvg_b200_device_synchronize(self.h)

    def ipc_open(self, handle):
        """device address, valid in this process, of the surface image another process exported with Surface.ipc_export()"""
        p = lib().vkvg_b200_ipc_open(self.h, C.c_char_p(handle))
        if not p:
            raise VkvgError("vkvg_b200_ipc_open failed (no peer access between the two GPUs, or not the same node)")
        return int(p)

    def ipc_close(self, address):
        lib().vkvg_b200_ipc_close(self.h, C.c_void_p(address))

    def time_resident(self, surf, steps, clear_first=True, flush_l2=True):
        """re-run the pipeline `steps` times on the batch kept by the last flush; returns summed stats (ms) as a dict."""
        s = Stats()
        st = lib().vkvg_b200_time_resident(self.h, surf.h, steps, int(clear_first), int(flush_l2), C.byref(s))
        if st:
            raise VkvgError("vkvg_b200_time_resident: status %d" % st)
        return s.as_dict()

    def winding(self, edges, width, height):
        """per-sample integer winding of raw 24.8 edges through the tile rasteriser: (H, W, samples) int32."""
        e = np.ascontiguousarray(edges, np.int32).reshape(-1, 4)
        out = np.zeros((height, width) if self.analytic else (height, width, self.samples), np.int32)
        st = lib().vkvg_b200_winding(self.h, e.ctypes.data, len(e), width, height, out.ctypes.data)
        if st:
            raise VkvgError("vkvg_b200_winding: status %d" % st)
        return out.view(np.float32) if self.analytic else out

    def area(self, edges, width, height):
        """analytic mode: integral of the winding number over each pixel, (H, W) float32, through the tile rasteriser."""
        assert self.analytic
        return self.winding(edges, width, height)


class Surface:
    def __init__(self, dev, width, height, full_height=None, origin_y=0, batch=None):
        """full_height / origin_y: this surface is the stripe [origin_y, origin_y + height) of a taller logical surface.
        batch=n: n independent width x height canvases stacked in one surface (pixels() returns (n * height, width, 4))."""
        self.dev = dev
        self.width, self.height = width, height
        self.full_height, self.origin_y = full_height or height, origin_y
        self.batch = batch
        if batch:
            self.h = lib().vkvg_b200_surface_create_batch(dev.h, width, height, batch)
            self.height = height * batch
        elif full_height is None:
            self.h = lib().vkvg_surface_create(dev.h, width, height)
        else:
            self.h = lib().vkvg_b200_surface_create_stripe(dev.h, width, full_height, origin_y, height)
        st = lib().vkvg_surface_status(self.h)
        if st:
            raise VkvgError("vkvg_surface_create: status %d" % st)

    def close(self):
        if self.h:
            lib().vkvg_surface_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def clear(self):
        lib().vkvg_surface_clear(self.h)

    def device_pointer(self):
        """address of the premultiplied RGBA8 image in device memory (row-major, width * 4 bytes per row)"""
        return int(lib().vkvg_b200_surface_device_pointer(self.h))

    def set_readback(self, address):
        """every later flush also delivers the image to `address`, band by band while later bands render: pinned host memory, or device
        memory of this or another GPU (Device.ipc_open).  None / 0 switches it off."""
        st = lib().vkvg_b200_surface_set_readback(self.h, C.c_void_p(address or None))
        if st:
            raise VkvgError("vkvg_b200_surface_set_readback: status %d" % st)

    def wait_delivered(self, address):
        """returns once the last flush's image is at the read-back target `address` (no copy when the flush delivered it itself)"""
        st = lib().vkvg_b200_surface_read_premultiplied(self.h, C.c_void_p(address))
        if st:
            raise VkvgError("vkvg_b200_surface_read_premultiplied: status %d" % st)

    def ipc_export(self):
        """64-byte inter-process handle of the image (cudaIpcMemHandle_t) for Device.ipc_open in another process of this node"""
        buf = C.create_string_buffer(64)
        st = lib().vkvg_b200_surface_ipc_export(self.h, buf)
        if st:
            raise VkvgError("vkvg_b200_surface_ipc_export: status %d" % st)
        return bytes(buf.raw)

    def as_tensor(self):
        """the premultiplied RGBA8 image in device memory as a (H, W, 4) uint8 torch tensor WITHOUT a copy (what an NCCL gather of
        the stripes sends from).  The library renders on a stream of its own: call Device.synchronize() before reading it."""
        import torch

        class _View:
            pass
        view = _View()
        view.__cuda_array_interface__ = {"shape": (self.height, self.width, 4), "typestr": "|u1", "data": (int(lib().vkvg_b200_surface_device_pointer(self.h)), False),
                                         "version": 3, "strides": None}
        t = torch.as_tensor(view, device="cuda:%d" % lib().vkvg_b200_device_ordinal(self.dev.h))
        t._vkvg_surface = self   # (keeps the surface alive as long as the view)
        return t

    def pixels(self):
        """premultiplied RGBA8 exactly as stored, (H, W, 4)."""
        out = np.zeros((self.height, self.width, 4), np.uint8)
        st = lib().vkvg_b200_surface_read_premultiplied(self.h, out.ctypes.data)
        if st:
            raise VkvgError("surface read: status %d" % st)
        return out

    def write_to_memory(self, out=None):
        """vkvg_surface_write_to_memory: un-premultiplied RGBA8."""
        if out is None:
            out = np.zeros((self.height, self.width, 4), np.uint8)
        st = lib().vkvg_surface_write_to_memory(self.h, out.ctypes.data)
        if st:
            raise VkvgError("vkvg_surface_write_to_memory: status %d" % st)
        return out

    def copy_to_device(self, device_ptr):
        """premultiplied RGBA8 rows into caller-owned device memory (int pointer, e.g. torch.Tensor.data_ptr())."""
        st = lib().vkvg_b200_surface_copy_to_device(self.h, device_ptr)
        if st:
            raise VkvgError("vkvg_b200_surface_copy_to_device: status %d" % st)

    def write_to_png(self, path):
        return lib().vkvg_surface_write_to_png(self.h, path.encode())


_CTX_CALLS = ["new_path", "close_path", "new_sub_path", "line_to", "rel_line_to", "move_to", "rel_move_to", "arc", "arc_negative",
              "curve_to", "rel_curve_to", "quadratic_to", "rel_quadratic_to", "rectangle", "rounded_rectangle", "rounded_rectangle2", "elliptic_arc_to", "rel_elliptic_arc_to", "ellipse", "stroke",
              "stroke_preserve", "fill", "fill_preserve", "paint", "clear", "clip", "clip_preserve", "reset_clip", "set_opacity", "set_source_color", "set_source_rgba",
              "set_source_rgb", "set_line_width", "set_miter_limit", "set_line_cap", "set_line_join", "set_operator", "set_fill_rule",
              "save", "restore", "translate", "scale", "rotate", "identity_matrix", "flush"]


class Context:
    """vkvg context; drawing methods have the vkvg_* names without the prefix."""

    def __init__(self, surf):
        self.surf = surf
        self.h = lib().vkvg_create(surf.h)
        st = lib().vkvg_status(self.h)
        if st:
            raise VkvgError("vkvg_create: status %d" % st)

    def __getattr__(self, name):
        if name in _CTX_CALLS:
            fn = getattr(lib(), "vkvg_" + name)
            h = self.h
            return lambda *a: fn(h, *a)
        raise AttributeError(name)

    def close(self):
        if self.h:
            lib().vkvg_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def status(self):
        return lib().vkvg_status(self.h)

    def set_dash(self, dashes, offset=0.0):
        arr = (_f * max(len(dashes), 1))(*dashes)
        lib().vkvg_set_dash(self.h, arr, len(dashes), offset)

    def set_matrix(self, m):
        a = np.asarray(m, np.float32)
        lib().vkvg_set_matrix(self.h, a.ctypes.data)

    def get_matrix(self):
        a = np.zeros(6, np.float32)
        lib().vkvg_get_matrix(self.h, a.ctypes.data)
        return a

    def get_current_point(self):
        x, y = _f(), _f()
        lib().vkvg_get_current_point(self.h, C.byref(x), C.byref(y))
        return x.value, y.value

    def _grad(self, pat, stops):
        L = lib()
        for s in np.asarray(stops, np.float32).reshape(-1, 5):
            L.vkvg_pattern_add_color_stop(pat, *[float(v) for v in s])
        L.vkvg_set_source(self.h, pat)
        L.vkvg_pattern_destroy(pat)

    def set_source_linear(self, x0, y0, x1, y1, stops):
        self._grad(lib().vkvg_pattern_create_linear(x0, y0, x1, y1), stops)

    def set_source_radial(self, cx0, cy0, r0, cx1, cy1, r1, stops):
        self._grad(lib().vkvg_pattern_create_radial(cx0, cy0, r0, cx1, cy1, r1), stops)

    def set_source_surface(self, src, x=0.0, y=0.0, extend=None, filter=None, matrix=None):
        """src: Surface.  With extend / filter / matrix the source goes through an explicit pattern (vkvg_pattern_create_for_surface
        + vkvg_set_source), otherwise through vkvg_set_source_surface(ctx, surf, x, y)."""
        L = lib()
        if extend is None and filter is None and matrix is None:
            L.vkvg_set_source_surface(self.h, src.h, x, y)
            return
        L.vkvg_set_source_surface(self.h, src.h, x, y)      # sets the source offset
        pat = L.vkvg_pattern_create_for_surface(src.h)
        if extend is not None:
            L.vkvg_pattern_set_extend(pat, extend)
        if filter is not None:
            L.vkvg_pattern_set_filter(pat, filter)
        if matrix is not None:
            m = np.asarray(matrix, np.float32)
            L.vkvg_pattern_set_matrix(pat, m.ctypes.data)
        L.vkvg_set_source(self.h, pat)
        L.vkvg_pattern_destroy(pat)

    def start_recording(self):
        lib().vkvg_start_recording(self.h)

    def stop_recording(self):
        """returns an opaque VkvgRecording handle (or None when nothing was recorded); free it with lib().vkvg_recording_destroy"""
        return lib().vkvg_stop_recording(self.h)

    def replay_recording(self, rec):
        lib().vkvg_replay(self.h, rec)

    def source_push(self):
        out = np.zeros(10, np.float32)
        lib().vkvg_b200_get_source_push(self.h, out.ctypes.data)
        return out

    def set_canvas(self, index):
        st = lib().vkvg_b200_set_canvas(self.h, index)
        if st:
            raise VkvgError("vkvg_b200_set_canvas: status %d" % st)

    def path_extents(self):
        x1, y1, x2, y2 = _f(), _f(), _f(), _f()
        lib().vkvg_path_extents(self.h, C.byref(x1), C.byref(y1), C.byref(x2), C.byref(y2))
        return x1.value, y1.value, x2.value, y2.value

    def render_svg(self, svg, sub_id=None):
        lib().vkvg_svg_render(svg.h, self.h, sub_id.encode() if sub_id else None)

    # ---- stage introspection (vkvg_b200.h) ----
    def path_points(self):
        """flattened points of the current path, computed by the CUDA flatten kernels: (n,2) float32."""
        L = lib()
        ns = _u()
        n = L.vkvg_b200_flatten_path(self.h, None, None, 0, None, None, 0, C.byref(ns))
        xy = np.zeros((n, 2), np.float32)
        cur = np.zeros(n, np.uint8)
        first = np.zeros(ns.value, np.uint32)
        cnt = np.zeros(ns.value, np.uint32)
        L.vkvg_b200_flatten_path(self.h, xy.ctypes.data, cur.ctypes.data, n, first.ctypes.data, cnt.ctypes.data, ns.value, C.byref(ns))
        self._last_subpaths = (first, cnt, cur)
        return xy

    def path_subpaths(self):
        self.path_points()
        return self._last_subpaths

    def stroke_geometry(self):
        L = lib()
        nv, ni = _u(), _u()
        L.vkvg_b200_stroke_geometry(self.h, None, 0, C.byref(nv), None, 0, C.byref(ni))
        v = np.zeros((nv.value, 2), np.float32)
        ix = np.zeros(ni.value, np.uint32)
        L.vkvg_b200_stroke_geometry(self.h, v.ctypes.data, nv.value, C.byref(nv), ix.ctypes.data, ni.value, C.byref(ni))
        return v, ix

    def path_edges(self, stroke=False):
        L = lib()
        n = L.vkvg_b200_path_edges(self.h, int(stroke), None, 0)
        e = np.zeros((n, 4), np.int32)
        L.vkvg_b200_path_edges(self.h, int(stroke), e.ctypes.data, n)
        return e

    def flush_capture_winding(self):
        d = self.surf.dev
        out = np.zeros((self.surf.height, self.surf.width) if d.analytic else (self.surf.height, self.surf.width, d.samples), np.int32)
        lib().vkvg_b200_flush_capture_winding(self.h, out.ctypes.data)
        return out.view(np.float32) if d.analytic else out   # analytic mode: the area integral A of the last draw

    def submit(self, cmds, args):
        """vkvg_b200_submit: the stream with explicit argument counts (CommandStream.arrays2), decoded on the device when it can be, then flushed"""
        cmds = np.ascontiguousarray(cmds, np.uint32)
        args = np.ascontiguousarray(args, np.float32)
        return lib().vkvg_b200_submit(self.h, cmds.ctypes.data, len(cmds), args.ctypes.data, len(args))

    def replay(self, ops, args):
        ops = np.ascontiguousarray(ops, np.uint8)
        args = np.ascontiguousarray(args, np.float32)
        return lib().vkvg_b200_replay(self.h, ops.ctypes.data, len(ops), args.ctypes.data, len(args))


class Svg:
    """vkvg-svg.h: a parsed SVG document (host only: no device is needed to load or inspect one)."""

    def __init__(self, path=None, fragment=None):
        L = lib()
        self.h = L.vkvg_svg_load(os.fsencode(path)) if path is not None else L.vkvg_svg_load_fragment(fragment.encode() if isinstance(fragment, str) else fragment)
        if not self.h:
            raise VkvgError("vkvg_svg_load failed: %r" % (path,))

    def close(self):
        if self.h:
            lib().vkvg_svg_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def dimensions(self):
        w, h = _u(), _u()
        lib().vkvg_svg_get_dimensions(self.h, C.byref(w), C.byref(h))
        return w.value, h.value

    def serialize(self):
        """flat shape dump in the layout of oracle/nsvg_dump.c (bytes)."""
        L = lib()
        n = L.vkvg_b200_svg_serialize(self.h, None, 0)
        buf = np.zeros(n, np.uint8)
        L.vkvg_b200_svg_serialize(self.h, buf.ctypes.data, n)
        return buf.tobytes()


class CommandStream:
    """Builds the packed (ops, args) arrays consumed by vkvg_b200_replay; method names as on Context."""

    def __init__(self):
        self.ops = []
        self.args = []

    def _op(self, name, *a):
        self.ops.append(OPS[name])
        self.args.extend(float(v) for v in a)

    def __getattr__(self, name):
        key = name.upper()
        if key in OPS and key not in ("SET_DASH", "SET_SOURCE_LINEAR", "SET_SOURCE_RADIAL", "POLYLINE"):
            return lambda *a: self._op(key, *a)
        raise AttributeError(name)

    def set_dash(self, dashes, offset=0.0):
        self._op("SET_DASH", len(dashes), offset, *dashes)

    def set_source_linear(self, x0, y0, x1, y1, stops):
        s = np.asarray(stops, np.float32).reshape(-1, 5)
        self._op("SET_SOURCE_LINEAR", x0, y0, x1, y1, len(s), *s.ravel())

    def set_source_radial(self, cx0, cy0, r0, cx1, cy1, r1, stops):
        s = np.asarray(stops, np.float32).reshape(-1, 5)
        self._op("SET_SOURCE_RADIAL", cx0, cy0, r0, cx1, cy1, r1, len(s), *s.ravel())

    def polyline(self, xy):
        xy = np.ascontiguousarray(xy, np.float32).reshape(-1, 2)
        self.ops.append(OPS["POLYLINE"])
        self.args.append(np.array([len(xy)], np.uint32).view(np.float32)[0])
        self.args.append(xy.ravel())

    def arrays2(self):
        """(cmds uint32 = op | n_args << 8, args float32): the form vkvg_b200_submit takes - counts in the command words, not in the arguments"""
        ops, args = self.arrays()
        cmds = np.zeros(len(ops), np.uint32)
        out = []
        k = 0
        for i, op in enumerate(ops.tolist()):
            if op == OPS["POLYLINE"]:
                n = int(args[k:k + 1].view(np.uint32)[0])
                out.append(args[k + 1:k + 1 + 2 * n])
                cmds[i] = op | ((2 * n) << 8)
                k += 1 + 2 * n
            elif op == OPS["SET_DASH"]:
                n = int(args[k])
                out.append(args[k + 1:k + 2 + n])          # offset d0 .. dn-1
                cmds[i] = op | ((1 + n) << 8)
                k += 2 + n
            elif op in (OPS["SET_SOURCE_LINEAR"], OPS["SET_SOURCE_RADIAL"]):
                np_ = 4 if op == OPS["SET_SOURCE_LINEAR"] else 6
                ns = int(args[k + np_])
                out.append(args[k:k + np_])
                out.append(args[k + np_ + 1:k + np_ + 1 + 5 * ns])
                cmds[i] = op | ((np_ + 5 * ns) << 8)
                k += np_ + 1 + 5 * ns
            else:
                n = _FIXED_ARGS[op]
                out.append(args[k:k + n])
                cmds[i] = op | (n << 8)
                k += n
        assert k == len(args)
        return cmds, (np.concatenate(out) if out else np.zeros(0, np.float32)).astype(np.float32, copy=False)

    def arrays(self):
        parts = []
        run = []
        for a in self.args:
            if isinstance(a, np.ndarray):
                if run:
                    parts.append(np.asarray(run, np.float32))
                    run = []
                parts.append(a.astype(np.float32, copy=False))
            elif isinstance(a, np.floating):
                if run:
                    parts.append(np.asarray(run, np.float32))
                    run = []
                parts.append(np.array([a], np.float32))
            else:
                run.append(a)
        if run:
            parts.append(np.asarray(run, np.float32))
        args = np.concatenate(parts) if parts else np.zeros(0, np.float32)
        return np.asarray(self.ops, np.uint8), args
