"""CPU: the C-ABI library loads, exports every symbol include/*.h declares, and the host-only parts of the API
(matrices, patterns, status strings, NULL-safety) behave as the reference's gunit tests pin them
(gunit_tests/matrices.cpp:10-70, gunit_tests/context.cpp:30-164, gunit_tests/patternDraw.cpp:18-37).
No compute call is made here."""
import ctypes as C
import math
import os
import re

import numpy as np
import pytest

import vkvg_b200 as v
from vkvg_b200 import build as vbuild

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def L():
    vbuild.build()
    return v.lib()


def _declared(header):
    src = open(os.path.join(ROOT, "include", header)).read()
    return sorted(set(re.findall(r"vkvg_public[^;(]*?\b(vkvg_\w+)\s*\(", src)))


@pytest.mark.parametrize("header", ["vkvg.h", "vkvg-svg.h", "vkvg_b200.h"])
def test_every_declared_symbol_is_exported(L, header):
    names = _declared(header)
    assert len(names) > {"vkvg.h": 100, "vkvg_b200.h": 10, "vkvg-svg.h": 6}[header]
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing


def test_python_binding_table_matches_header(L):
    declared = set(_declared("vkvg.h")) | set(_declared("vkvg_b200.h")) | set(_declared("vkvg-svg.h"))
    assert set(v.exported_symbols()) <= declared


def test_headers_compile_as_c(tmp_path):
    import subprocess
    src = tmp_path / "t.c"
    src.write_text('#include "vkvg.h"\n#include "vkvg-svg.h"\n#include "vkvg_b200.h"\nint main(void){vkvg_matrix_t m; vkvg_matrix_init_identity(&m); return (int)m.x0;}\n')
    r = subprocess.run(["gcc", "-std=c11", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), "-c", str(src), "-o", str(tmp_path / "t.o")],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


class Mat(C.Structure):
    _fields_ = [(n, C.c_float) for n in ("xx", "yx", "xy", "yy", "x0", "y0")]

    def t(self):
        return tuple(getattr(self, n) for n, _ in self._fields_)


def _close(a, b):
    return np.allclose(np.array(a, np.float32), np.array(b, np.float32), rtol=4e-7, atol=0)


def test_matrix_init_known_answers(L):  # gunit_tests/matrices.cpp:10-36
    m = Mat()
    L.vkvg_matrix_init_identity(C.byref(m))
    assert m.t() == (1, 0, 0, 1, 0, 0)
    L.vkvg_matrix_init(C.byref(m), *[C.c_float(x) for x in (1.3, 2.5, 0.3, 0.7, 1.2, 1.7)])
    assert _close(m.t(), (1.3, 2.5, 0.3, 0.7, 1.2, 1.7))
    L.vkvg_matrix_init_translate(C.byref(m), 1.3, 2.5)
    assert _close(m.t(), (1, 0, 0, 1, 1.3, 2.5))
    L.vkvg_matrix_init_scale(C.byref(m), 2.1, 1.5)
    assert _close(m.t(), (2.1, 0, 0, 1.5, 0, 0))
    a, b = C.c_float(), C.c_float()
    L.vkvg_matrix_get_scale(C.byref(m), C.byref(a), C.byref(b))
    assert _close((a.value, b.value), (2.1, 1.5))
    L.vkvg_matrix_init_rotate(C.byref(m), 2.0)
    c, s = np.cos(np.float32(2)), np.sin(np.float32(2))
    assert _close(m.t(), (c, s, -s, c, 0, 0))


def test_matrix_invert_known_answers(L):  # gunit_tests/matrices.cpp:43-70
    INVALID_MATRIX, SUCCESS = 5, 0
    m = Mat(0, 0, 0, 0, 0, 0)
    assert L.vkvg_matrix_invert(C.byref(m)) == INVALID_MATRIX
    m = Mat(1, 1, 0, 0, 0, 0)
    assert L.vkvg_matrix_invert(C.byref(m)) == INVALID_MATRIX
    m = Mat(1, 0, 0, 1, 0, 0)
    assert L.vkvg_matrix_invert(C.byref(m)) == SUCCESS and m.t() == (1, 0, 0, 1, 0, 0)
    L.vkvg_matrix_init_scale(C.byref(m), 2.1, 1.5)
    assert L.vkvg_matrix_invert(C.byref(m)) == SUCCESS
    assert _close(m.t(), (1 / np.float32(2.1), 0, 0, 1 / np.float32(1.5), 0, 0))
    L.vkvg_matrix_init_translate(C.byref(m), 2.1, 1.5)
    assert L.vkvg_matrix_invert(C.byref(m)) == SUCCESS
    assert _close(m.t(), (1, 0, 0, 1, -2.1, -1.5))
    L.vkvg_matrix_init_rotate(C.byref(m), 2.0)
    assert L.vkvg_matrix_invert(C.byref(m)) == SUCCESS
    c, s = np.cos(np.float32(2)), np.sin(np.float32(2))
    assert _close(m.t(), (c, -s, s, c, 0, 0))


def test_matrix_multiply_and_transform_against_reference_object_code(L, oracle_lib):
    """vkvg_matrix_multiply/translate/scale/rotate/transform_point vs the reference's src/vkvg_matrix.c (compiled into oracle/_ref)."""
    if not oracle_lib.ref_available():
        pytest.skip("no oracle/_ref")
    R = oracle_lib.Ref.lib()
    rng = np.random.default_rng(5)
    for _ in range(50):
        vals = rng.uniform(-3, 3, 6).astype(np.float32)
        a, b = Mat(*vals), Mat(*vals)
        for lib, m in ((L, a), (R, b)):
            lib.vkvg_matrix_translate.argtypes = [C.c_void_p, C.c_float, C.c_float]
            lib.vkvg_matrix_scale.argtypes = [C.c_void_p, C.c_float, C.c_float]
            lib.vkvg_matrix_rotate.argtypes = [C.c_void_p, C.c_float]
            lib.vkvg_matrix_translate(C.byref(m), float(vals[0]), float(vals[1]))
            lib.vkvg_matrix_rotate(C.byref(m), float(vals[2]))
            lib.vkvg_matrix_scale(C.byref(m), float(vals[3]) + 4, float(vals[4]) + 4)
        assert a.t() == b.t()
        x1, y1, x2, y2 = C.c_float(1.25), C.c_float(-7.5), C.c_float(1.25), C.c_float(-7.5)
        R.vkvg_matrix_transform_point.argtypes = [C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_float)]
        L.vkvg_matrix_transform_point(C.byref(a), C.byref(x1), C.byref(y1))
        R.vkvg_matrix_transform_point(C.byref(b), C.byref(x2), C.byref(y2))
        assert (x1.value, y1.value) == (x2.value, y2.value)
        R.vkvg_matrix_invert.argtypes = [C.c_void_p]
        assert L.vkvg_matrix_invert(C.byref(a)) == R.vkvg_matrix_invert(C.byref(b))
        assert a.t() == b.t()


def test_null_handles_are_safe(L):  # gunit_tests/context.cpp:30-41, :73-164
    NULL_POINTER, INVALID_SURFACE = 2, 19
    assert L.vkvg_status(None) == NULL_POINTER
    assert L.vkvg_get_reference_count(None) == 0
    ctx = L.vkvg_create(None)
    assert L.vkvg_status(ctx) == INVALID_SURFACE
    assert L.vkvg_get_reference_count(ctx) == 0
    for name in ("new_path", "close_path", "new_sub_path", "stroke", "stroke_preserve", "fill", "fill_preserve", "paint", "clear", "save",
                 "restore", "identity_matrix", "flush"):
        getattr(L, "vkvg_" + name)(ctx)
    L.vkvg_line_to(ctx, 0, 0)
    L.vkvg_move_to(ctx, 0, 0)
    L.vkvg_rel_line_to(ctx, 0, 0)
    L.vkvg_arc(ctx, 0, 0, 0, 0, 0)
    L.vkvg_curve_to(ctx, 0, 0, 0, 0, 0, 0)
    L.vkvg_rectangle(ctx, 0, 0, 0, 0)
    L.vkvg_set_source_rgba(ctx, 0, 0, 0, 0)
    L.vkvg_set_source(ctx, None)
    L.vkvg_set_dash(ctx, None, 0, 0)
    L.vkvg_get_dash(ctx, None, None, None)
    L.vkvg_get_current_point(ctx, None, None)
    L.vkvg_set_matrix(ctx, None)
    L.vkvg_get_matrix(ctx, None)
    L.vkvg_transform(ctx, None)
    assert L.vkvg_get_opacity(ctx) == 0 and L.vkvg_get_line_width(ctx) == 0 and L.vkvg_get_miter_limit(ctx) == 0
    assert L.vkvg_get_line_cap(ctx) == v.CAP_BUTT and L.vkvg_get_line_join(ctx) == v.JOIN_MITER
    assert L.vkvg_get_fill_rule(ctx) == v.FILL_NON_ZERO and L.vkvg_get_operator(ctx) == 2  # VKVG_OPERATOR_OVER
    assert not L.vkvg_has_current_point(ctx)
    L.vkvg_destroy(ctx)
    assert L.vkvg_surface_status(None) == NULL_POINTER and L.vkvg_device_status(None) == NULL_POINTER
    assert L.vkvg_surface_get_width(None) == 0
    L.vkvg_surface_destroy(None)
    L.vkvg_device_destroy(None)
    surf = L.vkvg_surface_create(None, 16, 16)
    assert L.vkvg_surface_status(surf) != 0
    L.vkvg_surface_destroy(surf)


def test_pattern_objects(L):  # gunit_tests/patternDraw.cpp:18-37, src/vkvg_pattern.c:95-167
    p = L.vkvg_pattern_create_linear(0, 0, 10, 10)
    assert L.vkvg_pattern_status(p) == 0 and L.vkvg_pattern_get_reference_count(p) == 1
    assert L.vkvg_pattern_get_type(p) == 2  # VKVG_PATTERN_TYPE_LINEAR
    L.vkvg_pattern_reference(p)
    assert L.vkvg_pattern_get_reference_count(p) == 2
    L.vkvg_pattern_destroy(p)
    assert L.vkvg_pattern_get_reference_count(p) == 1
    n = C.c_uint32(99)
    assert L.vkvg_pattern_get_color_stop_count(p, C.byref(n)) == 0 and n.value == 0
    assert L.vkvg_pattern_add_color_stop(p, 0.0, 1, 0, 0, 1) == 0
    assert L.vkvg_pattern_add_color_stop(p, 1.0, 0, 0, 1, 0.5) == 0
    L.vkvg_pattern_get_color_stop_count(p, C.byref(n))
    assert n.value == 2
    L.vkvg_pattern_set_extend(p, 3)
    assert L.vkvg_pattern_get_extend(p) == 3
    L.vkvg_pattern_destroy(p)
    r = L.vkvg_pattern_create_radial(10, 10, 50, 12, 12, 20)  # inner radius clamped to r1 - 1 (src/vkvg_pattern.c:95-118)
    assert L.vkvg_pattern_status(r) == 0 and L.vkvg_pattern_get_type(r) == 3
    L.vkvg_pattern_destroy(r)
    assert L.vkvg_pattern_status(None) == 2


def test_status_strings(L):  # include/vkvg.h:1647-1692
    assert L.vkvg_status_to_string(0) == b"no error has occurred"
    assert b"dash" in L.vkvg_status_to_string(13)
    assert L.vkvg_status_to_string(12345).startswith(b"<unknown")


def test_no_cpu_fallback(L):
    """without a CUDA device the library must refuse to create a device (loudly), never render on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(v.VkvgError):
        v.Device(4)


def test_product_never_imports_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "vkvg_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                s = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "import oracle" not in s and "liboracle" not in s and "ovk_" not in s, f
