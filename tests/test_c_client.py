"""A compiled C client of the drop-in boundary: tests/c_client/offscreen.c (shaped like the reference's tests/offscreen.c) built with gcc
against include/vkvg.h and linked with vkvg_b200/libvkvg_b200.so.  CPU: it compiles and links with -Wall -Werror as C11.  GPU: it runs,
and the PNG it writes equals the oracle's rendering of the same calls (un-premultiplied as vkvg_surface_write_to_png does)."""
import os
import subprocess
import zlib

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "c_client", "offscreen.c")


def _build(tmp_path):
    import vkvg_b200.build as vb
    vb.build()
    exe = str(tmp_path / "offscreen")
    libdir = os.path.join(ROOT, "vkvg_b200")
    r = subprocess.run(["gcc", "-std=c11", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), SRC, "-o", exe, "-L", libdir, "-lvkvg_b200",
                        "-Wl,-rpath," + libdir], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def test_c_client_compiles_and_links(tmp_path):
    exe = _build(tmp_path)
    assert os.path.exists(exe)


def _read_png(path):
    data = open(path, "rb").read()
    assert data[:8] == b"\x89PNG\r\n\x1a\n"
    pos, idat, dims = 8, b"", None
    while pos < len(data):
        n = int.from_bytes(data[pos:pos + 4], "big")
        typ, body = data[pos + 4:pos + 8], data[pos + 8:pos + 8 + n]
        assert zlib.crc32(typ + body) == int.from_bytes(data[pos + 8 + n:pos + 12 + n], "big")
        if typ == b"IHDR":
            dims = (int.from_bytes(body[:4], "big"), int.from_bytes(body[4:8], "big"), body[8], body[9])
        if typ == b"IDAT":
            idat += body
        pos += 12 + n
    w, h, depth, ctype = dims
    assert (depth, ctype) == (8, 6)
    raw = np.frombuffer(zlib.decompress(idat), np.uint8).reshape(h, 1 + 4 * w)
    assert not raw[:, 0].any()   # filter type 0 on every row
    return raw[:, 1:].reshape(h, w, 4)


@pytest.mark.gpu
def test_c_client_png_equals_the_oracle(tmp_path, oracle_lib):
    exe = _build(tmp_path)
    out = str(tmp_path / "offscreen.png")
    r = subprocess.run([exe, out], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    img = _read_png(out)
    o = oracle_lib.Oracle(256, 192, 4)
    o.rectangle(10, 10, 120, 90)
    o.set_source_rgba(1, 0, 0, 1)
    o.fill()
    o.set_fill_rule(0)
    o.move_to(60.5, 40.25)
    o.curve_to(200, 10, 240, 180, 100.75, 150)
    o.line_to(180, 60)
    o.close_path()
    o.set_source_linear(40, 20, 220, 170, [(0.0, 0.1, 0.3, 0.9, 1.0), (1.0, 0.9, 0.8, 0.1, 0.5)])
    o.fill_preserve()
    o.set_source_rgba(0.0, 0.4, 0.1, 0.8)
    o.set_line_width(5.0)
    o.set_line_join(1)
    o.set_dash([9.0, 4.0], 1.5)
    o.stroke()
    assert np.array_equal(img, o.write_to_memory())
