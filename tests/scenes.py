"""Scenes written once against the common drawing vocabulary (move_to, line_to, curve_to, arc, close_path,
set_*, fill, stroke, paint) and replayed on the oracle, on the reference and on the CUDA library.

Random inputs follow BASELINE.md: splitmix64(seed) -> uniform floats in [0,1).
"""
import math

import numpy as np


class SplitMix64:
    def __init__(self, seed):
        self.s = np.uint64(seed)

    def next_u64(self):
        with np.errstate(over="ignore"):
            self.s = self.s + np.uint64(0x9E3779B97F4A7C15)
            z = self.s
            z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
            z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
            return z ^ (z >> np.uint64(31))

    def u(self):
        return float(self.next_u64() >> np.uint64(11)) / float(1 << 53)

    def uniform(self, a, b):
        return a + (b - a) * self.u()

    def randint(self, a, b):  # inclusive
        return a + int(self.u() * (b - a + 1))


def splitmix_array(seed, n):
    """n uniform doubles in [0,1), vectorised splitmix64 (same sequence as SplitMix64(seed).u())."""
    with np.errstate(over="ignore"):
        idx = np.arange(1, n + 1, dtype=np.uint64)
        z = np.uint64(seed) + idx * np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return (z >> np.uint64(11)).astype(np.float64) / float(1 << 53)


def random_path(g, seed, size=256, curves=True, arcs=True):
    r = SplitMix64(seed)
    for _ in range(r.randint(1, 3)):
        g.move_to(r.uniform(20, size - 20), r.uniform(20, size - 20))
        for _ in range(r.randint(2, 8)):
            t = r.randint(0, 3)
            if t == 0 or (t == 1 and not curves) or (t >= 2 and not arcs):
                g.line_to(r.uniform(10, size - 10), r.uniform(10, size - 10))
            elif t == 1:
                g.curve_to(*[r.uniform(0, size) for _ in range(6)])
            elif t == 2:
                g.arc(r.uniform(50, size - 50), r.uniform(50, size - 50), r.uniform(3, 60), r.uniform(-4, 4), r.uniform(-4, 7))
            else:
                g.arc_negative(r.uniform(50, size - 50), r.uniform(50, size - 50), r.uniform(3, 60), r.uniform(-4, 4), r.uniform(-7, 4))
        if r.u() < 0.5:
            g.close_path()


def star(g, cx, cy, r1, r2, n):
    for i in range(2 * n):
        a = math.pi * i / n
        rr = r1 if i % 2 == 0 else r2
        (g.move_to if i == 0 else g.line_to)(cx + rr * math.cos(a), cy + rr * math.sin(a))
    g.close_path()


def polygons_c2(n, size, seed):
    """BASELINE C2: n closed polygons, K~U{5..16} vertices at radius U[4,32] around a uniform centre with
    unsorted uniform angles (self-intersecting).  Returns (list of (k,2) float32 arrays, (n,4) rgba float32)."""
    u = splitmix_array(seed, n * (3 + 2 * 16 + 4))
    u = u.reshape(n, -1)
    cx, cy = u[:, 0] * size, u[:, 1] * size
    k = 5 + (u[:, 2] * 12).astype(int)
    polys = []
    for i in range(n):
        rad = 4 + 28 * u[i, 3:3 + k[i]]
        ang = 2 * math.pi * u[i, 19:19 + k[i]]
        polys.append(np.stack([cx[i] + rad * np.cos(ang), cy[i] + rad * np.sin(ang)], 1).astype(np.float32))
    cols = u[:, 35:39].astype(np.float32)
    return polys, cols


def polyline_c3(n_points, size, seed, margin=10.0):
    """BASELINE C3: random walk, step U[2,14] px, turn U[-60,60] degrees, reflected at a margin."""
    u = splitmix_array(seed, 2 * n_points + 3)
    step = 2 + 12 * u[3:3 + n_points]
    turn = (u[3 + n_points:3 + 2 * n_points] - 0.5) * (2 * math.pi / 3)
    pts = np.zeros((n_points, 2), np.float64)
    x, y = margin + u[0] * (size - 2 * margin), margin + u[1] * (size - 2 * margin)
    a = u[2] * 2 * math.pi
    lo, hi = margin, size - margin
    for i in range(n_points):
        pts[i] = (x, y)
        a += turn[i]
        nx, ny = x + step[i] * math.cos(a), y + step[i] * math.sin(a)
        if nx < lo or nx > hi:
            a = math.pi - a
            nx = x + step[i] * math.cos(a)
        if ny < lo or ny > hi:
            a = -a
            ny = y + step[i] * math.sin(a)
        x, y = min(max(nx, lo), hi), min(max(ny, lo), hi)
    return pts.astype(np.float32)


def curves_c4(n, size, seed):
    """BASELINE C4: n closed paths of 4..8 cubic segments, control points within U[16,96] px of a uniform centre;
    alternating 3-stop linear / radial gradients with alpha in {1, 0.5}."""
    u = splitmix_array(seed, n * (3 + 8 * 6 * 2 + 16)).reshape(n, -1)
    out = []
    for i in range(n):
        cx, cy = u[i, 0] * size, u[i, 1] * size
        k = 4 + int(u[i, 2] * 5)
        rr = 16 + 80 * u[i, 3:3 + 3 * k]
        aa = 2 * math.pi * u[i, 51:51 + 3 * k]
        px, py = cx + rr * np.cos(aa), cy + rr * np.sin(aa)
        pts = np.stack([px, py], 1).astype(np.float32).reshape(k, 3, 2)
        alpha = 1.0 if (i // 2) % 2 == 0 else 0.5
        c = u[i, 99:99 + 9].reshape(3, 3)
        stops = [(0.0, c[0, 0], c[0, 1], c[0, 2], alpha), (0.5, c[1, 0], c[1, 1], c[1, 2], alpha), (1.0, c[2, 0], c[2, 1], c[2, 2], alpha)]
        out.append((pts, i % 2, (cx, cy), stops))
    return out


def load_nsvg(path):
    """shape list written by oracle/nsvg_dump.c (nanoSVG of the reference at 96 dpi)."""
    import struct
    data = open(path, "rb").read()
    assert data[:4] == b"NSVG"
    w, h, n = struct.unpack_from("<ffI", data, 4)
    off = 16
    shapes = []
    for _ in range(n):
        ft, fc, st, sc, op, sw, npaths = struct.unpack_from("<IIIIffI", data, off)
        off += 28
        paths = []
        for _ in range(npaths):
            npts, closed = struct.unpack_from("<II", data, off)
            off += 8
            pts = np.frombuffer(data, np.float32, 2 * npts, off).reshape(npts, 2).copy()
            off += 8 * npts
            paths.append((pts, bool(closed)))
        shapes.append(dict(fill_type=ft, fill_color=fc, stroke_type=st, stroke_color=sc, opacity=op, stroke_width=sw, paths=paths))
    return w, h, shapes


def _svg_color(g, c, alpha):
    a = (c >> 24 & 255) / 255.0
    b = (c >> 16 & 255) / 255.0
    gg = (c >> 8 & 255) / 255.0
    r = (c & 255) / 255.0
    g.set_source_rgba(r, gg, b, a * alpha)


def render_nsvg(g, shapes):
    """vkvg_svg_render, reference src/nsvg/vkvg_nsvg.c:79-136 (save/restore are the caller's business)."""
    g.set_fill_rule(0)
    g.set_source_rgba(0.0, 0.0, 0.0, 1.0)
    for s in shapes:
        g.new_path()
        g.set_line_width(s["stroke_width"])
        for pts, closed in s["paths"]:
            g.move_to(float(pts[0, 0]), float(pts[0, 1]))
            for i in range(1, len(pts) - 2, 3):
                g.curve_to(*[float(v) for v in pts[i:i + 3].ravel()])
            if closed:
                g.close_path()
        if s["fill_type"] in (1, 2):
            _svg_color(g, s["fill_color"], s["opacity"])
        if s["fill_type"] != 0:
            if s["stroke_type"] == 0:
                g.fill()
                continue
            g.fill_preserve()
        if s["stroke_type"] in (1, 2):
            _svg_color(g, s["stroke_color"], s["opacity"])
        g.stroke()
