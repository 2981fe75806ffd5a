"""CPU: the decomposition behind flatten_count_warp_k (vkvg_b200/csrc/flatten.cu: bez_lane_walk + bez_dfs).  A warp flattens one cubic: lane L
walks the subtree below the depth-5 node whose path from the root is L (bit 4 first, 0 = left child), a node on the way down that passes the
leaf test belongs to the leftmost lane below it, and the depth limit counts the right siblings pending above the lane's start node.  The
claim is that the lanes' points, concatenated in lane order, are exactly the points of the serial depth-first walk (flatten_cubic) - same
leaf set, same order, same behaviour at the depth limit.  Both walks are restated here over ONE leaf test, so the comparison is about the
traversal alone; the arithmetic itself is pinned by the GPU tests (tests/test_gpu_parity.py against the reference's goldens,
tests/test_gpu_variants.py: the two kernels agree bit for bit)."""
import math
import random

PI = 3.14159265358979323846
TWO_OVER_PI = 0.63661977236758134308


def mids(b):
    x1, y1, x2, y2, x3, y3, x4, y4 = b
    x12, y12, x23, y23, x34, y34 = (x1 + x2) / 2, (y1 + y2) / 2, (x2 + x3) / 2, (y2 + y3) / 2, (x3 + x4) / 2, (y3 + y4) / 2
    x123, y123, x234, y234 = (x12 + x23) / 2, (y12 + y23) / 2, (x23 + x34) / 2, (y23 + y34) / 2
    x1234, y1234 = (x123 + x234) / 2, (y123 + y234) / 2
    left = (x1, y1, x12, y12, x123, y123, x1234, y1234)
    right = (x1234, y1234, x234, y234, x34, y34, x4, y4)
    return left, right, x1234, y1234


def leaf_test(b, x1234, y1234, tol):
    """the points a leaf emits, or None when the node must be subdivided (bez_leaf_test)"""
    x1, y1, x2, y2, x3, y3, x4, y4 = b
    dx, dy = x4 - x1, y4 - y1
    d2 = abs((x2 - x4) * dy - (y2 - y4) * dx)
    d3 = abs((x3 - x4) * dy - (y3 - y4) * dx)
    if d2 > 1.7 and d3 > 1.7:
        if (d2 + d3) * (d2 + d3) <= (dx * dx + dy * dy) * tol:
            a23 = math.atan2(y3 - y2, x3 - x2)
            da1 = abs(a23 - math.atan2(y2 - y1, x2 - x1))
            da2 = abs(math.atan2(y4 - y3, x4 - x3) - a23)
            if da1 >= PI:
                da1 = TWO_OVER_PI - da1
            if da2 >= PI:
                da2 = TWO_OVER_PI - da2
            if da1 + da2 < 0.01:
                return [(x1234, y1234)]
            if da1 > 0.01:
                return [(x2, y2)]
            if da2 > 0.01:
                return [(x3, y3)]
    elif d2 > 1.7:
        if d2 * d2 <= tol * (dx * dx + dy * dy):
            da1 = abs(math.atan2(y3 - y2, x3 - x2) - math.atan2(y2 - y1, x2 - x1))
            if da1 >= PI:
                da1 = TWO_OVER_PI - da1
            if da1 < 0.01:
                return [(x2, y2), (x3, y3)]
            if da1 > 0.01:
                return [(x2, y2)]
    elif d3 > 1.7:
        if d3 * d3 <= tol * (dx * dx + dy * dy):
            da1 = abs(math.atan2(y4 - y3, x4 - x3) - math.atan2(y3 - y2, x3 - x2))
            if da1 >= PI:
                da1 = TWO_OVER_PI - da1
            if da1 < 0.01:
                return [(x2, y2), (x3, y3)]
            if da1 > 0.01:
                return [(x3, y3)]
    else:
        ddx, ddy = x1234 - (x1 + x4) / 2, y1234 - (y1 + y4) / 2
        if ddx * ddx + ddy * ddy <= tol:
            return [(x1234, y1234)]
    return None


def dfs(out, cur, level, pending, tol, limit):
    """bez_dfs: the serial walk from any node; `pending` right siblings wait above it"""
    stack = []
    while True:
        left, right, x1234, y1234 = mids(cur)
        pts = leaf_test(cur, x1234, y1234, tol) if level > 0 else None
        if pts is None and pending + len(stack) < limit:
            stack.append((right, level + 1))
            cur, level = left, level + 1
            continue
        if pts:
            out.extend(pts)
        if not stack:
            return
        cur, level = stack.pop()


def serial(e, tol, limit):
    out = []
    dfs(out, e, 0, 0, tol, limit)
    return out


def lane_walk(out, e, lane, tol, limit):
    cur, pending = e, 0
    for d in range(5):
        left, right, x1234, y1234 = mids(cur)
        if d > 0:
            pts = leaf_test(cur, x1234, y1234, tol)
            if pts is not None:
                if lane & ((1 << (5 - d)) - 1) == 0:
                    out.extend(pts)
                return
        if (lane >> (4 - d)) & 1:
            cur = right
        else:
            cur, pending = left, pending + 1
    dfs(out, cur, 5, pending, tol, limit)


def by_lanes(e, tol, limit):
    out = []
    for lane in range(32):
        lane_walk(out, e, lane, tol, limit)
    return out


def curves(n, seed):
    r = random.Random(seed)
    for i in range(n):
        scale = 10.0 ** r.uniform(-1, 4)          # from sub-pixel wiggles to curves thousands of pixels long
        kind = i % 5
        if kind == 0:
            e = tuple(r.uniform(-scale, scale) for _ in range(8))
        elif kind == 1:   # nearly straight: leaves high in the tree
            x0, y0, x1, y1 = (r.uniform(-scale, scale) for _ in range(4))
            e = (x0, y0, x0 + (x1 - x0) / 3 + r.uniform(-1e-3, 1e-3), y0 + (y1 - y0) / 3, x0 + 2 * (x1 - x0) / 3, y0 + 2 * (y1 - y0) / 3 + r.uniform(-1e-3, 1e-3), x1, y1)
        elif kind == 2:   # a cusp / loop: deep on one side only
            x0, y0 = r.uniform(-scale, scale), r.uniform(-scale, scale)
            e = (x0, y0, x0 + scale, y0 + scale, x0 - scale, y0 + scale, x0 + r.uniform(-1, 1), y0 + r.uniform(-1, 1))
        elif kind == 3:   # degenerate: coincident control points
            x0, y0 = r.uniform(-scale, scale), r.uniform(-scale, scale)
            e = (x0, y0, x0, y0, x0, y0, x0 + r.choice([0.0, scale]), y0)
        else:             # one end far away: many levels on the far side
            e = (0.0, 0.0, r.uniform(0, 5), r.uniform(0, 5), r.uniform(0, 5), r.uniform(0, 5), scale * 50, scale * 30)
        yield e, 10.0 ** r.uniform(-3, 1)


def test_lanes_in_order_are_the_serial_walk():
    depth_seen = 0
    for e, tol in curves(3000, 11):
        a, b = serial(e, tol, 40), by_lanes(e, tol, 40)
        assert a == b, (e, tol, len(a), len(b))
        depth_seen = max(depth_seen, len(a))
    assert depth_seen > 200   # (the set holds curves that really subdivide)


def test_depth_limit_cuts_in_the_same_place():
    """with a limit the tree really reaches: nodes that may not be subdivided any more emit nothing, in both walks alike"""
    cut = 0
    for limit in (5, 6, 8, 12):
        for e, tol in curves(1500, 100 + limit):
            a, b = serial(e, tol, limit), by_lanes(e, tol, limit)
            assert a == b, (limit, e, tol, len(a), len(b))
            cut += a != serial(e, tol, 40)
    assert cut > 100   # (the limits did change the output: the test is about something)
