// Flattening kernels: path elements -> points.
//
// Replaces the CPU recursion of the reference (vkvg_curve_to -> _recursive_bezier,
// src/vkvg_context.c:541-566, src/vkvg_context_internal.c:1305-1461) and the arc stepping loops
// (src/vkvg_context.c:394-503).  One thread per element walks the same adaptive subdivision tree with an
// explicit stack (left child first), once to count and once to emit at the offset given by a prefix scan,
// so the output order is exactly the reference's DFS order.
//
// Compiled with --fmad=false: every float expression is evaluated as the reference's gcc build does.
#include "pipeline.h"
#include <stdlib.h>

#define PIF 3.14159265358979323846f
#define TWO_OVER_PIF 0.63661977236758134308f  // the reference's M_2_PIF (2/pi), used where 2*pi was meant
#define VKB_BEZ_STACK 40                      // reference limit is 100 levels; finite input never gets near 40

// The counting pass keeps the first cache_n points (VKB_FLAT_CACHE, or VKB_FLAT_CACHE_SMALL_SCENE for batches of few elements, where the
// longest curve of the frame is what the emitting pass waits for) of every curved element (most cubics flatten to fewer), so the
// emitting pass copies them instead of walking the subdivision tree a second time.
#define VKB_FLAT_CACHE 16
struct PointSink {
    float2  *pts;
    uint8_t *flags;
    uint32_t n;
    uint8_t  flag;
    bool     emit;
    float2  *cache;    // counting pass: this element's cache_cap slots, or null
    uint32_t n_cached, cache_cap;
    __device__ __forceinline__ void add(float x, float y) {
        if (isnan(x) || isnan(y)) return;  // _add_point drops NaN, internal.c:224
        if (emit) {
            pts[n]   = make_float2(x, y);
            flags[n] = flag;
        } else if (cache && n_cached < cache_cap) cache[n_cached++] = make_float2(x, y);
        n++;
    }
};

struct BezNode {
    float x1, y1, x2, y2, x3, y3, x4, y4;
};

// returns true when the node is a leaf (points were emitted), false when it must be subdivided
__device__ __forceinline__ bool bez_leaf_test(PointSink &s, float tol, const BezNode &b, float x1234, float y1234) {
    float dx = b.x4 - b.x1, dy = b.y4 - b.y1;
    float d2 = fabsf(((b.x2 - b.x4) * dy - (b.y2 - b.y4) * dx));
    float d3 = fabsf(((b.x3 - b.x4) * dy - (b.y3 - b.y4) * dx));
    float da1, da2;
    // thresholds are double literals in the reference (1.7, 0.01); compare in double like it does
    if ((double)d2 > 1.7 && (double)d3 > 1.7) {
        if ((d2 + d3) * (d2 + d3) <= (dx * dx + dy * dy) * tol) {
            float a23 = atan2f(b.y3 - b.y2, b.x3 - b.x2);
            da1       = fabsf(a23 - atan2f(b.y2 - b.y1, b.x2 - b.x1));
            da2       = fabsf(atan2f(b.y4 - b.y3, b.x4 - b.x3) - a23);
            if (da1 >= PIF) da1 = TWO_OVER_PIF - da1;
            if (da2 >= PIF) da2 = TWO_OVER_PIF - da2;
            if (da1 + da2 < (float)0.01) { s.add(x1234, y1234); return true; }
            if ((double)da1 > 0.01) { s.add(b.x2, b.y2); return true; }
            if ((double)da2 > 0.01) { s.add(b.x3, b.y3); return true; }
        }
    } else {
        if ((double)d2 > 1.7) {
            if (d2 * d2 <= tol * (dx * dx + dy * dy)) {
                da1 = fabsf(atan2f(b.y3 - b.y2, b.x3 - b.x2) - atan2f(b.y2 - b.y1, b.x2 - b.x1));
                if (da1 >= PIF) da1 = TWO_OVER_PIF - da1;
                if ((double)da1 < 0.01) { s.add(b.x2, b.y2); s.add(b.x3, b.y3); return true; }
                if ((double)da1 > 0.01) { s.add(b.x2, b.y2); return true; }
            }
        } else if ((double)d3 > 1.7) {
            if (d3 * d3 <= tol * (dx * dx + dy * dy)) {
                da1 = fabsf(atan2f(b.y4 - b.y3, b.x4 - b.x3) - atan2f(b.y3 - b.y2, b.x3 - b.x2));
                if (da1 >= PIF) da1 = TWO_OVER_PIF - da1;
                if ((double)da1 < 0.01) { s.add(b.x2, b.y2); s.add(b.x3, b.y3); return true; }
                if ((double)da1 > 0.01) { s.add(b.x3, b.y3); return true; }
            }
        } else {
            dx = x1234 - (b.x1 + b.x4) / 2;
            dy = y1234 - (b.y1 + b.y4) / 2;
            if (dx * dx + dy * dy <= tol) { s.add(x1234, y1234); return true; }
        }
    }
    return false;
}

// The reference's depth-first walk, startable at any node of the subdivision tree: `pending` = right siblings that wait above the node
// (one per left turn on the way down from the root), so that the depth limit cuts where it would in a walk from the root.
__device__ void bez_dfs(PointSink &s, BezNode cur, unsigned level, int pending, const float tol) {
    BezNode  stack[VKB_BEZ_STACK];
    uint8_t  lvl[VKB_BEZ_STACK];
    int      sp  = 0;
    for (;;) {
        // de Casteljau midpoints, internal.c:1321-1332
        float x12 = (cur.x1 + cur.x2) / 2, y12 = (cur.y1 + cur.y2) / 2;
        float x23 = (cur.x2 + cur.x3) / 2, y23 = (cur.y2 + cur.y3) / 2;
        float x34 = (cur.x3 + cur.x4) / 2, y34 = (cur.y3 + cur.y4) / 2;
        float x123 = (x12 + x23) / 2, y123 = (y12 + y23) / 2;
        float x234 = (x23 + x34) / 2, y234 = (y23 + y34) / 2;
        float x1234 = (x123 + x234) / 2, y1234 = (y123 + y234) / 2;
        bool  leaf = level > 0 && bez_leaf_test(s, tol, cur, x1234, y1234);  // level 0 always subdivides
        if (!leaf && pending + sp < VKB_BEZ_STACK) {
            // right child waits on the stack, continue with the left one (internal.c:1459-1460)
            stack[sp] = BezNode{x1234, y1234, x234, y234, x34, y34, cur.x4, cur.y4};
            lvl[sp]   = (uint8_t)(level + 1);
            sp++;
            cur = BezNode{cur.x1, cur.y1, x12, y12, x123, y123, x1234, y1234};
            level++;
            continue;
        }
        // leaf, or depth limit reached (the reference returns without a point past its own limit)
        if (sp == 0) break;
        sp--;
        cur   = stack[sp];
        level = lvl[sp];
    }
}
__device__ void flatten_cubic(PointSink &s, const float *e) {
    bez_dfs(s, BezNode{e[0], e[1], e[2], e[3], e[4], e[5], e[6], e[7]}, 0, 0, e[8]);
    s.add(e[6], e[7]);  // end point appended unconditionally, vkvg_context.c:564
}

// One lane's share of a cubic when a WARP flattens it: the subtree below the depth-5 node whose path from the root is the lane number (bit 4
// decides first, 0 = left child), so the lanes in order are the tree in order and the curve's points are the lanes' points concatenated.
// A node on the way down that passes the leaf test belongs to the leftmost lane below it; the other lanes below it find nothing.  Every
// lane recomputes its up to five ancestors with the operations the serial walk performs on them: same floats, same leaf set.
__device__ void bez_lane_walk(PointSink &s, const float *e, uint32_t lane) {
    BezNode     cur = {e[0], e[1], e[2], e[3], e[4], e[5], e[6], e[7]};
    const float tol = e[8];
    int         pending = 0;
    for (int d = 0; d < 5; d++) {
        float x12 = (cur.x1 + cur.x2) / 2, y12 = (cur.y1 + cur.y2) / 2;
        float x23 = (cur.x2 + cur.x3) / 2, y23 = (cur.y2 + cur.y3) / 2;
        float x34 = (cur.x3 + cur.x4) / 2, y34 = (cur.y3 + cur.y4) / 2;
        float x123 = (x12 + x23) / 2, y123 = (y12 + y23) / 2;
        float x234 = (x23 + x34) / 2, y234 = (y23 + y34) / 2;
        float x1234 = (x123 + x234) / 2, y1234 = (y123 + y234) / 2;
        if (d > 0) {  // (level 0 always subdivides)
            float2    tmp[2];
            PointSink t;
            t.pts = nullptr; t.flags = nullptr; t.n = 0; t.flag = 0; t.emit = false; t.cache = tmp; t.n_cached = 0; t.cache_cap = 2;
            if (bez_leaf_test(t, tol, cur, x1234, y1234)) {
                if ((lane & ((1u << (5 - d)) - 1u)) == 0u) {
                    if (t.n_cached > 0) s.add(tmp[0].x, tmp[0].y);
                    if (t.n_cached > 1) s.add(tmp[1].x, tmp[1].y);
                }
                return;
            }
        }
        if ((lane >> (4 - d)) & 1u) cur = BezNode{x1234, y1234, x234, y234, x34, y34, cur.x4, cur.y4};
        else { cur = BezNode{cur.x1, cur.y1, x12, y12, x123, y123, x1234, y1234}; pending++; }
    }
    bez_dfs(s, cur, 5, pending, tol);
}

__device__ __forceinline__ void flatten_arc(PointSink &s, const float *e) {
    // interior points only; the host emits the start and end points as VKB_EL_POINT (it needs them as the
    // current point anyway).  Sequential float accumulation of the angle, as vkvg_context.c:425-435/478-488.
    const float xc = e[0], yc = e[1], radius = e[2], a2 = e[4], step = e[5];
    float       a = e[3];
    if (step > 0.f)
        while (a < a2) { s.add(cosf(a) * radius + xc, sinf(a) * radius + yc); a += step; }
    else
        while (a > a2) { s.add(cosf(a) * radius + xc, sinf(a) * radius + yc); a += step; }
}

template <bool EMIT>
__global__ void __launch_bounds__(128) flatten_k(const uint32_t *elem_hdr, const float *elem_data, uint32_t n_elems, uint32_t *counts,
                                                const uint32_t *offsets, float2 *pts, uint8_t *flags, const vkb_counts *C, float2 *cache, uint32_t cache_n) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_elems) return;
    if (EMIT && C->overflow) return;  // the points do not fit the buffers of this attempt (dev_util.cuh: vkb_counts)
    const uint32_t hdr = elem_hdr[i];
    const float   *e   = elem_data + (hdr >> VKB_EL_PAYLOAD_SHIFT);
    PointSink s;
    s.emit  = EMIT;
    s.n     = EMIT ? offsets[i] : 0;
    s.pts   = pts;
    s.flags = flags;
    s.flag  = (hdr & VKB_EL_CURVED) ? 1 : 0;
    s.cache = (!EMIT && cache && (hdr & VKB_EL_TYPE_MASK) != VKB_EL_POINT) ? cache + (size_t)i * cache_n : nullptr;
    s.n_cached = 0; s.cache_cap = cache_n;
    uint32_t start = s.n;
    if (EMIT && cache && (hdr & VKB_EL_TYPE_MASK) != VKB_EL_POINT) {
        const uint32_t cnt = (i + 1 < n_elems ? offsets[i + 1] : C->n[VKC_POINTS]) - start;
        if (cnt <= cache_n) {  // the counting pass already produced every point of this element
            const float2 *c = cache + (size_t)i * cache_n;
            for (uint32_t k = 0; k < cnt; k++) { pts[start + k] = c[k]; flags[start + k] = s.flag; }
            return;
        }
    }
    switch (hdr & VKB_EL_TYPE_MASK) {
    case VKB_EL_POINT: s.add(e[0], e[1]); break;
    case VKB_EL_CUBIC: flatten_cubic(s, e); break;
    case VKB_EL_ARC: flatten_arc(s, e); break;
    }
    if (!EMIT) counts[i] = s.n - start;
}

// Counting pass for batches of few elements (a tiger frame: 2500 cubics), where the pass lasts as long as the serial walk of the frame's
// longest curve (33 us of a 255 us frame): one WARP per element.  The lanes count their shares (bez_lane_walk), a prefix sum orders them,
// and - when the curve fits the cache - a second walk writes every lane's points where the emitting pass expects them.  Points and arcs are
// lane 0's, as in flatten_k.
__global__ void __launch_bounds__(128) flatten_count_warp_k(const uint32_t *elem_hdr, const float *elem_data, uint32_t n_elems, uint32_t *counts, float2 *cache, uint32_t cache_n) {
    const uint32_t i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31u;
    if (i >= n_elems) return;  // (whole warps)
    const uint32_t hdr = elem_hdr[i], type = hdr & VKB_EL_TYPE_MASK;
    const float   *e   = elem_data + (hdr >> VKB_EL_PAYLOAD_SHIFT);
    PointSink s;
    s.pts = nullptr; s.flags = nullptr; s.n = 0; s.flag = 0; s.emit = false; s.cache = nullptr; s.n_cached = 0; s.cache_cap = 0;
    if (type == VKB_EL_CUBIC) {
        bez_lane_walk(s, e, lane);
        const uint32_t n = s.n, incl = warp_incl_scan(n), total = __shfl_sync(0xffffffffu, incl, 31), off = incl - n;
        const uint32_t end_ok = (isnan(e[6]) || isnan(e[7])) ? 0u : 1u;  // (PointSink::add drops NaN)
        if (total + end_ok <= cache_n) {  // (a longer curve is walked again by the emitting pass: nothing of it is read from the cache)
            float2 *c = cache + (size_t)i * cache_n;
            if (n) {
                PointSink w;
                w.pts = nullptr; w.flags = nullptr; w.n = 0; w.flag = 0; w.emit = false; w.cache = c + off; w.n_cached = 0; w.cache_cap = n;
                bez_lane_walk(w, e, lane);
            }
            if (lane == 31 && end_ok) c[total] = make_float2(e[6], e[7]);
        }
        if (lane == 0) counts[i] = total + end_ok;
        return;
    }
    if (lane) return;
    s.cache = type != VKB_EL_POINT ? cache + (size_t)i * cache_n : nullptr;
    s.cache_cap = cache_n;
    if (type == VKB_EL_POINT) s.add(e[0], e[1]);
    else if (type == VKB_EL_ARC) flatten_arc(s, e);
    counts[i] = s.n;
}

// sub-path point ranges from the element offsets: first point = offset of the first element,
// count = offset(end) - offset(first) minus one if close_path dropped the duplicated last point
__global__ void subpath_ranges_k(const vkb_subpath *sps, uint32_t n_sp, const uint32_t *elem_off, uint32_t n_elems, const uint32_t *total,
                                 uint32_t *sp_first, uint32_t *sp_count) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_sp) return;
    vkb_subpath sp = sps[i];
    uint32_t    a  = sp.n_elems ? elem_off[sp.first_elem] : 0;
    uint32_t    e  = sp.first_elem + sp.n_elems;
    uint32_t    b  = sp.n_elems ? (e < n_elems ? elem_off[e] : *total) : 0;
    uint32_t    n  = b - a;
    if ((sp.flags & VKB_SP_DROP_LAST) && n > 0) n--;
    sp_first[i] = a;
    sp_count[i] = n;
}

void vkb_launch_flatten_count(const uint32_t *elem_hdr, const float *elem_data, uint32_t n, uint32_t *counts, float2 *cache, uint32_t cache_n, cudaStream_t s) {
    if (!n) return;
    static const bool per_thread = [] { const char *e = getenv("VKVG_B200_FLATTEN"); return e && e[0] == 't'; }();   // =thread: A/B against the kernel below
    if (cache && cache_n >= 64 && !per_thread) flatten_count_warp_k<<<vkb_div_up((uint64_t)n * 32, 128), 128, 0, s>>>(elem_hdr, elem_data, n, counts, cache, cache_n);  // (few elements: a warp each)
    else flatten_k<false><<<vkb_div_up(n, 128), 128, 0, s>>>(elem_hdr, elem_data, n, counts, nullptr, nullptr, nullptr, nullptr, cache, cache_n);
    VKB_LAUNCHED();
}
void vkb_launch_flatten_emit(const uint32_t *elem_hdr, const float *elem_data, uint32_t n, const uint32_t *offsets, float2 *pts, uint8_t *flags,
                             const vkb_counts *C, float2 *cache, uint32_t cache_n, cudaStream_t s) {
    if (!n) return;
    flatten_k<true><<<vkb_div_up(n, 128), 128, 0, s>>>(elem_hdr, elem_data, n, nullptr, offsets, pts, flags, C, cache, cache_n);
    VKB_LAUNCHED();
}
void vkb_launch_subpath_ranges(const vkb_subpath *sps, uint32_t n_sp, const uint32_t *elem_off, uint32_t n_elems, const uint32_t *total,
                               uint32_t *sp_first, uint32_t *sp_count, cudaStream_t s) {
    if (!n_sp) return;
    subpath_ranges_k<<<vkb_div_up(n_sp, 256), 256, 0, s>>>(sps, n_sp, elem_off, n_elems, total, sp_first, sp_count);
    VKB_LAUNCHED();
}
