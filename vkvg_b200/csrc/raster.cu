// Tile-binned winding rasteriser + paint/composite pass.
//
// Replaces the reference's stencil-then-cover GPU passes (pipelinePolyFill fan + scissored cover,
// src/vkvg_context_internal.c:1583-1655,1919-1952; direct blending of stroke / libtess triangles through
// pipe_OVER, src/vkvg_device_internal.c:200-246) and the ICD's rasterisation they rely on.
//
// Geometry reaches this file as directed edges in 24.8 fixed-point window coordinates:
//   even-odd / non-zero fills: the polygon edges of each sub-path (implicitly closed),
//   strokes: three edges per stroke triangle, every triangle oriented so that it winds +1,
//   paint: a rectangle larger than the surface.
// For a sample s the integer winding is  W(s) = sum_e sign(dy_e) [e spans s.y] [x_e(s.y) <= s.x]  with the
// sample displaced by (+eps', +eps), eps << eps' — exactly the top-left rule of the triangle rasteriser the
// reference runs on (oracle/vkvg_oracle.c: edge_winding / tri_edge_in).  Per 16x16 tile it is evaluated as
//   W(s) = B + V(s.y) + H(s)
//   B    = W at the virtual sample C = (X0+1/2, Y0+1/2) units of the tile origin   ("backdrop")
//   V    = signed crossings of the vertical segment from C down to (X0+1/2, s.y)
//   H    = signed crossings of the horizontal segment from (X0+1/2, s.y) to s
// B needs every edge of the draw and is accumulated by the binning kernel (atomic add at the first tile to
// the right of the crossing, then a prefix sum along each tile row); V and H only need edges that touch the
// tile, which the binning kernel scatters into per-tile lists.  All tests are exact (int64 products of
// 32-bit differences; half-unit coordinates are handled by doubling).
//
// The fine pass keeps the S per-sample colours of its pixel in registers across every draw that touches the
// tile, applies Porter-Duff OVER with the reference's UNORM8 rounding per sample, resolves (box filter) and
// writes each pixel once.
#include "pipeline.h"
#include <cstdlib>

#define VKB_EDGE_GRID (148u * 8u * 8u)  // most blocks of a per-edge / per-vertex / per-triangle kernel: eight waves of eight 256-thread blocks per SM, grid-stride beyond

// (vertex stage: vs_snap in pipeline.h - the stroke emitter runs it too)

__device__ __forceinline__ uint32_t find_job(const uint32_t *job_base, uint32_t n_jobs, uint32_t item) {
    uint32_t lo = 0, hi = n_jobs;
    while (hi - lo > 1) {
        uint32_t mid = (lo + hi) >> 1;
        if (job_base[mid] <= item) lo = mid; else hi = mid;
    }
    return lo;
}

// ---- bounding box of every sub-path in user space, from its ELEMENTS (a cubic lies inside the hull of its control points, an arc inside
//      the square about its circle): one warp per sub-path, lanes stride the elements.  What job_counts_k culls with. ----
// (boxes are kept as order-preserving ints so that the long sub-paths can be reduced with atomicMin / atomicMax)
__device__ __forceinline__ int32_t f2ord(float f) { const int32_t k = __float_as_int(f); return k >= 0 ? k : k ^ 0x7FFFFFFF; }
__device__ __forceinline__ float   ord2f(int32_t k) { return __int_as_float(k >= 0 ? k : k ^ 0x7FFFFFFF); }
__device__ __forceinline__ void elem_box(uint32_t h, const float *elem_data, float &x0, float &y0, float &x1, float &y1) {
    const float   *e = elem_data + (h >> VKB_EL_PAYLOAD_SHIFT);
    const uint32_t t = h & VKB_EL_TYPE_MASK;
    if (t == VKB_EL_ARC) {
        const float r = fabsf(e[2]);
        x0 = fminf(x0, e[0] - r); x1 = fmaxf(x1, e[0] + r); y0 = fminf(y0, e[1] - r); y1 = fmaxf(y1, e[1] + r);
    } else {
        const int np = t == VKB_EL_CUBIC ? 4 : 1;
        for (int k = 0; k < np; k++) { x0 = fminf(x0, e[2 * k]); x1 = fmaxf(x1, e[2 * k]); y0 = fminf(y0, e[2 * k + 1]); y1 = fmaxf(y1, e[2 * k + 1]); }
    }
}
__device__ __forceinline__ void warp_box(float &x0, float &y0, float &x1, float &y1) {
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        x0 = fminf(x0, __shfl_xor_sync(0xffffffffu, x0, o)); y0 = fminf(y0, __shfl_xor_sync(0xffffffffu, y0, o));
        x1 = fmaxf(x1, __shfl_xor_sync(0xffffffffu, x1, o)); y1 = fmaxf(y1, __shfl_xor_sync(0xffffffffu, y1, o));
    }
}
__global__ void __launch_bounds__(256)
sp_bounds_k(const vkb_subpath *sps, uint32_t n_sp, const uint32_t *elem_hdr, const float *elem_data, int4 *sp_bbox, uint32_t *long_blocks) {
    const uint32_t s = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (s >= n_sp) return;
    const vkb_subpath sp = sps[s];
    float x0 = 3.0e38f, y0 = 3.0e38f, x1 = -3.0e38f, y1 = -3.0e38f;
    if (sp.n_elems <= VKB_SP_LONG) {
        for (uint32_t i = lane; i < sp.n_elems; i += 32) elem_box(elem_hdr[sp.first_elem + i], elem_data, x0, y0, x1, y1);
        warp_box(x0, y0, x1, y1);
    }
    if (lane == 0) {
        sp_bbox[s]     = make_int4(f2ord(x0), f2ord(y0), f2ord(x1), f2ord(y1));  // (empty for the long ones: sp_bounds_long_k grows it)
        long_blocks[s] = sp.n_elems > VKB_SP_LONG ? (sp.n_elems + 255) / 256 : 0u;  // blocks of 256 elements the long ones are reduced in (scanned next)
    }
}
// the long sub-paths (a 1M-point polyline is ONE sub-path: a single warp would walk it for a millisecond): one block per 256 of their
// elements; first_block = exclusive scan of the per-sub-path block counts (zero for the short ones), *n_blocks its total
__global__ void __launch_bounds__(256)
sp_bounds_long_k(const vkb_subpath *sps, uint32_t n_sp, const uint32_t *first_block, const uint32_t *n_blocks, const uint32_t *elem_hdr, const float *elem_data, int4 *sp_bbox) {
    if (blockIdx.x >= *n_blocks) return;   // (the grid is sized for a bound the host can know: elements / 256 + elements / 1024)
    uint32_t lo = 0, hi = n_sp;  // the last sub-path whose first block is <= this one (short ones have none: they share the next one's)
    while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if (first_block[mid] <= blockIdx.x) lo = mid; else hi = mid;
    }
    const vkb_subpath sp = sps[lo];
    const uint32_t    i  = (blockIdx.x - first_block[lo]) * 256 + threadIdx.x;
    float x0 = 3.0e38f, y0 = 3.0e38f, x1 = -3.0e38f, y1 = -3.0e38f;
    if (i < sp.n_elems) elem_box(elem_hdr[sp.first_elem + i], elem_data, x0, y0, x1, y1);
    warp_box(x0, y0, x1, y1);
    __shared__ float red[4][8];
    if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = x0; red[1][threadIdx.x >> 5] = y0; red[2][threadIdx.x >> 5] = x1; red[3][threadIdx.x >> 5] = y1; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; w++) { x0 = fminf(x0, red[0][w]); y0 = fminf(y0, red[1][w]); x1 = fmaxf(x1, red[2][w]); y1 = fmaxf(y1, red[3][w]); }
        int32_t *b = (int32_t *)(sp_bbox + lo);
        atomicMin(b, f2ord(x0)); atomicMin(b + 1, f2ord(y0)); atomicMax(b + 2, f2ord(x1)); atomicMax(b + 3, f2ord(y1));
    }
}
void vkb_launch_sp_bounds(const vkb_subpath *sps, uint32_t n_sp, const uint32_t *elem_hdr, const float *elem_data, uint32_t n_elems, uint32_t *long_blocks, uint32_t *n_long_blocks,
                          ScanScratch &scan, int4 *sp_bbox, bool any_long, cudaStream_t s) {
    if (!n_sp) return;
    sp_bounds_k<<<vkb_div_up((uint64_t)n_sp * 32, 256), 256, 0, s>>>(sps, n_sp, elem_hdr, elem_data, sp_bbox, long_blocks);
    VKB_LAUNCHED();
    if (any_long && n_elems > VKB_SP_LONG) {
        vkb_exclusive_scan<uint32_t, uint32_t>(long_blocks, long_blocks, n_sp, n_long_blocks, scan, s);
        sp_bounds_long_k<<<n_elems / 256 + n_elems / VKB_SP_LONG + 1, 256, 0, s>>>(sps, n_sp, long_blocks, n_long_blocks, elem_hdr, elem_data, sp_bbox);
        VKB_LAUNCHED();
    }
}
// Work items of a job = the points of its sub-path, or none when the sub-path is too short - or when it cannot touch the surface: a
// sub-path is a closed curve of its own for fills and clips (and a stroke stays within its half width / miter length of the path), so
// one whose box lies wholly above, below, left or right of the surface leaves every sample's winding as it is.  On a stripe surface
// (multi-GPU tile rows) that removes most of the scene before anything is tessellated, snapped or binned.
__device__ __forceinline__ void job_count_one(uint32_t j, const uint32_t *job_sp, const uint32_t *job_draw, const uint32_t *sp_count, uint32_t min_points, const int4 *sp_bbox,
                                              const vkb_draw *draws, const vkb_xform *xforms, const vkb_stroke *strokes, const SurfaceDesc &sd, uint32_t *job_n) {
    const uint32_t s = job_sp[j];
    uint32_t       n = sp_count[s];
    if (n < min_points) n = 0;
    if (n && sp_bbox) {
        const vkb_draw  &dr = draws[job_draw[j]];
        const vkb_xform &xf = xforms[dr.xform_stroke & 0xFFFF];
        const float     *m  = xf.mat;
        const int4       bi = sp_bbox[s];
        const float4     b  = make_float4(ord2f(bi.x), ord2f(bi.y), ord2f(bi.z), ord2f(bi.w));
        float            ext = 2.0f;  // (vertex-stage rounding, 1/256 snap, sample offsets: far below one pixel; two for good measure)
        if (dr.kind == VKB_DRAW_STROKE) {
            const vkb_stroke &st = strokes[dr.xform_stroke >> 16];
            ext += fmaxf(st.lhMax, 2.0f * st.hw) * sqrtf(m[0] * m[0] + m[1] * m[1] + m[2] * m[2] + m[3] * m[3]);
        }
        float xlo = 3.0e38f, ylo = 3.0e38f, xhi = -3.0e38f, yhi = -3.0e38f;
#pragma unroll
        for (int c = 0; c < 4; c++) {
            const float x = (c & 1) ? b.z : b.x, y = (c & 2) ? b.w : b.y;
            const float px = m[0] * x + m[2] * y + m[4], py = m[1] * x + m[3] * y + m[5];
            xlo = fminf(xlo, px); xhi = fmaxf(xhi, px); ylo = fminf(ylo, py); yhi = fmaxf(yhi, py);
        }
        const float yoff = (float)(xf.band * sd.band_tiles * VKB_TILE) - (float)sd.origin_y;  // canvas of a batch surface / stripe of a taller one
        ylo += yoff; yhi += yoff;
        // (comparisons written so that NaN / inf boxes are kept)
        if (yhi + ext < 0.0f || ylo - ext > (float)sd.height || xhi + ext < 0.0f || xlo - ext > (float)sd.width) n = 0;
    }
    job_n[j] = n;
}
__global__ void job_counts_k(const uint32_t *job_sp, const uint32_t *job_draw, uint32_t n_jobs, const uint32_t *sp_count, uint32_t min_points, const int4 *sp_bbox,
                             const vkb_draw *draws, const vkb_xform *xforms, const vkb_stroke *strokes, SurfaceDesc sd, uint32_t *job_n) {
    uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < n_jobs) job_count_one(j, job_sp, job_draw, sp_count, min_points, sp_bbox, draws, xforms, strokes, sd, job_n);
}
// both job tables of a batch that fills AND strokes (a tiger): threads [0, nf) size the fill jobs (> 2 points), the next ns the stroke jobs (>= 2)
__global__ void job_counts2_k(const uint32_t *fjob_sp, const uint32_t *fjob_draw, uint32_t nf, uint32_t *fjob_n, const uint32_t *sjob_sp, const uint32_t *sjob_draw, uint32_t ns,
                              uint32_t *sjob_n, const uint32_t *sp_count, const int4 *sp_bbox, const vkb_draw *draws, const vkb_xform *xforms, const vkb_stroke *strokes,
                              SurfaceDesc sd) {
    uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < nf) job_count_one(j, fjob_sp, fjob_draw, sp_count, 3, sp_bbox, draws, xforms, strokes, sd, fjob_n);
    else if (j - nf < ns) job_count_one(j - nf, sjob_sp, sjob_draw, sp_count, 2, sp_bbox, draws, xforms, strokes, sd, sjob_n);
}
void vkb_launch_job_counts2(const uint32_t *fjob_sp, const uint32_t *fjob_draw, uint32_t nf, uint32_t *fjob_n, const uint32_t *sjob_sp, const uint32_t *sjob_draw, uint32_t ns,
                            uint32_t *sjob_n, const uint32_t *sp_count, const int4 *sp_bbox, const vkb_draw *draws, const vkb_xform *xforms, const vkb_stroke *strokes,
                            SurfaceDesc sd, cudaStream_t s) {
    job_counts2_k<<<vkb_div_up((uint64_t)nf + ns, 256), 256, 0, s>>>(fjob_sp, fjob_draw, nf, fjob_n, sjob_sp, sjob_draw, ns, sjob_n, sp_count, sp_bbox, draws, xforms, strokes, sd);
    VKB_LAUNCHED();
}
void vkb_launch_job_counts(const uint32_t *job_sp, const uint32_t *job_draw, uint32_t n_jobs, const uint32_t *sp_count, uint32_t min_points, const int4 *sp_bbox,
                           const vkb_draw *draws, const vkb_xform *xforms, const vkb_stroke *strokes, SurfaceDesc sd, uint32_t *job_n, cudaStream_t s) {
    if (!n_jobs) return;
    job_counts_k<<<vkb_div_up(n_jobs, 256), 256, 0, s>>>(job_sp, job_draw, n_jobs, sp_count, min_points, sp_bbox, draws, xforms, strokes, sd, job_n);
    VKB_LAUNCHED();
}

// An edge wholly above, below or to the right of the surface changes no sample of it: the winding of a sample only counts edges that
// span its height and cross at or left of it.  (Wholly to the LEFT does count: it is part of every backdrop of its rows.)  On a stripe
// surface (multi-GPU tile rows) this removes most of the scene before binning looks at it.
__device__ __forceinline__ bool edge_off_surface(const vkb_edge &e, const SurfaceDesc &sd) {
    return max(e.y0, e.y1) < 0 || min(e.y0, e.y1) > (int32_t)sd.height * 256 || min(e.x0, e.x1) > (int32_t)sd.width * 256;
}
// ---- per-draw bounding boxes of the stored edges, grown by the kernels that EMIT the edges (a pass of its own over 25 M edges cost
//      0.87 ms on C5a).  A warp whose edges all belong to one draw reduces with redux.sync and issues at most four atomics; every tier looks
//      at the box first: a stale read only costs a redundant atomic, never a missed one (the box only grows). ----
__device__ __forceinline__ bool edge_degenerate(const vkb_edge &e) { return e.x0 == e.x1 && e.y0 == e.y1; }
__device__ __forceinline__ void bbox_grow(int32_t *bbox, uint32_t d, int32_t mnx, int32_t mny, int32_t mxx, int32_t mxy) {
    if (mnx > mxx) return;
    volatile int32_t *vb = bbox + 4 * (size_t)d;
    if (mnx < vb[0]) atomicMin(&bbox[4 * (size_t)d], mnx);
    if (mny < vb[1]) atomicMin(&bbox[4 * (size_t)d + 1], mny);
    if (mxx > vb[2]) atomicMax(&bbox[4 * (size_t)d + 2], mxx);
    if (mxy > vb[3]) atomicMax(&bbox[4 * (size_t)d + 3], mxy);
}
struct BoxAcc {  // what one thread contributes: the box of the (non-degenerate) edges it stored for draw d
    int32_t mnx = INT32_MAX, mny = INT32_MAX, mxx = INT32_MIN, mxy = INT32_MIN;
    __device__ __forceinline__ void add(const vkb_edge &e) {
        if (edge_degenerate(e)) return;
        mnx = min(mnx, min(e.x0, e.x1)); mny = min(mny, min(e.y0, e.y1)); mxx = max(mxx, max(e.x0, e.x1)); mxy = max(mxy, max(e.y0, e.y1));
    }
};
// every lane of the warp must call this (lanes without edges pass an empty box and any d)
__device__ __forceinline__ void bbox_accumulate(int32_t *bbox, uint32_t d, BoxAcc b) {
    const bool     has = b.mnx <= b.mxx;
    const uint32_t hm  = __ballot_sync(0xffffffffu, has);
    if (!hm) return;
    const uint32_t d0 = __shfl_sync(0xffffffffu, d, __ffs((int)hm) - 1);
    if (__all_sync(0xffffffffu, !has || d == d0)) {
        b.mnx = __reduce_min_sync(0xffffffffu, b.mnx); b.mny = __reduce_min_sync(0xffffffffu, b.mny);
        b.mxx = __reduce_max_sync(0xffffffffu, b.mxx); b.mxy = __reduce_max_sync(0xffffffffu, b.mxy);
        if ((threadIdx.x & 31) == 0) bbox_grow(bbox, d0, b.mnx, b.mny, b.mxx, b.mxy);
    } else if (has) bbox_grow(bbox, d, b.mnx, b.mny, b.mxx, b.mxy);
}

// ---- fill: one edge per point of every sub-path with > 2 points (the fan of _poly_fill covers exactly the
//      implicitly closed polygon, internal.c:1617-1642) ----
struct FillItem {
    uint32_t j, k, first, n, d;
    float2   a, b;
};
__device__ __forceinline__ FillItem fill_item(uint32_t item, const float2 *pts, const uint32_t *job_draw, const uint32_t *job_sp, const uint32_t *job_base,
                                              uint32_t n_jobs, const uint32_t *sp_first, const uint32_t *sp_count) {
    FillItem f;
    f.j = find_job(job_base, n_jobs, item); f.k = item - job_base[f.j];
    const uint32_t s = job_sp[f.j];
    f.first = sp_first[s]; f.n = sp_count[s]; f.d = job_draw[f.j];
    f.a = pts[f.first + f.k]; f.b = pts[f.first + (f.k + 1 == f.n ? 0 : f.k + 1)];
    return f;
}
__device__ __forceinline__ vkb_edge fill_snap_edge(const vkb_xform &xf, const SurfaceDesc &sd, float2 a, float2 b) {
    vkb_edge e;
    vs_snap(xf.mat, (float)sd.width, (float)sd.full_height, a.x, a.y, e.x0, e.y0);
    vs_snap(xf.mat, (float)sd.width, (float)sd.full_height, b.x, b.y, e.x1, e.y1);
    const int32_t yoff = (int32_t)(xf.band * sd.band_tiles) * VKB_TILE_FX - (int32_t)sd.origin_y * 256;
    e.y0 += yoff; e.y1 += yoff;
    if (edge_off_surface(e, sd)) e = vkb_edge{0, 0, 0, 0};  // (a degenerate edge: every later stage skips it)
    return e;
}
__global__ void __launch_bounds__(256)
fill_edges_k(const float2 *pts, const vkb_draw *draws, const vkb_xform *xforms, const uint32_t *job_draw, const uint32_t *job_sp, const uint32_t *job_base, uint32_t n_jobs,
             const uint32_t *sp_first, const uint32_t *sp_count, const vkb_counts *C, SurfaceDesc sd, vkb_edge *edges, uint32_t *edge_draw, int32_t *bbox) {
    uint32_t item = blockIdx.x * blockDim.x + threadIdx.x;
    if (C->overflow) return;
    const uint32_t n_items = C->n[VKC_FILL];
    if ((blockIdx.x * blockDim.x + (threadIdx.x & ~31u)) >= n_items) return;  // whole warps leave; a partial warp stays for bbox_accumulate
    BoxAcc   box;
    uint32_t d = 0;
    if (item < n_items) {
        const FillItem f = fill_item(item, pts, job_draw, job_sp, job_base, n_jobs, sp_first, sp_count);
        const vkb_edge e = fill_snap_edge(xforms[draws[f.d].xform_stroke & 0xFFFF], sd, f.a, f.b);
        edges[item]     = e;
        edge_draw[item] = d = f.d;
        box.add(e);
    }
    bbox_accumulate(bbox, d, box);
}

// ---- NON_ZERO fills and clips the way the reference's libtess makes them (src/vkvg_context_internal.c:1720-1793, external/glutess) ----
// The reference does not evaluate a winding rule: it blends the triangles libtess returns.  Two regimes decide what those cover:
//  * one contour (sub-path of > 2 points) of at most 100 vertices whose fan about the first vertex turns one way throughout:
//    libtess's cache fast path (tess.c:376-443, render.c:362-516 __gl_renderCache) emits that fan as it is - original vertices, and
//    where the polygon winds twice the triangles overlap and the reference blends twice.  Here: the polygon's own edges with the COUNT
//    rule (the fan's spokes cancel in the winding sum, what is left is |winding| blends) - mode 1;
//  * everything else goes through the sweep, which adds a vertex wherever two edges cross; combine2 (:1706-1712) stores it as a FLOAT
//    vec2 and it reaches the rasteriser on the 1/256 grid like any vertex, so the non-overlapping triangles tile { winding != 0 } of
//    the path whose edges are BENT at those vertices.  Here: every edge is split at its proper crossings with the other edges of the
//    draw (crossing in double, as libtess computes, rounded to float) and the NON_ZERO rule runs on the pieces - mode 2.  Draws of more
//    than VKB_NZ_SPLIT_MAX edges are not split (the search is quadratic) - mode 3; the difference is then single samples next to
//    self-intersections (measured on C2 without any of this: 0.43 % of the pixels, p99.9 = 11/255).
// Same arithmetic as nz_fan / nz_cross / nz_edges in oracle/vkvg_oracle.c.
#define VKB_NZ_SPLIT_MAX 1024
__global__ void __launch_bounds__(128)
nz_classify_k(const vkb_draw *draws, uint32_t n_draws, const uint32_t *sp_first, const uint32_t *sp_count, const float2 *pts, const vkb_counts *C,
              vkb_paint *paints, uint8_t *nz_mode) {
    const uint32_t d = blockIdx.x * blockDim.x + threadIdx.x;
    if (d >= n_draws || C->overflow) return;
    const vkb_draw dr = draws[d];
    const uint32_t rule = dr.rule_pattern & 0xFF;
    uint8_t        mode = 0;
    if ((dr.kind == VKB_DRAW_FILL && rule == VKB_RULE_NON_ZERO) || (dr.kind == VKB_DRAW_CLIP && rule == VKB_RULE_CLIP_NZ)) {
        uint32_t contours = 0, first = 0, n = 0, items = 0;
        for (uint32_t s = dr.first_subpath; s < dr.first_subpath + dr.n_subpaths; s++) {
            const uint32_t c = sp_count[s];
            if (c > 2) {
                if (contours++ == 0) { first = sp_first[s]; n = c; }
                items += c;
            }
        }
        mode = items > VKB_NZ_SPLIT_MAX ? 3 : 2;
        if (contours == 1 && n <= 100) {
            const float2 *p = pts + first;
            double norm = 0.0;
            int    sign = 0;
            bool   consistent = true;
            for (int pass = 0; pass < 2 && consistent; pass++) {
                double xc = (double)p[1].x - (double)p[0].x, yc = (double)p[1].y - (double)p[0].y;
                for (uint32_t k = 2; k < n; k++) {
                    const double xp = xc, yp = yc;
                    xc = (double)p[k].x - (double)p[0].x; yc = (double)p[k].y - (double)p[0].y;
                    const double nz = xp * yc - yp * xc, dot = nz * norm;
                    if (pass == 0) { if (dot >= 0) norm += nz; else norm -= nz; }
                    else if (dot != 0) {
                        if (dot > 0) { if (sign < 0) { consistent = false; break; } sign = 1; }
                        else { if (sign > 0) { consistent = false; break; } sign = -1; }
                    }
                }
            }
            if (consistent) mode = 1;
        }
    }
    nz_mode[d] = mode;
    if (dr.kind == VKB_DRAW_FILL && rule == VKB_RULE_NON_ZERO)  // (every flush: the paint table outlives the flush on resident replays)
        paints[d].rule_pattern = (dr.rule_pattern & ~0xFFu) | (mode == 1 ? VKB_RULE_COUNT : VKB_RULE_NON_ZERO);
}
// proper crossing of A = a -> b with B = c -> d, A the edge that comes first in the draw: decided without a division (t = tn / den and
// u = un / den strictly inside (0, 1)); nz_cross_point then gives the parameters along both and the point as a float
struct NzPair { double rx, ry, tn, un, den; };
__device__ __forceinline__ bool nz_cross(float2 a, float2 b, float2 c, float2 d, NzPair &o) {
    const double rx = (double)b.x - (double)a.x, ry = (double)b.y - (double)a.y, sx = (double)d.x - (double)c.x, sy = (double)d.y - (double)c.y;
    const double den = rx * sy - ry * sx;
    if (den == 0.0) return false;
    const double qx = (double)c.x - (double)a.x, qy = (double)c.y - (double)a.y;
    const double tn = qx * sy - qy * sx, un = qx * ry - qy * rx;
    if (den > 0.0 ? !(tn > 0.0 && tn < den && un > 0.0 && un < den) : !(tn < 0.0 && tn > den && un < 0.0 && un > den)) return false;
    o.rx = rx; o.ry = ry; o.tn = tn; o.un = un; o.den = den;
    return true;
}
struct NzHit { double key; uint32_t other; float2 p; };
__device__ __forceinline__ bool nz_hit_less(const NzHit &x, const NzHit &y) { return x.key < y.key || (x.key == y.key && x.other < y.other); }
// visits the crossings of fill item `item` (edge f.a -> f.b of draw f.d) with every other edge of the draw: visit(other, pair, item_is_first)
template <class F>
__device__ __forceinline__ void nz_for_each_crossing(uint32_t item_unused, const FillItem &f, const float2 *pts, const uint32_t *job_sp, const uint32_t *job_base, uint32_t n_jobs,
                                                     const uint32_t *sp_first, const uint32_t *sp_count, const uint32_t *draw_first_job, uint32_t n_draws, F &&visit) {
    const uint32_t j0 = draw_first_job[f.d], j1 = f.d + 1 < n_draws ? draw_first_job[f.d + 1] : n_jobs;
    const uint32_t item = f.first + f.k;
    (void)item_unused; (void)job_base;
    const float    lox = fminf(f.a.x, f.b.x), hix = fmaxf(f.a.x, f.b.x), loy = fminf(f.a.y, f.b.y), hiy = fmaxf(f.a.y, f.b.y);
    for (uint32_t jj = j0; jj < j1; jj++) {
        const uint32_t s2 = job_sp[jj], n2 = sp_count[s2];
        if (n2 < 3) continue;
        const uint32_t base = sp_first[s2];   // edges are identified by the index of their first point: the same on every stripe / canvas
        const float2  *q = pts + base;        // (the work items of a culled sub-path are gone, its edges still split the others)
        float2         c = __ldg(q);
        for (uint32_t k2 = 0; k2 < n2; k2++) {
            const float2   dd = __ldg(q + (k2 + 1 == n2 ? 0 : k2 + 1));
            const uint32_t other = base + k2;
            // (boxes that do not overlap cannot cross: the same early-out as the oracle's, it never changes a decision)
            if (other != item && !(hix < fminf(c.x, dd.x) || fmaxf(c.x, dd.x) < lox || hiy < fminf(c.y, dd.y) || fmaxf(c.y, dd.y) < loy)) {
                NzPair pr;
                const bool first = item < other;
                if (first ? nz_cross(f.a, f.b, c, dd, pr) : nz_cross(c, dd, f.a, f.b, pr)) visit(other, pr, first, c);
            }
            c = dd;
        }
    }
}
// parameter along THIS item's edge and the crossing point (computed from the edge that comes first, so both edges get the same float)
__device__ __forceinline__ NzHit nz_hit(uint32_t other, const NzPair &pr, bool first, float2 a_first) {
    NzHit h;
    const double t = pr.tn / pr.den;
    h.key   = first ? t : pr.un / pr.den;
    h.other = other;
    h.p.x = (float)((double)a_first.x + t * pr.rx);
    h.p.y = (float)((double)a_first.y + t * pr.ry);
    return h;
}
// One pass: every fill item writes its pieces (one edge, or 1 + its crossings when its draw is split) where its warp reserved room with a
// single atomic on *n_out - the order of the edges of a draw is immaterial (windings and backdrops are sums).  Nothing is written past
// `cap`; commit_fedges_k then checks the total against it (the host replays the batch with room when it did not fit).
#define VKB_NZ_LOCAL 12
__global__ void __launch_bounds__(128)
nz_split_k(const float2 *pts, const vkb_draw *draws, const vkb_xform *xforms, const uint32_t *job_draw, const uint32_t *job_sp, const uint32_t *job_base,
           uint32_t n_jobs, const uint32_t *sp_first, const uint32_t *sp_count, const uint32_t *draw_first_job, uint32_t n_draws, const uint8_t *nz_mode,
           const vkb_counts *C, SurfaceDesc sd, vkb_edge *edges, uint32_t *edge_draw, uint32_t *n_out, uint32_t cap, int32_t *bbox) {
    if (C->overflow) return;
    const uint32_t n_items = C->n[VKC_FILL];
    // whole warps stride the LIVE items (the grid comes from the capacity of the item space, capped: on a stripe most of the scene's sub-paths
    // were culled and have no items); a partial warp stays for the shuffles below
    for (uint32_t wb = blockIdx.x * blockDim.x + (threadIdx.x & ~31u); wb < n_items; wb += gridDim.x * blockDim.x) {
    const uint32_t item = wb + (threadIdx.x & 31u);
    const bool live = item < n_items;
    FillItem   f;
    NzHit      loc[VKB_NZ_LOCAL];
    uint32_t   m = 0;
    bool       split = false;
    if (live) {
        f     = fill_item(item, pts, job_draw, job_sp, job_base, n_jobs, sp_first, sp_count);
        split = nz_mode[f.d] == 2;
        if (split)
            nz_for_each_crossing(item, f, pts, job_sp, job_base, n_jobs, sp_first, sp_count, draw_first_job, n_draws, [&](uint32_t other, const NzPair &pr, bool first, float2 c) {
                if (m < VKB_NZ_LOCAL) {  // insertion sort among the first few (what nearly every edge has)
                    const NzHit h = nz_hit(other, pr, first, first ? f.a : c);
                    uint32_t i = m;
                    while (i > 0 && nz_hit_less(h, loc[i - 1])) { loc[i] = loc[i - 1]; i--; }
                    loc[i] = h;
                }
                m++;
            });
    }
    const uint32_t cnt  = live ? m + 1 : 0;
    const uint32_t incl = warp_incl_scan(cnt);
    uint32_t       base = 0;
    if ((threadIdx.x & 31) == 31 && incl) base = atomicAdd(n_out, incl);
    uint32_t o = __shfl_sync(0xffffffffu, base, 31) + incl - cnt;
    BoxAcc   box;
    if (!live) { bbox_accumulate(bbox, 0, box); continue; }
    const vkb_xform &xf = xforms[draws[f.d].xform_stroke & 0xFFFF];
    float2           prev = f.a;
    auto put = [&](float2 to) {
        const vkb_edge e = fill_snap_edge(xf, sd, prev, to);
        if (o < cap) { edges[o] = e; edge_draw[o] = f.d; }
        box.add(e);
        o++;
        prev = to;
    };
    if (m <= VKB_NZ_LOCAL) {
        for (uint32_t i = 0; i < m; i++) put(loc[i].p);
    } else {  // more crossings than fit: repeated selection of the next one in order
        NzHit last;
        last.key = -1.0; last.other = 0;
        for (uint32_t i = 0; i < m; i++) {
            NzHit best;
            best.key = 2.0; best.other = 0xffffffffu; best.p = f.b;
            nz_for_each_crossing(item, f, pts, job_sp, job_base, n_jobs, sp_first, sp_count, draw_first_job, n_draws, [&](uint32_t other, const NzPair &pr, bool first, float2 c) {
                const NzHit h = nz_hit(other, pr, first, first ? f.a : c);
                if (nz_hit_less(last, h) && nz_hit_less(h, best)) best = h;
            });
            put(best.p);
            last = best;
        }
    }
    put(f.b);
    bbox_accumulate(bbox, f.d, box);
    }
}
__global__ void commit_fedges_k(vkb_counts *C, const uint32_t *n_out) {
    if (C->overflow) return;
    vkc_commit(C, VKC_FEDGES, *n_out);
}
void vkb_launch_fill_edges(const float2 *pts, const vkb_draw *draws, const vkb_xform *xforms, const uint32_t *job_draw, const uint32_t *job_sp, const uint32_t *job_base,
                           uint32_t n_jobs, const uint32_t *sp_first, const uint32_t *sp_count, uint32_t cap_items, const vkb_counts *C, SurfaceDesc sd,
                           vkb_edge *edges, uint32_t *edge_draw, int32_t *draw_bbox, cudaStream_t s) {
    if (!cap_items || !n_jobs) return;
    fill_edges_k<<<vkb_div_up(cap_items, 256), 256, 0, s>>>(pts, draws, xforms, job_draw, job_sp, job_base, n_jobs, sp_first, sp_count, C, sd, edges, edge_draw, draw_bbox);
    VKB_LAUNCHED();
}
void vkb_launch_nz_classify(const vkb_draw *draws, uint32_t n_draws, const uint32_t *sp_first, const uint32_t *sp_count, const float2 *pts, const vkb_counts *C,
                            vkb_paint *paints, uint8_t *nz_mode, cudaStream_t s) {
    if (!n_draws) return;
    nz_classify_k<<<vkb_div_up(n_draws, 128), 128, 0, s>>>(draws, n_draws, sp_first, sp_count, pts, C, paints, nz_mode);
    VKB_LAUNCHED();
}
void vkb_launch_nz_split(const float2 *pts, const vkb_draw *draws, const vkb_xform *xforms, const uint32_t *job_draw, const uint32_t *job_sp, const uint32_t *job_base,
                         uint32_t n_jobs, const uint32_t *sp_first, const uint32_t *sp_count, const uint32_t *draw_first_job, uint32_t n_draws, const uint8_t *nz_mode,
                         uint32_t cap_items, vkb_counts *C, SurfaceDesc sd, vkb_edge *edges, uint32_t *edge_draw, uint32_t *n_out, uint32_t cap_edges, int32_t *draw_bbox, cudaStream_t s) {
    if (!cap_items || !n_jobs) return;
    nz_split_k<<<min(vkb_div_up(cap_items, 128), VKB_EDGE_GRID * 2u), 128, 0, s>>>(pts, draws, xforms, job_draw, job_sp, job_base, n_jobs, sp_first, sp_count, draw_first_job, n_draws, nz_mode, C, sd,
                                                          edges, edge_draw, n_out, cap_edges, draw_bbox);
    VKB_LAUNCHED();
    commit_fedges_k<<<1, 1, 0, s>>>(C, n_out);
    VKB_LAUNCHED();
}

// ---- stroke: three edges per triangle, oriented to wind +1 (so the sum over triangles is the number of
//      triangles covering the sample, which is how many times the reference blends it).
//      An edge shared by two consecutive triangles of the index list that traverse it in opposite directions
//      contributes exactly zero winding to every sample (quad diagonals, fan spokes, the seam between a quad and the
//      next join): both copies are dropped here, which roughly halves what the rasteriser has to look at. ----
struct TriV {
    uint32_t i[3];
    int32_t  x[3], y[3];
    int      sign;  // +1: index order is kept, -1: reversed, 0: degenerate / invalid (emits nothing)
};
__device__ __forceinline__ bool tri_has(const uint32_t (&i)[3], uint32_t u, uint32_t v) {
    return (i[0] == u || i[1] == u || i[2] == u) && (i[0] == v || i[1] == v || i[2] == v);
}
// +1 if u is followed by v in the cyclic index order of i, -1 if v is followed by u
__device__ __forceinline__ int tri_dir(const uint32_t (&i)[3], uint32_t u, uint32_t v) {
    return ((i[0] == u && i[1] == v) || (i[1] == u && i[2] == v) || (i[2] == u && i[0] == v)) ? 1 : -1;
}
__device__ __forceinline__ void tri_idx(const uint32_t *inds, uint32_t n_tris, long long t, uint32_t (&i)[3]) {
    if (t < 0 || t >= (long long)n_tris) { i[0] = i[1] = i[2] = 0xffffffffu; return; }
    i[0] = inds[3 * t]; i[1] = inds[3 * t + 1]; i[2] = inds[3 * t + 2];
}
// every stroke vertex through the vertex stage once (a vertex is shared by up to six triangles, and tri_edges_k also looks at
// the two neighbouring triangles of each one: snapping inside it cost nine vertex transforms per triangle)
__global__ void __launch_bounds__(256)
snap_verts_k(const float2 *verts, const vkb_counts *C, const vkb_draw *draws, const vkb_xform *xforms, const uint32_t *sdraw_id,
             const uint32_t *sdraw_first_item, uint32_t n_sdraws, const unsigned long long *item_offsets, SurfaceDesc sd, int2 *snapped) {
    if (C->overflow) return;
    const uint32_t n_verts = C->n[VKC_VERTS];
    for (uint32_t v = blockIdx.x * blockDim.x + threadIdx.x; v < n_verts; v += gridDim.x * blockDim.x) {  // (grid from the capacity, capped)
        uint32_t lo = 0, hi = n_sdraws;  // stroke draw owning vertex v: last q whose first item's vertex offset <= v
        while (hi - lo > 1) {
            uint32_t mid = (lo + hi) >> 1;
            if ((uint32_t)(item_offsets[sdraw_first_item[mid]] & 0xffffffffull) <= v) lo = mid; else hi = mid;
        }
        const vkb_xform &xf = xforms[draws[sdraw_id[lo]].xform_stroke & 0xFFFF];
        const float2     p  = verts[v];
        int32_t          x, y;
        vs_snap(xf.mat, (float)sd.width, (float)sd.full_height, p.x, p.y, x, y);
        snapped[v] = make_int2(x, y + (int32_t)(xf.band * sd.band_tiles) * VKB_TILE_FX - (int32_t)sd.origin_y * 256);
    }
}
__device__ __forceinline__ int tri_sign(const int2 *snapped, uint32_t n_verts, const uint32_t (&i)[3], int32_t (&x)[3], int32_t (&y)[3]) {
    if (i[0] >= n_verts || i[1] >= n_verts || i[2] >= n_verts) return 0;
#pragma unroll
    for (int k = 0; k < 3; k++) {
        const int2 p = snapped[i[k]];
        x[k] = p.x; y[k] = p.y;
    }
    long long area = (long long)(x[1] - x[0]) * (y[2] - y[0]) - (long long)(x[2] - x[0]) * (y[1] - y[0]);
    return area > 0 ? -1 : (area < 0 ? 1 : 0);  // cross > 0 winds -1 under our convention: such a triangle is reversed
}
// Only the edges that survive are stored: every warp reserves room for its live edges with one atomic on *live (the order of
// the edges of a draw is immaterial: windings and backdrops are sums), so the binning kernels never see the ~60 % of slots that
// cancelled.  Layout of the edge array: fill edges, whole-surface rectangles (n_extra), then the live stroke edges.
__global__ void __launch_bounds__(256)
tri_edges_k(const int2 *snapped, const uint32_t *inds, const vkb_counts *C, const uint32_t *sdraw_id, const uint32_t *sdraw_first_item,
            uint32_t n_sdraws, const unsigned long long *item_offsets, vkb_edge *edges, uint32_t *edge_draw, uint32_t n_extra, uint32_t *live,
            SurfaceDesc sd, int32_t *bbox) {
    if (C->overflow) return;
    const uint32_t n_tris = C->n[VKC_TRIS], n_verts = C->n[VKC_VERTS];
    edges += C->n[VKC_FEDGES] + n_extra; edge_draw += C->n[VKC_FEDGES] + n_extra;
    // whole warps stride the LIVE triangles (the grid comes from the capacity of the index buffer, capped); a partial warp stays for the shuffles
    for (uint32_t wb = blockIdx.x * blockDim.x + (threadIdx.x & ~31u); wb < n_tris; wb += gridDim.x * blockDim.x) {
        uint32_t   t = wb + (threadIdx.x & 31u);
        const bool in_range = t < n_tris;
        if (!in_range) t = n_tris - 1;
        uint32_t i0[3];
        tri_idx(inds, n_tris, (long long)t, i0);
        int32_t x[3], y[3], nx[3], ny[3];
        int     sg = tri_sign(snapped, n_verts, i0, x, y);
        // a triangle wholly above, below or right of the surface: each of its edges is (edge_off_surface), so none of them is stored whether
        // or not a neighbour cancels it - skip the neighbours.  On a stripe of a taller surface that is most of the triangles.
        if (sg != 0 && (max(max(y[0], y[1]), y[2]) < 0 || min(min(y[0], y[1]), y[2]) > (int32_t)sd.height * 256 ||
                        min(min(x[0], x[1]), x[2]) > (int32_t)sd.width * 256))
            sg = 0;
        vkb_edge e[3];
        uint32_t d = 0;
#pragma unroll
        for (int k = 0; k < 3; k++) e[k] = vkb_edge{0, 0, 0, 0};
        if (sg != 0) {
            // stroke draw owning index 3t: last q whose first item's index offset <= 3t
            uint32_t lo = 0, hi = n_sdraws;
            while (hi - lo > 1) {
                uint32_t mid = (lo + hi) >> 1;
                if ((uint32_t)(item_offsets[sdraw_first_item[mid]] >> 32) <= 3 * t) lo = mid; else hi = mid;
            }
            d = sdraw_id[lo];
            uint32_t ip1[3], ip2[3], im1[3], im2[3];
            tri_idx(inds, n_tris, (long long)t + 1, ip1);
            tri_idx(inds, n_tris, (long long)t + 2, ip2);
            tri_idx(inds, n_tris, (long long)t - 1, im1);
            tri_idx(inds, n_tris, (long long)t - 2, im2);
            int sg_next = 2, sg_prev = 2;  // 2: not evaluated yet (vertices of a neighbour share this draw's matrix: shared indices)
#pragma unroll
            for (int k = 0; k < 3; k++) {
                // k-th edge in normalised orientation: kept order a->b, b->c, c->a; reversed a->c, c->b, b->a
                const int ka = sg > 0 ? k : (3 - k) % 3, kb = sg > 0 ? (k + 1) % 3 : (5 - k) % 3;
                const uint32_t u = i0[ka], v = i0[kb];
                bool           drop = false;
                if (tri_has(ip1, u, v)) {
                    if (!tri_has(im1, u, v) && !tri_has(ip2, u, v)) {
                        if (sg_next == 2) sg_next = tri_sign(snapped, n_verts, ip1, nx, ny);
                        // direction of u->v inside the neighbour after ITS normalisation; opposite to ours (which is u->v) cancels
                        drop = sg_next != 0 && tri_dir(ip1, u, v) * sg_next < 0;
                    }
                } else if (tri_has(im1, u, v)) {
                    if (!tri_has(im2, u, v)) {
                        if (sg_prev == 2) sg_prev = tri_sign(snapped, n_verts, im1, nx, ny);
                        drop = sg_prev != 0 && tri_dir(im1, u, v) * sg_prev < 0;
                    }
                }
                if (!drop) e[k] = vkb_edge{x[ka], y[ka], x[kb], y[kb]};
            }
        }
        bool     keep[3];
        uint32_t cnt = 0;
#pragma unroll
        for (int k = 0; k < 3; k++) {
            keep[k] = in_range && !(e[k].x0 == e[k].x1 && e[k].y0 == e[k].y1) && !edge_off_surface(e[k], sd);
            cnt += keep[k] ? 1u : 0u;
        }
        const uint32_t incl = warp_incl_scan(cnt);
        uint32_t       base = 0;
        if ((threadIdx.x & 31) == 31 && incl) base = atomicAdd(live, incl);
        uint32_t pos = __shfl_sync(0xffffffffu, base, 31) + incl - cnt;
        BoxAcc box;
#pragma unroll
        for (int k = 0; k < 3; k++)
            if (keep[k]) { edges[pos] = e[k]; edge_draw[pos] = d; pos++; box.add(e[k]); }
        bbox_accumulate(bbox, d, box);
    }
}
// The same kernel with the neighbours taken from the neighbouring LANES: lane L holds triangle t, so the index triples of t +- 1, t +- 2 and
// the orientation of t +- 1 are already in the registers of lanes L +- 1, L +- 2 - five shuffles instead of twelve index loads, six vertex
// gathers and two 64-bit cross products per triangle.  Only the two lanes at either end of the warp (and the last triangles of the list)
// load what lies outside it.  Identical output, edge for edge (the order in which warps reserve room aside).
__global__ void __launch_bounds__(256)
tri_edges_warp_k(const int2 *snapped, const uint32_t *inds, const vkb_counts *C, const uint32_t *sdraw_id, const uint32_t *sdraw_first_item,
                 uint32_t n_sdraws, const unsigned long long *item_offsets, vkb_edge *edges, uint32_t *edge_draw, uint32_t n_extra, uint32_t *live,
                 SurfaceDesc sd, int32_t *bbox) {
    if (C->overflow) return;
    const uint32_t n_tris = C->n[VKC_TRIS], n_verts = C->n[VKC_VERTS], lane = threadIdx.x & 31u;
    edges += C->n[VKC_FEDGES] + n_extra; edge_draw += C->n[VKC_FEDGES] + n_extra;
    // (room for the surviving edges is reserved once per warp; once per block - two barriers per round - was measured slower: C3 0.222 -> 0.240 ms)
    for (uint32_t wb = blockIdx.x * blockDim.x + (threadIdx.x & ~31u); wb < n_tris; wb += gridDim.x * blockDim.x) {
        uint32_t   t = wb + lane;
        const bool in_range = t < n_tris;
        if (!in_range) t = n_tris - 1;
        uint32_t i0[3];
        tri_idx(inds, n_tris, (long long)t, i0);
        int32_t   x[3], y[3], nx[3], ny[3];
        const int sg_own = tri_sign(snapped, n_verts, i0, x, y);   // (what the neighbours ask for: before the off-surface test below)
        int       sg = sg_own;
        if (sg != 0 && (max(max(y[0], y[1]), y[2]) < 0 || min(min(y[0], y[1]), y[2]) > (int32_t)sd.height * 256 ||
                        min(min(x[0], x[1]), x[2]) > (int32_t)sd.width * 256))
            sg = 0;
        uint32_t ip1[3], ip2[3], im1[3], im2[3];
#pragma unroll
        for (int k = 0; k < 3; k++) {
            ip1[k] = __shfl_down_sync(0xffffffffu, i0[k], 1); ip2[k] = __shfl_down_sync(0xffffffffu, i0[k], 2);
            im1[k] = __shfl_up_sync(0xffffffffu, i0[k], 1);   im2[k] = __shfl_up_sync(0xffffffffu, i0[k], 2);
        }
        int sg_next = __shfl_down_sync(0xffffffffu, sg_own, 1), sg_prev = __shfl_up_sync(0xffffffffu, sg_own, 1);
        // what the lanes next door do not hold: beyond either end of the warp, or beyond the last triangle (those lanes hold a copy of it)
        if (lane > 30u || t + 1 >= n_tris) { tri_idx(inds, n_tris, (long long)t + 1, ip1); sg_next = 2; }
        if (lane > 29u || t + 2 >= n_tris) tri_idx(inds, n_tris, (long long)t + 2, ip2);
        if (lane < 1u) { tri_idx(inds, n_tris, (long long)t - 1, im1); sg_prev = 2; }
        if (lane < 2u) tri_idx(inds, n_tris, (long long)t - 2, im2);
        vkb_edge e[3];
        uint32_t d = 0;
#pragma unroll
        for (int k = 0; k < 3; k++) e[k] = vkb_edge{0, 0, 0, 0};
        if (sg != 0) {
            uint32_t lo = 0, hi = n_sdraws;  // stroke draw owning index 3t: last q whose first item's index offset <= 3t
            while (hi - lo > 1) {
                uint32_t mid = (lo + hi) >> 1;
                if ((uint32_t)(item_offsets[sdraw_first_item[mid]] >> 32) <= 3 * t) lo = mid; else hi = mid;
            }
            d = sdraw_id[lo];
            // which of this triangle's three vertices each neighbour holds (bit k: vertex k): 36 comparisons once instead of six per edge and neighbour
            uint32_t in_p1 = 0, in_p2 = 0, in_m1 = 0, in_m2 = 0;
#pragma unroll
            for (int k = 0; k < 3; k++) {
                in_p1 |= (uint32_t)(i0[k] == ip1[0] || i0[k] == ip1[1] || i0[k] == ip1[2]) << k;
                in_p2 |= (uint32_t)(i0[k] == ip2[0] || i0[k] == ip2[1] || i0[k] == ip2[2]) << k;
                in_m1 |= (uint32_t)(i0[k] == im1[0] || i0[k] == im1[1] || i0[k] == im1[2]) << k;
                in_m2 |= (uint32_t)(i0[k] == im2[0] || i0[k] == im2[1] || i0[k] == im2[2]) << k;
            }
#pragma unroll
            for (int k = 0; k < 3; k++) {
                const int ka = sg > 0 ? k : (3 - k) % 3, kb = sg > 0 ? (k + 1) % 3 : (5 - k) % 3;
                const uint32_t u = i0[ka], v = i0[kb];
                bool           drop = false;
                // (an edge lies in a neighbour when both its ends do: bits ka and kb of the neighbour's membership mask)
                const uint32_t eb = (1u << ka) | (1u << kb);
                if ((in_p1 & eb) == eb) {
                    if ((in_m1 & eb) != eb && (in_p2 & eb) != eb) {
                        if (sg_next == 2) sg_next = tri_sign(snapped, n_verts, ip1, nx, ny);
                        drop = sg_next != 0 && tri_dir(ip1, u, v) * sg_next < 0;
                    }
                } else if ((in_m1 & eb) == eb) {
                    if ((in_m2 & eb) != eb) {
                        if (sg_prev == 2) sg_prev = tri_sign(snapped, n_verts, im1, nx, ny);
                        drop = sg_prev != 0 && tri_dir(im1, u, v) * sg_prev < 0;
                    }
                }
                if (!drop) e[k] = vkb_edge{x[ka], y[ka], x[kb], y[kb]};
            }
        }
        bool     keep[3];
        uint32_t cnt = 0;
#pragma unroll
        for (int k = 0; k < 3; k++) {
            keep[k] = in_range && !(e[k].x0 == e[k].x1 && e[k].y0 == e[k].y1) && !edge_off_surface(e[k], sd);
            cnt += keep[k] ? 1u : 0u;
        }
        const uint32_t incl = warp_incl_scan(cnt);
        uint32_t       base = 0;
        if (lane == 31 && incl) base = atomicAdd(live, incl);
        uint32_t pos = __shfl_sync(0xffffffffu, base, 31) + incl - cnt;
        BoxAcc box;
#pragma unroll
        for (int k = 0; k < 3; k++)
            if (keep[k]) { edges[pos] = e[k]; edge_draw[pos] = d; pos++; box.add(e[k]); }
        bbox_accumulate(bbox, d, box);
    }
}
// the stroke edges that survived are only counted by tri_edges_k: C->n[VKC_EDGES] (so far the upper bound fill + 3 x triangles +
// rectangles, which sized the buffers) becomes the number actually stored
__global__ void commit_live_edges_k(vkb_counts *C, const uint32_t *live, uint32_t n_extra) {
    if (C->overflow) return;
    C->n[VKC_EDGES] = C->n[VKC_FEDGES] + n_extra + *live;
}
void vkb_launch_tri_edges(const float2 *verts, uint32_t cap_verts, int2 *snapped, const uint32_t *inds, uint32_t cap_tris, const vkb_counts *C,
                          const vkb_draw *draws, const vkb_xform *xforms, const uint32_t *sdraw_id, const uint32_t *sdraw_first_item, uint32_t n_sdraws,
                          const unsigned long long *item_offsets, SurfaceDesc sd, vkb_edge *edges, uint32_t *edge_draw, uint32_t n_extra, uint32_t *live,
                          vkb_counts *Cw, int32_t *draw_bbox, bool commit_live, cudaStream_t s) {
    if (!cap_tris || !n_sdraws) return;
    if (verts) {   // (null: the stroke emitter has run the vertex stage itself)
        snap_verts_k<<<min(vkb_div_up(cap_verts, 256), VKB_EDGE_GRID), 256, 0, s>>>(verts, C, draws, xforms, sdraw_id, sdraw_first_item, n_sdraws, item_offsets, sd, snapped);
        VKB_LAUNCHED();
    }
    if (vkb_stroke_emit_mode() == 1) tri_edges_k<<<min(vkb_div_up(cap_tris, 256), VKB_EDGE_GRID), 256, 0, s>>>(snapped, inds, C, sdraw_id, sdraw_first_item, n_sdraws, item_offsets, edges, edge_draw, n_extra, live, sd, draw_bbox);
    else tri_edges_warp_k<<<min(vkb_div_up(cap_tris, 256), VKB_EDGE_GRID), 256, 0, s>>>(snapped, inds, C, sdraw_id, sdraw_first_item, n_sdraws, item_offsets, edges, edge_draw, n_extra, live, sd, draw_bbox);
    VKB_LAUNCHED();
    if (commit_live) {   // (else vkb_launch_draw_rects does it: a frame that goes on to binning saves the launch)
        commit_live_edges_k<<<1, 1, 0, s>>>(Cw, live, n_extra);
        VKB_LAUNCHED();
    }
}

// ---- per-draw bounding boxes of a RAW edge list (vkb_winding_raw; draws' own edges grow their boxes where they are emitted) ----

__global__ void draw_bbox_init_k(int32_t *bbox, uint32_t n_draws) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_draws) return;
    bbox[4 * i] = INT32_MAX; bbox[4 * i + 1] = INT32_MAX; bbox[4 * i + 2] = INT32_MIN; bbox[4 * i + 3] = INT32_MIN;
}
__global__ void __launch_bounds__(256) draw_bbox_k(const vkb_edge *edges, const uint32_t *edge_draw, const vkb_counts *C, int32_t *bbox) {
    if (C->overflow) return;
    const uint64_t n_edges = C->n[VKC_EDGES];
    if ((uint64_t)blockIdx.x * blockDim.x >= n_edges) return;  // (whole block: the barriers below stay uniform)
    uint64_t i  = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool     ok = i < n_edges;
    vkb_edge e  = ok ? edges[i] : vkb_edge{0, 0, 0, 0};
    uint32_t d  = ok ? edge_draw[i] : 0xffffffffu;
    ok          = ok && !edge_degenerate(e);
    int32_t mnx = ok ? min(e.x0, e.x1) : INT32_MAX, mny = ok ? min(e.y0, e.y1) : INT32_MAX;
    int32_t mxx = ok ? max(e.x0, e.x1) : INT32_MIN, mxy = ok ? max(e.y0, e.y1) : INT32_MIN;
    __shared__ uint32_t d0;
    __shared__ int32_t  red[4][8];
    if (threadIdx.x == 0) d0 = d;
    __syncthreads();
    const bool block_uniform = __syncthreads_and(d == d0 || i >= n_edges);
    const uint32_t dw = __shfl_sync(0xffffffffu, d, 0);
    const bool warp_uniform = block_uniform || __all_sync(0xffffffffu, d == dw || i >= n_edges);
    if (warp_uniform) {
        mnx = __reduce_min_sync(0xffffffffu, mnx); mny = __reduce_min_sync(0xffffffffu, mny);
        mxx = __reduce_max_sync(0xffffffffu, mxx); mxy = __reduce_max_sync(0xffffffffu, mxy);
    }
    if (block_uniform) {
        if ((threadIdx.x & 31) == 0) {
            red[0][threadIdx.x >> 5] = mnx; red[1][threadIdx.x >> 5] = mny; red[2][threadIdx.x >> 5] = mxx; red[3][threadIdx.x >> 5] = mxy;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            for (int w = 1; w < 8; w++) {
                mnx = min(mnx, red[0][w]); mny = min(mny, red[1][w]); mxx = max(mxx, red[2][w]); mxy = max(mxy, red[3][w]);
            }
            bbox_grow(bbox, d0, mnx, mny, mxx, mxy);
        }
    } else if (warp_uniform) {
        if ((threadIdx.x & 31) == 0) bbox_grow(bbox, dw, mnx, mny, mxx, mxy);
    } else if (ok) {
        bbox_grow(bbox, d, mnx, mny, mxx, mxy);  // (looks first here too: where two long strokes meet, warps mix their edges)
    }
}
void vkb_launch_draw_bbox_init(uint32_t n_draws, int32_t *draw_bbox, cudaStream_t s) {
    if (!n_draws) return;
    draw_bbox_init_k<<<vkb_div_up(n_draws, 256), 256, 0, s>>>(draw_bbox, n_draws);
    VKB_LAUNCHED();
}
void vkb_launch_draw_bbox(const vkb_edge *edges, const uint32_t *edge_draw, uint64_t cap_edges, const vkb_counts *C, uint32_t n_draws, int32_t *draw_bbox,
                          cudaStream_t s) {
    vkb_launch_draw_bbox_init(n_draws, draw_bbox, s);
    if (!cap_edges) return;
    draw_bbox_k<<<vkb_div_up(cap_edges, 256), 256, 0, s>>>(edges, edge_draw, C, draw_bbox);
    VKB_LAUNCHED();
}

__device__ __forceinline__ int32_t floor_div(int32_t a, int32_t b) {  // b > 0
    int32_t q = a / b;
    return (a % b < 0) ? q - 1 : q;
}
// tile rectangle of each draw (clipped to the surface) and its path-tile / row counts packed as lo | hi<<32
#define VKB_GPREP_FLOATS 16
__device__ void grad_prep_one(const vkb_gradient *g, float W, float H, float *o);   // (with the paint evaluation, below)
// (live != null: thread 0 first turns the number of stroke edges tri_edges_k stored into the edge count the binning kernels behind it read)
__global__ void draw_rects_k(const int32_t *bbox, const vkb_draw *draws, const vkb_xform *xforms, uint32_t n_draws, SurfaceDesc sd, int32_t *rect,
                             unsigned long long *counts, vkb_counts *Cw, const uint32_t *live, uint32_t n_extra, GradPrep gp) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0 && live && !Cw->overflow) Cw->n[VKC_EDGES] = Cw->n[VKC_FEDGES] + n_extra + *live;
    // the position-independent terms of every gradient of the batch, for the fine pass (what used to be a launch of its own)
    for (uint32_t g = i; g < gp.n; g += gridDim.x * blockDim.x) grad_prep_one(gp.grads + g, gp.W, gp.H, gp.out + (size_t)g * VKB_GPREP_FLOATS);
    if (i >= n_draws) return;
    int32_t mnx = bbox[4 * i], mny = bbox[4 * i + 1], mxx = bbox[4 * i + 2], mxy = bbox[4 * i + 3];
    int32_t tx0 = 0, ty0 = 0, tw = 0, th = 0;
    // rows the draw may touch: the whole surface, or the band of its canvas on a batch surface
    int32_t row_lo = 0, row_hi = (int32_t)sd.tiles_y - 1;
    if (sd.band_tiles && draws) {
        const uint32_t band = xforms[draws[i].xform_stroke & 0xFFFF].band;
        row_lo = (int32_t)(band * sd.band_tiles);
        row_hi = min(row_hi, row_lo + (int32_t)sd.band_tiles - 1);
    }
    if (draws && draws[i].kind == VKB_DRAW_CLIP) {  // everything outside the clip path is affected too
        tw = (int32_t)sd.tiles_x; ty0 = row_lo; th = max(row_hi - row_lo + 1, 0);
        if (th == 0) tw = 0;
    } else if (mnx <= mxx && mxx >= 0 && mxy >= 0 && mnx < (int32_t)sd.width * 256 && mny < (int32_t)sd.height * 256) {
        tx0         = max(floor_div(mnx, VKB_TILE_FX), 0);
        ty0         = max(floor_div(mny, VKB_TILE_FX), row_lo);
        int32_t tx1 = min(floor_div(mxx, VKB_TILE_FX), (int32_t)sd.tiles_x - 1);
        int32_t ty1 = min(floor_div(mxy, VKB_TILE_FX), row_hi);
        tw = tx1 - tx0 + 1; th = ty1 - ty0 + 1;
        if (tw <= 0 || th <= 0) tw = th = 0;
    }
    rect[4 * i] = tx0; rect[4 * i + 1] = ty0; rect[4 * i + 2] = tw; rect[4 * i + 3] = th;
    counts[i] = (unsigned long long)((uint32_t)(tw * th)) | ((unsigned long long)(uint32_t)th << 32);
}
void vkb_launch_draw_rects(const int32_t *draw_bbox, const vkb_draw *draws, const vkb_xform *xforms, uint32_t n_draws, SurfaceDesc sd, int32_t *draw_rect,
                           unsigned long long *tile_row_counts, vkb_counts *Cw, const uint32_t *live, uint32_t n_extra, const GradPrep &gp, cudaStream_t s) {
    draw_rects_k<<<vkb_div_up(n_draws, 256), 256, 0, s>>>(draw_bbox, draws, xforms, n_draws, sd, draw_rect, tile_row_counts, Cw, live, n_extra, gp);
    VKB_LAUNCHED();
}
// (thread 0 also commits the totals of the scan - path-tiles and path-tile rows - as checked counts: one launch less per frame)
__global__ void split_bases_k(const unsigned long long *packed, uint32_t n, uint32_t *lo, uint32_t *hi, vkb_counts *C, const unsigned long long *total) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0 && C && !C->overflow) {
        vkc_commit(C, VKC_PT, (uint32_t)(*total & 0xffffffffull));
        vkc_commit(C, VKC_ROWS, (uint32_t)(*total >> 32));
    }
    if (i >= n) return;
    lo[i] = (uint32_t)(packed[i] & 0xffffffffull);
    hi[i] = (uint32_t)(packed[i] >> 32);
}
void vkb_launch_split_bases(const unsigned long long *packed, uint32_t n, uint32_t *lo, uint32_t *hi, vkb_counts *C, const unsigned long long *total, cudaStream_t s) {
    if (!n) return;
    split_bases_k<<<vkb_div_up(n, 256), 256, 0, s>>>(packed, n, lo, hi, C, total);
    VKB_LAUNCHED();
}

// ---- binning: walk the tile rows an edge touches ----
struct EdgeWalk {
    int32_t r0, r1;  // tile rows (clamped to the draw rectangle), r0 > r1 when nothing
};
// first absolute tile column whose virtual sample column X0+1/2 is at or right of the edge at height Y0+1/2
__device__ __forceinline__ long long backdrop_first_col(const vkb_edge &e, int32_t Y0) {
    long long dx = (long long)e.x1 - e.x0, dy = (long long)e.y1 - e.y0;
    long long num = dx * (2ll * Y0 + 1 - 2ll * e.y0);  // doubled coordinates: Q = 2*x0 + num/dy
    // ceil(num / dy) exactly
    long long q = num / dy, r = num % dy;
    if (r != 0 && ((r < 0) == (dy < 0))) q++;
    long long Qc = 2ll * e.x0 + q;            // smallest integer L2 with L2 >= Q
    // L2 = 2*4096*tc + 1 >= Qc  <=>  tc >= ceil((Qc - 1) / 8192)
    long long n2 = Qc - 1, t = n2 / 8192;
    if (n2 % 8192 > 0) t++;                   // (for negative n2 truncation already rounds toward +inf)
    return t;
}
// number of tile rows + tile columns of the edge's bounding box inside the draw rectangle: edges above VKB_LONG_EDGE go to a
// list that one warp per edge walks cooperatively (a 4096-pixel side of a rectangle touches 256 tiles: one thread would visit
// them one after the other while the rest of the grid has long finished)
#define VKB_LONG_EDGE 40
__device__ __forceinline__ int32_t edge_tile_span(const vkb_edge &e, const int32_t *rect) {
    const int32_t tx0 = rect[0], ty0 = rect[1], tw = rect[2], th = rect[3];
    if (tw <= 0) return 0;
    int32_t r0 = max(floor_div(min(e.y0, e.y1), VKB_TILE_FX), ty0), r1 = min(floor_div(max(e.y0, e.y1), VKB_TILE_FX), ty0 + th - 1);
    int32_t c0 = max(floor_div(min(e.x0, e.x1), VKB_TILE_FX), tx0), c1 = min(floor_div(max(e.x0, e.x1), VKB_TILE_FX), tx0 + tw - 1);
    if (r1 < r0) return 0;
    return (r1 - r0 + 1) + max(c1 - c0 + 1, 0);
}
// rows r0 + row_off, + row_step ...; in each row columns c0 + col_off, + col_step ... (1 thread: 0,1,0,1; a warp on a steep edge:
// lane,32,0,1; on a shallow one: 0,1,lane,32)
template <class F> __device__ __forceinline__ void for_each_tile_of_edge(const vkb_edge &e, const int32_t *rect, F &&f_tile, bool want_backdrop,
                                                                        int32_t *pt_backdrop, uint32_t ptbase, int32_t row_off = 0, int32_t row_step = 1,
                                                                        int32_t col_off = 0, int32_t col_step = 1) {
    const int32_t tx0 = rect[0], ty0 = rect[1], tw = rect[2], th = rect[3];
    if (tw <= 0) return;
    const int32_t ymin = min(e.y0, e.y1), ymax = max(e.y0, e.y1);
    int32_t r0 = max(floor_div(ymin, VKB_TILE_FX), ty0), r1 = min(floor_div(ymax, VKB_TILE_FX), ty0 + th - 1);
    if (r0 + row_off > r1) return;  // outside the draw's rows (on a stripe surface: most edges of the scene) before any double arithmetic
    const double dxdy = (e.y1 != e.y0) ? ((double)e.x1 - (double)e.x0) / ((double)e.y1 - (double)e.y0) : 0.0;
    const int    sgn  = e.y1 > e.y0 ? 1 : -1;
    for (int32_t r = r0 + row_off; r <= r1; r += row_step) {
        const int32_t Y0 = r * VKB_TILE_FX;
        if (want_backdrop && col_off == 0 && e.y0 != e.y1 && ((e.y0 <= Y0) != (e.y1 <= Y0))) {
            long long tc = backdrop_first_col(e, Y0);
            if (tc < tx0) tc = tx0;
            if (tc < tx0 + tw) atomicAdd(&pt_backdrop[ptbase + (uint32_t)(r - ty0) * tw + (uint32_t)(tc - tx0)], sgn);
        }
        // conservative column range of the edge inside this row band
        double xa, xb;
        if (e.y0 == e.y1) { xa = e.x0; xb = e.x1; }
        else {
            double ya = max(ymin, Y0), yb = min(ymax, Y0 + VKB_TILE_FX);
            xa = (double)e.x0 + (ya - (double)e.y0) * dxdy;
            xb = (double)e.x0 + (yb - (double)e.y0) * dxdy;
        }
        double  xl = floor(fmin(xa, xb)) - 1.0, xh = ceil(fmax(xa, xb)) + 1.0;
        int32_t c0 = (int32_t)fmax(floor(xl / VKB_TILE_FX), (double)tx0), c1 = (int32_t)fmin(floor(xh / VKB_TILE_FX), (double)(tx0 + tw - 1));
        for (int32_t c = c0 + col_off; c <= c1; c += col_step) f_tile(ptbase + (uint32_t)(r - ty0) * tw + (uint32_t)(c - tx0));
    }
}
__device__ __forceinline__ bool edge_is_shallow(const vkb_edge &e) { return llabs((long long)e.x1 - e.x0) >= llabs((long long)e.y1 - e.y0); }
__global__ void __launch_bounds__(256) bin_count_k(const vkb_edge *edges, const uint32_t *edge_draw, const vkb_counts *C, const int32_t *draw_rect,
                                                  const uint32_t *draw_ptbase, uint32_t *pt_count, int32_t *pt_backdrop, uint32_t *long_list,
                                                  uint32_t *long_n) {
    if (C->overflow) return;
    // grid-stride over the LIVE edges: the grid is sized from the capacity of the buffer (3 x triangles for strokes, of which a stripe keeps
    // an eighth), capped at a few waves
    const uint32_t n = C->n[VKC_EDGES], stride = gridDim.x * blockDim.x;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        vkb_edge e = edges[i];
        if (edge_degenerate(e)) continue;
        uint32_t d = edge_draw[i];
        if (edge_tile_span(e, draw_rect + 4 * d) > VKB_LONG_EDGE) { long_list[atomicAdd(long_n, 1u)] = i; continue; }
        for_each_tile_of_edge(e, draw_rect + 4 * d, [&](uint32_t pt) { atomicAdd(&pt_count[pt], 1u); }, true, pt_backdrop, draw_ptbase[d]);
    }
}
// one warp per long edge (grid-stride over the list)
__global__ void __launch_bounds__(256) bin_count_long_k(const vkb_edge *edges, const uint32_t *edge_draw, const vkb_counts *C, const int32_t *draw_rect,
                                                       const uint32_t *draw_ptbase, uint32_t *pt_count, int32_t *pt_backdrop, const uint32_t *long_list,
                                                       const uint32_t *long_n) {
    if (C->overflow) return;
    const uint32_t n = *long_n, lane = threadIdx.x & 31, warps = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t k = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; k < n; k += warps) {
        const uint32_t i = long_list[k], d = edge_draw[i];
        const vkb_edge e = edges[i];
        const bool     sh = edge_is_shallow(e);
        for_each_tile_of_edge(e, draw_rect + 4 * d, [&](uint32_t pt) { atomicAdd(&pt_count[pt], 1u); }, true, pt_backdrop, draw_ptbase[d],
                              sh ? 0 : (int32_t)lane, sh ? 1 : 32, sh ? (int32_t)lane : 0, sh ? 32 : 1);
    }
}
void vkb_launch_bin_count(const vkb_edge *edges, const uint32_t *edge_draw, uint64_t cap_edges, const vkb_counts *C, const int32_t *draw_rect,
                          const uint32_t *draw_ptbase, uint32_t *pt_count, int32_t *pt_backdrop, uint32_t *long_list, uint32_t *long_n, cudaStream_t s) {
    if (!cap_edges) return;
    bin_count_k<<<min(vkb_div_up(cap_edges, 256), VKB_EDGE_GRID), 256, 0, s>>>(edges, edge_draw, C, draw_rect, draw_ptbase, pt_count, pt_backdrop, long_list, long_n);
    VKB_LAUNCHED();
    bin_count_long_k<<<148 * 2, 256, 0, s>>>(edges, edge_draw, C, draw_rect, draw_ptbase, pt_count, pt_backdrop, long_list, long_n);
    VKB_LAUNCHED();
}
// the first main_blocks blocks stride the live edges, one thread per edge; the blocks behind them walk the long list bin_count_k made, one warp
// per edge (one launch for both: the list is complete before this kernel starts)
__global__ void __launch_bounds__(256) bin_scatter_k(const vkb_edge *edges, const uint32_t *edge_draw, const vkb_counts *C, const int32_t *draw_rect,
                                                    const uint32_t *draw_ptbase, const uint32_t *pt_slot, const uint32_t *eoff, uint32_t *cursor,
                                                    vkb_edge *tile_edges, const uint32_t *long_list, const uint32_t *long_n, uint32_t main_blocks) {
    if (C->overflow) return;
    if (blockIdx.x >= main_blocks) {
        const uint32_t n = *long_n, lane = threadIdx.x & 31, warps = ((gridDim.x - main_blocks) * blockDim.x) >> 5;
        for (uint32_t k = ((blockIdx.x - main_blocks) * blockDim.x + threadIdx.x) >> 5; k < n; k += warps) {
            const uint32_t i = long_list[k], d = edge_draw[i];
            const vkb_edge e = edges[i];
            const bool     sh = edge_is_shallow(e);
            for_each_tile_of_edge(
                e, draw_rect + 4 * d,
                [&](uint32_t pt) {
                    uint32_t p   = pt_slot[pt];
                    uint32_t pos = eoff[p] + atomicAdd(&cursor[p], 1u);
                    tile_edges[pos] = e;
                },
                false, nullptr, draw_ptbase[d], sh ? 0 : (int32_t)lane, sh ? 1 : 32, sh ? (int32_t)lane : 0, sh ? 32 : 1);
        }
        return;
    }
    const uint32_t n = C->n[VKC_EDGES], stride = main_blocks * blockDim.x;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        vkb_edge e = edges[i];
        if (edge_degenerate(e)) continue;
        uint32_t d = edge_draw[i];
        if (edge_tile_span(e, draw_rect + 4 * d) > VKB_LONG_EDGE) continue;  // on the long list
        for_each_tile_of_edge(
            e, draw_rect + 4 * d,
            [&](uint32_t pt) {
                uint32_t p   = pt_slot[pt];
                uint32_t pos = eoff[p] + atomicAdd(&cursor[p], 1u);
                tile_edges[pos] = e;
            },
            false, nullptr, draw_ptbase[d]);
    }
}
void vkb_launch_bin_scatter(const vkb_edge *edges, const uint32_t *edge_draw, uint64_t cap_edges, const vkb_counts *C, const int32_t *draw_rect,
                            const uint32_t *draw_ptbase, const uint32_t *pt_slot, const uint32_t *eoff, uint32_t *cursor, vkb_edge *tile_edges,
                            const uint32_t *long_list, const uint32_t *long_n, cudaStream_t s) {
    if (!cap_edges) return;
    const uint32_t main_blocks = min(vkb_div_up(cap_edges, 256), VKB_EDGE_GRID);
    bin_scatter_k<<<main_blocks + 148 * 2, 256, 0, s>>>(edges, edge_draw, C, draw_rect, draw_ptbase, pt_slot, eoff, cursor, tile_edges, long_list, long_n, main_blocks);
    VKB_LAUNCHED();
}

// ---- which draw owns a path-tile / a path-tile row: written once per draw (one warp each) so that the per-path-tile and
//      per-row kernels below read one word instead of binary-searching the draw table ----
__global__ void __launch_bounds__(256) owners_k(const int32_t *draw_rect, const uint32_t *draw_ptbase, const uint32_t *draw_rowbase, uint32_t n_draws,
                                               const vkb_counts *C, uint32_t *pt_owner, uint32_t *row_owner, uint32_t *pt_count, int32_t *pt_backdrop) {
    if (C->overflow) return;
    const uint32_t d = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (d >= n_draws) return;
    const uint32_t tw = (uint32_t)draw_rect[4 * d + 2], th = (uint32_t)draw_rect[4 * d + 3];
    const uint32_t base = draw_ptbase[d];
    uint32_t *po = pt_owner + base, *ro = row_owner + draw_rowbase[d];
    // (the path-tiles of the draws tile [0, C->n[VKC_PT]) exactly: what bin_count_k adds to starts from zero without a memset of the capacity)
    if (pt_count) for (uint32_t k = lane; k < tw * th; k += 32) { po[k] = d; pt_count[base + k] = 0u; pt_backdrop[base + k] = 0; }
    else for (uint32_t k = lane; k < tw * th; k += 32) po[k] = d;
    for (uint32_t k = lane; k < th; k += 32) ro[k] = d;
}
void vkb_launch_owners(const int32_t *draw_rect, const uint32_t *draw_ptbase, const uint32_t *draw_rowbase, uint32_t n_draws, const vkb_counts *C,
                       uint32_t *pt_owner, uint32_t *row_owner, uint32_t *pt_count, int32_t *pt_backdrop, cudaStream_t s) {
    if (!n_draws) return;
    owners_k<<<vkb_div_up((uint64_t)n_draws * 32, 256), 256, 0, s>>>(draw_rect, draw_ptbase, draw_rowbase, n_draws, C, pt_owner, row_owner, pt_count, pt_backdrop);
    VKB_LAUNCHED();
}

// ---- backdrop: inclusive prefix sum along every path-tile row; eight lanes per row (most draws are a few tiles wide: a whole
//      warp per row left 29 lanes idle on C2), grid-stride over rows ----
__global__ void __launch_bounds__(256) backdrop_prefix_k(const int32_t *draw_rect, const uint32_t *draw_ptbase, const uint32_t *draw_rowbase,
                                                        const uint32_t *row_owner, const vkb_counts *C, int32_t *pt_backdrop) {
    if (C->overflow) return;
    const uint32_t n_rows = C->n[VKC_ROWS];
    const uint32_t sub = threadIdx.x & 7, groups = (gridDim.x * blockDim.x) >> 3;
    const uint32_t gmask = 0xFFu << (threadIdx.x & 24);  // the eight lanes of this group
    for (uint32_t row = (blockIdx.x * blockDim.x + threadIdx.x) >> 3; row < n_rows; row += groups) {
        const uint32_t d = row_owner[row];
        uint32_t tw   = (uint32_t)draw_rect[4 * d + 2];
        int32_t *p    = pt_backdrop + draw_ptbase[d] + (row - draw_rowbase[d]) * tw;
        int32_t  carry = 0;
        for (uint32_t c = 0; c < tw; c += 8) {
            int32_t v = c + sub < tw ? p[c + sub] : 0;
#pragma unroll
            for (int o = 1; o < 8; o <<= 1) {
                const int32_t u = __shfl_up_sync(gmask, v, o, 8);
                if (sub >= (uint32_t)o) v += u;
            }
            v += carry;
            if (c + sub < tw) p[c + sub] = v;
            carry = __shfl_sync(gmask, v, 7, 8);
        }
    }
}
void vkb_launch_backdrop_prefix(const int32_t *draw_rect, const uint32_t *draw_ptbase, const uint32_t *draw_rowbase, const uint32_t *row_owner,
                                uint32_t cap_rows, const vkb_counts *C, int32_t *pt_backdrop, cudaStream_t s) {
    if (!cap_rows) return;
    const uint32_t blocks = cap_rows / 32 + 1 < 148 * 16 ? cap_rows / 32 + 1 : 148 * 16;  // eight lanes per row, grid-stride
    backdrop_prefix_k<<<blocks, 256, 0, s>>>(draw_rect, draw_ptbase, draw_rowbase, row_owner, C, pt_backdrop);
    VKB_LAUNCHED();
}

// ---- compaction of non-empty path-tiles and the per-tile ordered lists ----
__global__ void pt_flags_k(const uint32_t *pt_count, const int32_t *pt_backdrop, const vkb_counts *C, const vkb_draw *draws, const uint32_t *pt_owner,
                           bool keep_clip, uint32_t *flags) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (C->overflow || i >= C->n[VKC_PT]) return;
    bool keep = pt_count[i] != 0 || pt_backdrop[i] != 0;
    if (!keep && keep_clip) keep = draws[pt_owner[i]].kind == VKB_DRAW_CLIP;  // an empty path-tile of a clip draw still clips its whole tile out
    flags[i] = keep ? 1u : 0u;
}
void vkb_launch_pt_flags(const uint32_t *pt_count, const int32_t *pt_backdrop, uint32_t cap_pt, const vkb_counts *C, const vkb_draw *draws,
                         const uint32_t *pt_owner, bool keep_clip, uint32_t *flags, cudaStream_t s) {
    if (!cap_pt) return;
    pt_flags_k<<<vkb_div_up(cap_pt, 256), 256, 0, s>>>(pt_count, pt_backdrop, C, draws, pt_owner, keep_clip && draws, flags);
    VKB_LAUNCHED();
}
__global__ void pt_compact_k(const uint32_t *flags, const uint32_t *flag_scan, const vkb_counts *C, const int32_t *draw_rect, const uint32_t *draw_ptbase,
                             const uint32_t *pt_owner, SurfaceDesc sd, uint32_t *keys, uint32_t *vals, uint32_t *pt_draw) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (C->overflow || i >= C->n[VKC_PT] || !flags[i]) return;
    const uint32_t d = pt_owner[i];
    uint32_t tw = (uint32_t)draw_rect[4 * d + 2], local = i - draw_ptbase[d];
    uint32_t tx = (uint32_t)draw_rect[4 * d] + local % tw, ty = (uint32_t)draw_rect[4 * d + 1] + local / tw;
    uint32_t p = flag_scan[i];
    keys[p]    = ty * sd.tiles_x + tx;
    vals[p]    = i;
    pt_draw[p] = d;  // in compaction order == path-tile order; re-read through vals after the sort
}
void vkb_launch_pt_compact(const uint32_t *flags, const uint32_t *flag_scan, uint32_t cap_pt, const vkb_counts *C, const int32_t *draw_rect,
                           const uint32_t *draw_ptbase, const uint32_t *pt_owner, SurfaceDesc sd, uint32_t *keys, uint32_t *vals, uint32_t *pt_draw, cudaStream_t s) {
    if (!cap_pt) return;
    pt_compact_k<<<vkb_div_up(cap_pt, 256), 256, 0, s>>>(flags, flag_scan, C, draw_rect, draw_ptbase, pt_owner, sd, keys, vals, pt_draw);
    VKB_LAUNCHED();
}
// Also zeroes what the kernels behind it start from - the scatter cursor of every non-empty path-tile, the list bounds of every tile
// (headers_k only writes the tiles that have a list) and, when the frame starts from a cleared surface, the per-tile multisample flags
// (tile_ms_words: the flags as 32-bit words, or null) - in place of four memsets of whole capacities.
__global__ void sorted_counts_k(const uint32_t *vals, const vkb_counts *C, const uint32_t *pt_count, uint32_t *sorted_cnt, uint32_t *pt_slot, uint32_t *cursor,
                                uint32_t *tile_first, uint32_t *tile_end, uint32_t n_tiles, uint32_t *tile_ms_words) {
    uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    for (uint32_t t = p; t < n_tiles; t += gridDim.x * blockDim.x) {   // (whatever became of the counts: a freshly allocated flag plane is zeroed by the attempt that allocated it)
        tile_first[t] = 0u; tile_end[t] = 0u;
        if (tile_ms_words && t < (n_tiles + 3) / 4) tile_ms_words[t] = 0u;
    }
    if (C->overflow || p >= C->n[VKC_NE]) return;
    uint32_t pt   = vals[p];
    sorted_cnt[p] = pt_count[pt];
    pt_slot[pt]   = p;
    cursor[p]     = 0u;
}
void vkb_launch_sorted_counts(const uint32_t *vals, uint32_t cap_ne, const vkb_counts *C, const uint32_t *pt_count, uint32_t *sorted_cnt, uint32_t *pt_slot,
                              uint32_t *cursor, uint32_t *tile_first, uint32_t *tile_end, uint32_t n_tiles, uint32_t *tile_ms_words, cudaStream_t s) {
    if (!cap_ne) return;
    // (a few path-tiles on a large surface: enough blocks for the zeroing loop as well)
    sorted_counts_k<<<max(vkb_div_up(cap_ne, 256), min(vkb_div_up(n_tiles, 256), 148u * 8u)), 256, 0, s>>>(vals, C, pt_count, sorted_cnt, pt_slot, cursor, tile_first, tile_end, n_tiles, tile_ms_words);
    VKB_LAUNCHED();
}
// pt_draw_by_flagpos: draw of the path-tile at compaction position flag_scan[pt] (pt_compact_k output)
__global__ void headers_k(const uint32_t *keys, const uint32_t *vals, const vkb_counts *C, const uint32_t *pt_draw_by_flagpos, const uint32_t *flag_scan,
                          const int32_t *pt_backdrop, const uint32_t *pt_count, const uint32_t *eoff, const vkb_paint *paints, int4 *hdr,
                          uint32_t *tile_first, uint32_t *tile_end) {
    uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (C->overflow) return;
    const uint32_t n_ne = C->n[VKC_NE];
    if (p >= n_ne) return;
    uint32_t pt = vals[p], key = keys[p];
    uint32_t dr = pt_draw_by_flagpos[flag_scan[pt]];
    hdr[2 * p]     = make_int4((int)dr, pt_backdrop[pt], (int)eoff[p], (int)pt_count[pt]);
    hdr[2 * p + 1] = *(const int4 *)(paints + dr);  // the fine pass reads one 32-byte record per path-tile
    if (p == 0 || keys[p - 1] != key) tile_first[key] = p;
    if (p == n_ne - 1 || keys[p + 1] != key) tile_end[key] = p + 1;
}
void vkb_launch_headers(const uint32_t *keys, const uint32_t *vals, uint32_t cap_ne, const vkb_counts *C, const uint32_t *pt_draw_by_flagpos,
                        const uint32_t *flag_scan, const int32_t *pt_backdrop, const uint32_t *pt_count, const uint32_t *eoff, const vkb_paint *paints, int4 *hdr,
                        uint32_t *tile_first, uint32_t *tile_end, cudaStream_t s) {
    if (!cap_ne) return;
    headers_k<<<vkb_div_up(cap_ne, 256), 256, 0, s>>>(keys, vals, C, pt_draw_by_flagpos, flag_scan, pt_backdrop, pt_count, eoff, paints, hdr, tile_first, tile_end);
    VKB_LAUNCHED();
}

// ----------------------------------------------------------------------------------------------------
// fine pass
// ----------------------------------------------------------------------------------------------------
template <int S> struct SamplePos;
template <> struct SamplePos<1> { static __device__ __forceinline__ int x(int) { return 8; } static __device__ __forceinline__ int y(int) { return 8; } };
template <> struct SamplePos<2> {
    static __device__ __forceinline__ int x(int s) { return s == 0 ? 12 : 4; }
    static __device__ __forceinline__ int y(int s) { return s == 0 ? 12 : 4; }
};
template <> struct SamplePos<4> {
    static __device__ __forceinline__ int x(int s) { const int t[4] = {6, 14, 2, 10}; return t[s]; }
    static __device__ __forceinline__ int y(int s) { const int t[4] = {2, 6, 10, 14}; return t[s]; }
};
template <> struct SamplePos<8> {
    static __device__ __forceinline__ int x(int s) { const int t[8] = {9, 7, 13, 5, 3, 1, 11, 15}; return t[s]; }
    static __device__ __forceinline__ int y(int s) { const int t[8] = {5, 11, 9, 3, 13, 7, 15, 1}; return t[s]; }
};
template <> struct SamplePos<16> {
    static __device__ __forceinline__ int x(int s) { const int t[16] = {9, 7, 5, 12, 3, 10, 13, 11, 6, 8, 4, 2, 0, 15, 14, 1}; return t[s]; }
    static __device__ __forceinline__ int y(int s) { const int t[16] = {9, 5, 10, 7, 6, 13, 11, 3, 14, 1, 2, 12, 8, 4, 15, 0}; return t[s]; }
};

__device__ __forceinline__ float clamp01(float x) { return x < 0.0f ? 0.0f : (x > 1.0f ? 1.0f : x); }
__device__ __forceinline__ float smoothstepf(float e0, float e1, float x) {
    float t = clamp01((x - e0) / (e1 - e0));
    return t * t * (3.0f - 2.0f * t);
}
__device__ __forceinline__ void mix4(float *c, const float *b, float t) {
#pragma unroll
    for (int k = 0; k < 4; k++) c[k] = c[k] * (1.0f - t) + b[k] * t;
}
// shaders/vkvg_main.frag:68-157 (SOLID / LINEAR / RADIAL) at the pixel centre; identical arithmetic to
// oracle/vkvg_oracle.c: eval_paint.  Everything the shader derives from the gradient record alone (normalised control
// points, axis direction, line slope ...) is evaluated once per gradient by grad_prep_one (run by draw_rects_k) with the same float operations in
// the same order, so a pixel only pays for what depends on its position.
__device__ void grad_prep_one(const vkb_gradient *g, float W, float H, float *o) {
    {   // linear (frag :84-107)
        float p0x = g->cp[0][0] / W, p0y = g->cp[0][1] / H;
        float p1x = g->cp[0][2] / W, p1y = g->cp[0][3] / H;
        float dx = p1x - p0x, dy = p1y - p0y;
        float l  = sqrtf(dx * dx + dy * dy);
        float ux = dx / l, uy = dy / l;
        float m  = -ux / uy;
        float bb = p0y - m * p0x;
        o[0] = p0x; o[1] = l; o[2] = ux; o[3] = uy; o[4] = m; o[5] = bb; o[6] = sqrtf(1.0f + m * m);
    }
    {   // radial (frag :109-144)
        float c0x = g->cp[0][0] / W, c0y = g->cp[0][1] / H;
        float c1x = g->cp[1][0] / W, c1y = g->cp[1][1] / H;
        float r0 = g->cp[0][2] / W, r1 = g->cp[1][2] / W;
        float dfx = c0x - c1x, dfy = c0y - c1y;
        o[8] = c0x; o[9] = c0y; o[10] = r0; o[11] = dfx; o[12] = dfy; o[13] = (dfx * dfx + dfy * dfy) - r1 * r1;
    }
}
// px, py: the pixel centre divided by the surface size (gl_FragCoord.xy / size in the shader): fx / W, fy / H
__device__ __noinline__ void eval_gradient(uint32_t pattern, const vkb_gradient *g, const float *gp, float px, float py, float c[4]) {
    if (pattern == VKB_PAT_LINEAR) {
        const float p0x = gp[0], l = gp[1], ux = gp[2], uy = gp[3];
        float dist;
        if (uy == 0.0f) {
            if (ux < 0.0f) dist = -(px - p0x) / l;
            else dist = (px - p0x) / l;
        } else {
            const float m = gp[4], bb = gp[5], sq = gp[6];
            dist = ((py - m * px - bb) / sq) / l;
            if (uy < 0.0f) dist = -dist;
        }
        for (int k = 0; k < 4; k++) c[k] = g->colors[0][k];
        mix4(c, g->colors[1], smoothstepf(g->stops[0], g->stops[1], dist));
        for (uint32_t i = 1; i + 1 < g->count; ++i) mix4(c, g->colors[i + 1], smoothstepf(g->stops[i], g->stops[i + 1], dist));
    } else {
        const float c0x = gp[8], c0y = gp[9], r0 = gp[10], dfx = gp[11], dfy = gp[12], cc = gp[13];
        float gradLength = 1.0f;
        float rx = px - c0x, ry = py - c0y;
        float rl = sqrtf(rx * rx + ry * ry);
        float rdx = rx / rl, rdy = ry / rl;
        float a    = rdx * rdx + rdy * rdy;
        float b    = 2.0f * (rdx * dfx + rdy * dfy);
        float disc = b * b - 4.0f * a * cc;
        if (disc >= 0.0f) {
            float t   = (-b + sqrtf(fabsf(disc))) / (2.0f * a);
            float prx = c0x + rdx * t, pry = c0y + rdy * t;
            float ex = prx - c0x, ey = pry - c0y;
            gradLength = sqrtf(ex * ex + ey * ey) - r0;
        }
        float grad = (rl - r0) / gradLength;
        for (int k = 0; k < 4; k++) c[k] = g->colors[0][k];
        mix4(c, g->colors[1], smoothstepf(g->stops[0], g->stops[1], grad));
        for (uint32_t i = 2; i < g->count; i++) mix4(c, g->colors[i], smoothstepf(g->stops[i - 1], g->stops[i], grad));
    }
}
// texture(source, uv) of shaders/vkvg_main.frag:72-82 with the sampler src/vkvg_context_internal.c:730-755 builds: nearest or
// linear filtering of the unnormalised coordinate, address mode per vkvg_extend_t, transparent-black border — the Vulkan
// texel-addressing rules restated (same arithmetic as oracle/vkvg_oracle.c: sample_surface).
__device__ __forceinline__ int tex_wrap(int i, int n, uint32_t mode, bool &border) {
    if (mode == VKB_TEX_REPEAT) { i %= n; return i < 0 ? i + n : i; }
    if (mode == VKB_TEX_MIRROR) {
        int m = i % (2 * n);
        if (m < 0) m += 2 * n;
        m -= n;
        return (n - 1) - (m >= 0 ? m : -(1 + m));
    }
    if (mode == VKB_TEX_EDGE) return i < 0 ? 0 : (i >= n ? n - 1 : i);
    border = border || i < 0 || i >= n;
    return i < 0 ? 0 : (i >= n ? n - 1 : i);
}
__device__ __forceinline__ void tex_fetch(const vkb_surfpat &sp, int i, int j, uint32_t mode, const float *lut, float t[4]) {
    bool border = false;
    i = tex_wrap(i, (int)sp.width, mode, border);
    j = tex_wrap(j, (int)sp.height, mode, border);
    const uint32_t p = border ? 0u : ((const uint32_t *)sp.image)[(size_t)j * sp.width + i];
#pragma unroll
    for (int k = 0; k < 4; k++) t[k] = lut[(p >> (8 * k)) & 0xFF];
}
__device__ __noinline__ void eval_surface(const vkb_surfpat *spp, float fx, float fy, float c[4], const float *lut) {
    const vkb_surfpat sp = *spp;
    const float px = fx - sp.sx, py = fy - sp.sy;
    float u = (sp.minv[0] * px + sp.minv[2] * py + sp.minv[4]) / (float)sp.width;
    float v = (sp.minv[1] * px + sp.minv[3] * py + sp.minv[5]) / (float)sp.height;
    const uint32_t mode = sp.filter_extend >> 8;
    float U = u * (float)sp.width, V = v * (float)sp.height;
    if (!(fabsf(U) < 1.0e9f) || !(fabsf(V) < 1.0e9f)) { c[0] = c[1] = c[2] = c[3] = 0.0f; return; }  // NaN / far outside: border
    if ((sp.filter_extend & 0xFF) == VKB_TEX_NEAREST) {
        tex_fetch(sp, (int)floorf(U), (int)floorf(V), mode, lut, c);
        return;
    }
    U -= 0.5f; V -= 0.5f;
    const float fi = floorf(U), fj = floorf(V);
    const float al = U - fi, be = V - fj;
    float t00[4], t10[4], t01[4], t11[4];
    tex_fetch(sp, (int)fi, (int)fj, mode, lut, t00);
    tex_fetch(sp, (int)fi + 1, (int)fj, mode, lut, t10);
    tex_fetch(sp, (int)fi, (int)fj + 1, mode, lut, t01);
    tex_fetch(sp, (int)fi + 1, (int)fj + 1, mode, lut, t11);
#pragma unroll
    for (int k = 0; k < 4; k++)
        c[k] = ((1.0f - al) * (1.0f - be)) * t00[k] + (al * (1.0f - be)) * t10[k] + ((1.0f - al) * be) * t01[k] + (al * be) * t11[k];
}
__device__ __forceinline__ void eval_paint(uint32_t pattern, const vkb_gradient *g, const float *gp, float W, float H, uint32_t solid, float opacity, float fx,
                                           float fy, float out[4], const float *lut, const vkb_surfpat *surfpats = nullptr, uint32_t slot = 0) {
    float c[4];
    if (pattern == VKB_PAT_LINEAR || pattern == VKB_PAT_RADIAL) eval_gradient(pattern, g, gp, fx / W, fy / H, c);  // out of line: keeps the solid-colour loop small
    else if (pattern == VKB_PAT_SURFACE) eval_surface(surfpats + slot, fx, fy, c, lut);
    else {
        c[0] = lut[solid & 0xFF];  // lut[i] == (float)i / 255.0f exactly
        c[1] = lut[(solid >> 8) & 0xFF];
        c[2] = lut[(solid >> 16) & 0xFF];
        c[3] = lut[solid >> 24];
    }
#pragma unroll
    for (int k = 0; k < 4; k++) out[k] = c[k] * opacity;
}
// the same with fx / W and fy / H supplied (the warp-per-tile kernel divides once per tile column / row, not per pixel and path)
__device__ __forceinline__ void eval_paint_q(uint32_t pattern, const vkb_gradient *g, const float *gp, uint32_t solid, float opacity, float fx, float fy, float qx,
                                             float qy, float out[4], const float *lut, const vkb_surfpat *surfpats, uint32_t slot) {
    float c[4];
    if (pattern == VKB_PAT_LINEAR || pattern == VKB_PAT_RADIAL) eval_gradient(pattern, g, gp, qx, qy, c);
    else if (pattern == VKB_PAT_SURFACE) eval_surface(surfpats + slot, fx, fy, c, lut);
    else {
        c[0] = lut[solid & 0xFF];
        c[1] = lut[(solid >> 8) & 0xFF];
        c[2] = lut[(solid >> 16) & 0xFF];
        c[3] = lut[solid >> 24];
    }
#pragma unroll
    for (int k = 0; k < 4; k++) out[k] = c[k] * opacity;
}
// UNORM8 store conversion: round to nearest of v*255, clamped (NaN -> 0)
// (one saturating truncating convert: F2IP.U8.F32.TRUNC; NaN and negatives give 0, >= 255.5 gives 255)
__device__ __forceinline__ uint32_t unorm8(float v) {
    uint32_t r;
    float    q = v * 255.0f + 0.5f;
    asm("cvt.rzi.sat.u8.f32 %0, %1;" : "=r"(r) : "f"(q));
    return r;
}
// premultiplied OVER per channel with UNORM8 store, src/vkvg_device_internal.c:203-209.  `lut[i]` holds the exactly
// rounded float i / 255.0f (a table look-up instead of an IEEE division per channel per sample).
__device__ __forceinline__ uint32_t blend_over(uint32_t dst, const float s[4], float ia, const float *lut) {
    uint32_t out = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        float d = lut[(dst >> (8 * k)) & 0xFF];
        float r = s[k] + d * ia;
        out |= unorm8(r) << (8 * k);
    }
    return out;
}

// ----------------------------------------------------------------------------------------------------
// fine pass, "row-delta" formulation.
//
// A tile holds 16*S sample rows (pixel row ly, sample s: all 16 samples of a row share one y and one x offset).
// For one edge and one sample row the H term is a step function of the pixel column c:
//     dy > 0:  +[c >= c0] - [crossing at or left of L]          dy < 0:  -[c >= c0] + [crossing at or left of L]
// with c0 = the first column whose sample is at or right of the crossing, and the V term is constant along the row.
// So instead of testing every sample against every edge (16*S*16 tests per edge), one lane per sample row finds c0
// exactly (integer binary search), adds +-1 to a per-column DELTA and folds the row constants into a BASE; a prefix sum
// over the 16 columns at the end gives the partial winding of every sample of the row.  The work is organised in TASKS
// (a path-tile, or a chunk of at most FINE_CH of its edges): each of the 8 warps of the block turns one task into
// 16*S rows x 16 biased int8 partial windings in shared memory (phase A); then every thread, which owns one pixel,
// walks the tasks in order, adds its samples' partials into integer accumulators and, at the last task of a path-tile,
// applies the fill rule and blends (phase B).  Winding stays exact: same predicates as edge_winding() in the oracle.
// ----------------------------------------------------------------------------------------------------
#define FINE_SLOTS 8     // tasks per group = warps per block
#define FINE_CH 32       // edges per task: |partial winding| <= 3*FINE_CH = 96 fits the biased int8
#define FINE_STRIDE 20   // bytes per sample row in shared memory (16 columns + pad: lanes hit distinct banks)
#define FINE_DBIAS 40u   // bias of the packed per-column deltas (4 per register): bytes stay in [8, 72]

template <int S> __device__ __forceinline__ void sample_pos16(int s, int32_t &x16, int32_t &y16) {
    // (runtime sample index: a select chain over the compile-time tables; used once per lane per pass)
    int x = 0, y = 0;
#pragma unroll
    for (int k = 0; k < S; k++)
        if (k == s) { x = SamplePos<S>::x(k); y = SamplePos<S>::y(k); }
    x16 = x * 16; y16 = y * 16;
}

struct FineTask {
    uint32_t  eoff;   // first edge in tile_edges
    int32_t   n;      // edges in this task (0: backdrop-only)
    int32_t   draw;
    int32_t   bd;     // backdrop, added by the first task of a path-tile
    uint32_t  last;   // 1: the path-tile is complete after this task
    vkb_paint paint;
};

// sign of (2*q - r) without forming it: > 0 <=> q > floor(r/2), < 0 <=> q < ceil(r/2)
template <class T> __device__ __forceinline__ bool twice_gt(T q, int32_t r) { return q > (T)(r >> 1); }
template <class T> __device__ __forceinline__ bool twice_lt(T q, int32_t r) { return q < (T)((r + 1) >> 1); }

__device__ __forceinline__ void delta_add(uint32_t (&d)[4], int c0, int sgn) {  // column c0 (1..15) += sgn; c0 = 16: none
    const uint32_t inc = (uint32_t)sgn << ((c0 & 3) * 8);
    const int      t   = c0 >> 2;
    d[0] += t == 0 ? inc : 0u;
    d[1] += t == 1 ? inc : 0u;
    d[2] += t == 2 ? inc : 0u;
    d[3] += t == 3 ? inc : 0u;
}

// One edge (tile-relative 24.8 coordinates; L at x = 1/2, C at (1/2, 1/2)) against P sample rows of this lane.
// T = int32_t when every coordinate lies in [-16384, 20479] (products < 7.6e8, sums < 1.6e9), else long long.
template <int P, class T>
__device__ __forceinline__ void row_edge(int32_t ax, int32_t ay, int32_t bx, int32_t by, bool crossL, const int32_t (&ry)[P], const int32_t (&rxo)[P],
                                         uint32_t (&d)[P][4], int32_t (&base)[P]) {
    const int32_t dx = bx - ax, dy = by - ay;
    const T       K  = (T)dy * ax - (T)dx * ay;  // m(y) = dx*y + K ;  2*E(L,y) = 2*m - dy ;  E(sample) = m - dy*sx
    if (dy != 0) {
        const int     sgn = dy > 0 ? 1 : -1;
        const int32_t h   = dy > 0 ? (dy >> 1) : ((dy + 1) >> 1);
        const T       D   = (T)256 * (dy > 0 ? dy : -dy);
#pragma unroll
        for (int j = 0; j < P; j++) {
            const int32_t y = ry[j];
            if ((ay <= y) != (by <= y)) {
                const T m = (T)dx * y + K;
                // at or left of L: dy > 0: m <= h, dy < 0: m >= h
                if (dy > 0 ? (m <= (T)h) : (m >= (T)h)) base[j] -= sgn;
                // first column c with E(sample c) on the "left of sample" side: 256*|dy|*c >= qq
                T qq = m - (T)dy * rxo[j];
                if (dy < 0) qq = -qq;
                if (qq <= 0) base[j] += sgn;
                else {
                    int c = 0;
                    T   Dc = 0;
                    if (D * 8 < qq) { c = 8; Dc = D * 8; }
                    if (Dc + D * 4 < qq) { c += 4; Dc += D * 4; }
                    if (Dc + D * 2 < qq) { c += 2; Dc += D * 2; }
                    if (Dc + D < qq) { c += 1; }
                    delta_add(d[j], c + 1, sgn);  // c = last column still left of the crossing
                }
            }
        }
    }
    if (crossL) {  // V term: crossings of the vertical segment from C down to (L, y)
        const bool    tie = dy == 0 || ((dx > 0) != (dy > 0));
        const int32_t rc  = dy - dx;  // 2*E(C) = 2*K - rc
        if (dx > 0) {
            const int belowC = twice_gt<T>(K, rc) || (!twice_lt<T>(K, rc) && tie);
#pragma unroll
            for (int j = 0; j < P; j++) {
                const T m = (T)dx * ry[j] + K;
                base[j] -= (int)(twice_gt<T>(m, dy) || (!twice_lt<T>(m, dy) && tie)) - belowC;
            }
        } else {
            const int belowC = twice_lt<T>(K, rc) || (!twice_gt<T>(K, rc) && tie);
#pragma unroll
            for (int j = 0; j < P; j++) {
                const T m = (T)dx * ry[j] + K;
                base[j] += (int)(twice_lt<T>(m, dy) || (!twice_gt<T>(m, dy) && tie)) - belowC;
            }
        }
    }
}

// The same edge against the same rows, but as MASKS for the bit-sliced winding counters of fine_warp_k (bit 16 * j + c = sample row j of
// this lane, column c):
//   hm   columns at or right of the crossing of each row the edge spans, when that crossing lies right of L (a crossing at or left of L is
//        already counted by the backdrop and the V term: its two H contributions cancel) - winding += sign(dy) there;
//   vm   rows whose V term differs from the one of C - winding += (vneg ? -1 : +1) in all their columns.
// Same exact predicates as row_edge above (which stays for the COUNT-rule / very long lists).
// Returned packed in 64 bits (hm | rows of vm << 32 | vneg << 34) so that the out-of-line int64 variant hands it back in registers.
template <int P, class T>
__device__ __forceinline__ unsigned long long row_edge_masks(int32_t ax, int32_t ay, int32_t bx, int32_t by, bool crossL, int32_t ry0, int32_t ry1, int32_t rxo0, int32_t rxo1) {
    const int32_t ry[2] = {ry0, ry1}, rxo[2] = {rxo0, rxo1};
    const int32_t dx = bx - ax, dy = by - ay;
    const T       K  = (T)dy * ax - (T)dx * ay;
    uint32_t      hm = 0, vrows = 0, vneg = 0;
    if (dy != 0) {
        const int32_t h = dy > 0 ? (dy >> 1) : ((dy + 1) >> 1);
        const T       D = (T)256 * (dy > 0 ? dy : -dy);
#pragma unroll
        for (int j = 0; j < P; j++) {
            const int32_t y = ry[j];
            if ((ay <= y) != (by <= y)) {
                const T m = (T)dx * y + K;
                if (!(dy > 0 ? (m <= (T)h) : (m >= (T)h))) {
                    T qq = m - (T)dy * rxo[j];
                    if (dy < 0) qq = -qq;
                    uint32_t cm = 0xFFFFu;  // crossing between L and the first sample: every column
                    if (qq > 0) {
                        int c = 0;
                        T   Dc = 0;
                        if (D * 8 < qq) { c = 8; Dc = D * 8; }
                        if (Dc + D * 4 < qq) { c += 4; Dc += D * 4; }
                        if (Dc + D * 2 < qq) { c += 2; Dc += D * 2; }
                        if (Dc + D < qq) { c += 1; }
                        cm = (0xFFFEu << c) & 0xFFFFu;  // c = last column still left of the crossing
                    }
                    hm |= cm << (16 * j);
                }
            }
        }
    }
    if (crossL) {
        const bool    tie = dy == 0 || ((dx > 0) != (dy > 0));
        const int32_t rc  = dy - dx;
        if (dx > 0) {
            const bool belowC = twice_gt<T>(K, rc) || (!twice_lt<T>(K, rc) && tie);
#pragma unroll
            for (int j = 0; j < P; j++) {
                const T    m    = (T)dx * ry[j] + K;
                const bool cond = twice_gt<T>(m, dy) || (!twice_lt<T>(m, dy) && tie);
                if (cond != belowC) vrows |= 1u << j;
            }
            vneg = !belowC;
        } else {
            const bool belowC = twice_lt<T>(K, rc) || (!twice_gt<T>(K, rc) && tie);
#pragma unroll
            for (int j = 0; j < P; j++) {
                const T    m    = (T)dx * ry[j] + K;
                const bool cond = twice_lt<T>(m, dy) || (!twice_gt<T>(m, dy) && tie);
                if (cond != belowC) vrows |= 1u << j;
            }
            vneg = belowC;
        }
    }
    return (unsigned long long)hm | ((unsigned long long)(vrows | (vneg << 2)) << 32);
}
// edges far from the tile (64-bit products): rare, kept out of line so that the hot loop stays small
template <int P>
__device__ __noinline__ unsigned long long row_edge_masks_far(int32_t ax, int32_t ay, int32_t bx, int32_t by, bool crossL, int32_t ry0, int32_t ry1, int32_t rxo0, int32_t rxo1) {
    return row_edge_masks<P, long long>(ax, ay, bx, by, crossL, ry0, ry1, rxo0, rxo1);
}

// Bit-sliced winding counters: plane p holds bit p of the (two's complement) winding of the 16 columns of this lane's sample rows.  Adding
// +-1 to the columns of `mask` is a ripple carry over the planes: two LOP3 per plane for all 16 * P samples at once, against one byte-wise
// add per sample row for the packed deltas (and the coverage of the path-tile is an OR of the planes instead of a prefix sum per row).
// mode 0: parity only (even-odd); 1: four planes (|winding| <= 15); 2: eight planes (|winding| <= 255).  neg and mode are warp-uniform.
__device__ __forceinline__ void plane_add(uint32_t (&pl)[8], uint32_t mask, bool neg, int mode) {
    if (mode == 0) { pl[0] ^= mask; return; }
    const uint32_t s = neg ? 0xffffffffu : 0u;
    uint32_t       c = mask;
#pragma unroll
    for (int p = 0; p < 4; p++) { const uint32_t t = (pl[p] ^ s) & c; pl[p] ^= c; c = t; }
    if (mode == 2) {
#pragma unroll
        for (int p = 4; p < 8; p++) { const uint32_t t = (pl[p] ^ s) & c; pl[p] ^= c; c = t; }
    }
}

// Premultiplied OVER on packed bytes for a source whose channels are exact multiples of 1/255 (a solid colour at opacity 1) with every
// colour channel <= alpha: r = S + round(D * (255 - A) / 255) per channel, two channels per 32-bit multiply, the division by 255 as
// (v + (v >> 8)) >> 8 with v = D * IA + 128.  Identical to blend_over for all 2^24 (S, A, D) with S <= A: the fp32 chain there errs by
// < 1e-4 while D * IA / 255 + 1/2 is never closer than 1/510 to an integer (tests/test_blend_int.py enumerates every case).
__device__ __forceinline__ uint32_t blend_int(uint32_t dst, uint32_t s_lo, uint32_t s_hi, uint32_t IA) {
    uint32_t x = __byte_perm(dst, 0, 0x4240) * IA + 0x00800080u;  // R, B
    uint32_t y = __byte_perm(dst, 0, 0x4341) * IA + 0x00800080u;  // G, A
    x += __byte_perm(x, 0, 0x4341);
    y += __byte_perm(y, 0, 0x4341);
    x = __byte_perm(x, 0, 0x4341) + s_lo;
    y = __byte_perm(y, 0, 0x4341) + s_hi;
    return __byte_perm(x, y, 0x6240);
}

// Warp 0 turns the next path-tiles of the tile (headers prefetched in nh0 / nh1, one per lane) into at most FINE_SLOTS
// tasks, advances the cursor (s_p = path-tile, s_k = edges of it already consumed) and publishes the task count.
__device__ __forceinline__ void fine_build_group(uint32_t lane, uint32_t end, uint32_t &s_p, uint32_t &s_k, uint32_t &s_nslots, FineTask *tasks,
                                                 const int4 &nh0, const int4 &nh1) {
    const uint32_t p0 = s_p, k0 = s_k;
    __syncwarp();
    const uint32_t p  = p0 + lane;
    const bool     ok = lane < FINE_SLOTS && p < end;
    const int4     h  = nh0, hp = nh1;
    const int      kk = lane == 0 ? (int)k0 : 0;  // edges of path-tile p0 consumed by earlier groups
    // chunk size of a path-tile: one task up to 16 edges; up to 128 edges chunks of <= 16; beyond that a whole
    // number of groups of FINE_SLOTS equal chunks (<= FINE_CH) so that no warp idles while one finishes a long list
    int chunk = FINE_CH;
    if (h.w <= 128) { const int q = max(1, (h.w + 15) / 16); chunk = max(1, (h.w + q - 1) / q); }
    else { const int q = FINE_SLOTS * ((h.w + FINE_SLOTS * FINE_CH - 1) / (FINE_SLOTS * FINE_CH)); chunk = (h.w + q - 1) / q; }
    const int      nt = ok ? max(1, (h.w - kk + chunk - 1) / chunk) : 0;
    int            incl = nt;
#pragma unroll
    for (int o = 1; o < FINE_SLOTS; o <<= 1) {
        int v = __shfl_up_sync(0xffffffffu, incl, o);
        if ((int)lane >= o) incl += v;
    }
    const int start = incl - nt;
    for (int t = 0; t < nt && start + t < FINE_SLOTS; t++) {
        FineTask ft;
        const int e0 = kk + t * chunk;
        ft.eoff = (uint32_t)h.z + (uint32_t)e0;
        ft.n    = min(chunk, h.w - e0);
        ft.draw = h.x;
        ft.bd   = e0 == 0 ? h.y : 0;
        ft.last = e0 + ft.n >= h.w ? 1u : 0u;
        ft.paint = vkb_paint{(uint32_t)hp.x, (uint32_t)hp.y, __int_as_float(hp.z), (uint32_t)hp.w};
        tasks[start + t] = ft;
    }
    const int total = __shfl_sync(0xffffffffu, incl, FINE_SLOTS - 1);
    // cursor after this group: the path-tile that owns slot FINE_SLOTS-1 (or the first one not started)
    if (ok && start < FINE_SLOTS && start + nt >= FINE_SLOTS) {
        const int done = FINE_SLOTS - start;             // tasks of this path-tile issued in this group
        const int e1   = kk + done * chunk;
        if (e1 >= h.w) { s_p = p + 1; s_k = 0; }
        else { s_p = p; s_k = (uint32_t)e1; }
    }
    if (lane == 0) {
        s_nslots = (uint32_t)min(total, FINE_SLOTS);
        if (total < FINE_SLOTS) { s_p = end; s_k = 0; }
    }
}

// Per-sample stencil bytes packed four to a word (sample s = byte s & 3 of word s >> 2): bit 1 = clipped out, bits 2-7 =
// save levels — the reference's stencil attachment layout (src/vkvg_device_internal.h:28-30) minus its transient FILL bit.
// Clip / save / restore are whole-surface stencil passes there (pipelineClipping, src/vkvg_device_internal.c:233-239;
// src/vkvg_context.c:754-795, :1320-1336, :1402-1418); here they are ordinary entries of the per-tile draw lists.
template <int SW> __device__ __forceinline__ void stencil_word_op(uint32_t (&st)[SW], uint32_t rule, uint32_t arg) {
    const uint32_t bit = arg & 0xFF, sh = (arg >> 8) & 0xFF;
#pragma unroll
    for (int k = 0; k < SW; k++) {
        uint32_t w = st[k];
        if (rule == VKB_RULE_ST_CLEAR) w = 0u;
        else if (rule == VKB_RULE_ST_SAVE) w = (w & ~(bit * 0x01010101u)) | (((w & 0x02020202u) >> 1) * bit);
        else w = (w & ~0x02020202u) | ((((w & (bit * 0x01010101u)) >> sh) & 0x01010101u) << 1);
        st[k] = w;
    }
}

template <int S, bool CAPTURE, bool CLIP> __global__ void __launch_bounds__(256) fine_k(FineArgs a) {
    constexpr int ROWS   = 16 * S;
    constexpr int P      = ROWS >= 64 ? 2 : 1;                 // sample rows per lane per pass
    constexpr int PASSES = (ROWS + 32 * P - 1) / (32 * P);
    const uint32_t tile  = a.tile_lo + blockIdx.x;  // (a band of tile rows when the surface has a read-back target, else all of them)
    if (a.counts->overflow) return;  // some intermediate did not fit this attempt's buffers: the host replays the batch (vkb_counts)
    const uint32_t first = a.tile_first[tile], end = a.tile_end[tile];
    if (first == end) return;  // no draw of this batch touches the tile: the stored pixels stay as they are

    __shared__ float    lut[256];
    __shared__ uint32_t cnt[FINE_SLOTS * ROWS * (FINE_STRIDE / 4)];
    __shared__ FineTask tasks[FINE_SLOTS];
    __shared__ uint32_t s_p, s_k, s_nslots;

    lut[threadIdx.x] = (float)threadIdx.x / 255.0f;
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t tx = tile % a.sd.tiles_x, ty = tile / a.sd.tiles_x;
    // warp 0 keeps the headers of the next group in registers, fetched while the current group is processed
    int4 nh0 = make_int4(0, 0, 0, 0), nh1 = nh0;
    if (warp == 0 && lane < FINE_SLOTS && first + lane < end) { nh0 = a.hdr[2 * (first + lane)]; nh1 = a.hdr[2 * (first + lane) + 1]; }
    // a warp owns an 8x4 pixel block (more often wholly inside or wholly outside a shape than a 16x2 strip)
    const uint32_t lx = (lane & 7) + 8 * (warp & 1), ly = (lane >> 3) + 4 * (warp >> 1);
    const uint32_t px = tx * VKB_TILE + lx, py = ty * VKB_TILE + ly;
    const bool     inside = px < a.sd.width && py < a.sd.height;
    const size_t   pix    = (size_t)py * a.sd.width + px;
    const size_t   mspix  = ((size_t)tile * 256 + threadIdx.x) * S;  // per-sample plane is tile-major, in thread order
    const int32_t  X0 = (int32_t)tx * VKB_TILE_FX, Y0 = (int32_t)ty * VKB_TILE_FX;
    const uint32_t band_y0 = a.sd.band_tiles ? (ty / a.sd.band_tiles) * a.sd.band_tiles * VKB_TILE : 0u;  // first pixel row of this tile's canvas

    uint32_t col[S];
    {
        bool ms = false;
        if (!a.dst_is_clear && a.tile_ms[tile]) ms = (a.ms_mask[tile * 8 + warp] >> lane) & 1u;
        if (ms) {  // this pixel's samples differed after an earlier flush
#pragma unroll
            for (int s = 0; s < S; s++) col[s] = a.ms_image[mspix + s];
        } else {
            uint32_t c = (!a.dst_is_clear && inside) ? a.image[pix] : 0u;
#pragma unroll
            for (int s = 0; s < S; s++) col[s] = c;
        }
    }
    constexpr int SW = (S + 3) / 4;
    uint32_t      stw[SW];
#pragma unroll
    for (int k = 0; k < SW; k++) stw[k] = (CLIP && a.stencil_in) ? a.stencil[((size_t)tile * 256 + threadIdx.x) * SW + k] : 0u;
    int32_t wacc[S];
#pragma unroll
    for (int s = 0; s < S; s++) wacc[s] = 0;
    if (threadIdx.x == 0) { s_p = first; s_k = 0; }
    __syncthreads();

    for (;;) {
        // ---- build the next group of tasks (warp 0, one path-tile per lane) ----
        if (warp == 0) fine_build_group(lane, end, s_p, s_k, s_nslots, tasks, nh0, nh1);
        __syncthreads();
        const uint32_t nslots = s_nslots;
        if (warp == 0) {  // prefetch the headers the next group starts from (cursor already advanced)
            const uint32_t p = s_p + lane;
            if (lane < FINE_SLOTS && p < end) { nh0 = a.hdr[2 * p]; nh1 = a.hdr[2 * p + 1]; }
        }

        // ---- phase A: warp w turns task w into per-row partial windings ----
        if (warp < nslots && tasks[warp].n > 0) {
            const FineTask  ft = tasks[warp];
            const vkb_edge *ep = a.tile_edges + ft.eoff;
#pragma unroll 1
            for (int ps = 0; ps < PASSES; ps++) {
                int32_t  ry[P], rxo[P], base[P];
                uint32_t d[P][4];
                int      row[P];
#pragma unroll
                for (int j = 0; j < P; j++) {
                    row[j] = ps * 32 * P + j * 32 + (int)lane;   // rows of one lane are 32 apart: conflict-free stores
                    const int r = min(row[j], ROWS - 1);
                    int32_t   x16, y16;
                    sample_pos16<S>(r % S, x16, y16);
                    ry[j]  = (r / S) * 256 + y16;
                    rxo[j] = x16;
                    base[j] = 0;
#pragma unroll
                    for (int w4 = 0; w4 < 4; w4++) d[j][w4] = FINE_DBIAS * 0x01010101u;
                }
                for (int k = 0; k < ft.n; k++) {
                    const int4    ev = __ldg((const int4 *)(ep + k));
                    const int32_t ax = ev.x - X0, ay = ev.y - Y0, bx = ev.z - X0, by = ev.w - Y0;
                    const bool    crossL = (ax <= 0) != (bx <= 0);
                    const bool    near = (uint32_t)(ax + 16384) < 36864u && (uint32_t)(ay + 16384) < 36864u && (uint32_t)(bx + 16384) < 36864u &&
                                         (uint32_t)(by + 16384) < 36864u;
                    if (near) row_edge<P, int32_t>(ax, ay, bx, by, crossL, ry, rxo, d, base);
                    else row_edge<P, long long>(ax, ay, bx, by, crossL, ry, rxo, d, base);
                }
#pragma unroll
                for (int j = 0; j < P; j++) {
                    if (row[j] < ROWS) {
                        int       run = 128 + base[j];
                        uint32_t *dst = cnt + ((size_t)warp * ROWS + row[j]) * (FINE_STRIDE / 4);
#pragma unroll
                        for (int w4 = 0; w4 < 4; w4++) {
                            const uint32_t x = d[j][w4];
                            const int v0 = run + (int)(x & 0xFF) - (int)FINE_DBIAS;
                            const int v1 = v0 + (int)((x >> 8) & 0xFF) - (int)FINE_DBIAS;
                            const int v2 = v1 + (int)((x >> 16) & 0xFF) - (int)FINE_DBIAS;
                            const int v3 = v2 + (int)(x >> 24) - (int)FINE_DBIAS;
                            run          = v3;
                            dst[w4]      = (uint32_t)v0 | ((uint32_t)v1 << 8) | ((uint32_t)v2 << 16) | ((uint32_t)v3 << 24);
                        }
                    }
                }
            }
        }
        __syncthreads();

        // ---- phase B: every thread owns one pixel; tasks in order ----
        const uint8_t *cb = (const uint8_t *)cnt;
        for (uint32_t g = 0; g < nslots; g++) {
            const FineTask ft = tasks[g];
            if (ft.n > 0) {
#pragma unroll
                for (int s = 0; s < S; s++) wacc[s] += (int)cb[((size_t)g * ROWS + ly * S + s) * FINE_STRIDE + lx] - 128;
            }
#pragma unroll
            for (int s = 0; s < S; s++) wacc[s] += ft.bd;  // the backdrop rides on the first task of a path-tile (0 on the others)
            if (!ft.last) continue;
            int32_t w[S], wor = 0;
#pragma unroll
            for (int s = 0; s < S; s++) {
                w[s]    = wacc[s];
                wacc[s] = 0;
                wor |= w[s];
            }
            if (CAPTURE) {
                if ((uint32_t)ft.draw == a.winding_draw && inside) {
#pragma unroll
                    for (int s = 0; s < S; s++) a.winding_out[pix * S + s] = w[s];
                }
            }
            const vkb_paint pt   = ft.paint;
            const uint32_t  rule = pt.rule_pattern & 0xFF, pattern = (pt.rule_pattern >> 8) & 0xFF, op = (pt.rule_pattern >> 16) & 0xFF;
            if (CLIP && rule >= VKB_RULE_CLIP_EO) {  // stencil-only entries (warp uniform: the rule belongs to the draw)
                if (rule <= VKB_RULE_CLIP_NZ) {
#pragma unroll
                    for (int s = 0; s < S; s++) {
                        const bool in = rule == VKB_RULE_CLIP_EO ? (w[s] & 1) != 0 : w[s] != 0;
                        if (!in) stw[s >> 2] |= VKB_STENCIL_CLIP << (8 * (s & 3));
                    }
                } else stencil_word_op<SW>(stw, rule, pt.color);
                continue;
            }
            if (!__any_sync(0xffffffffu, wor != 0)) continue;  // nothing of this draw reaches the pixel block of this warp
            int32_t         n[S], nmax = 0;
            bool            uni = true, two = true;
            if (rule == VKB_RULE_EVEN_ODD) {
#pragma unroll
                for (int s = 0; s < S; s++) n[s] = w[s] & 1;
            } else if (rule == VKB_RULE_NON_ZERO) {
#pragma unroll
                for (int s = 0; s < S; s++) n[s] = w[s] != 0;
            } else {
#pragma unroll
                for (int s = 0; s < S; s++) n[s] = abs(w[s]);
            }
            if (CLIP) {
#pragma unroll
                for (int s = 0; s < S; s++)
                    if ((stw[s >> 2] >> (8 * (s & 3))) & VKB_STENCIL_CLIP) n[s] = 0;  // stencil test of every colour pipeline: CLIP bit clear
            }
#pragma unroll
            for (int s = 0; s < S; s++) {
                nmax = max(nmax, n[s]);
                uni  = uni && col[s] == col[0];
            }
            if (__any_sync(0xffffffffu, nmax != 0)) {
                float src[4];
                eval_paint(pattern, a.grads + pt.gradient, a.gprep + (size_t)pt.gradient * VKB_GPREP_FLOATS, (float)a.sd.width, (float)a.sd.full_height, pt.color, pt.opacity, (float)px + 0.5f,
                           (float)(py + a.sd.origin_y - band_y0) + 0.5f, src, lut, a.surfpats, pt.gradient);
                // every operator is r = src + dst * ia per channel with UNORM8 store (negatives and > 1 saturate): OVER ia = 1 - a,
                // DIFFERENCE (blend op SUBTRACT) ia = -(1 - a), CLEAR (logic op CLEAR) src = 0, ia = 0
                float ia = 1.0f - src[3];
                if (op == VKB_OP_SUB) ia = -ia;
                else if (op == VKB_OP_CLEAR) { src[0] = src[1] = src[2] = src[3] = 0.0f; ia = 0.0f; }
                if (ia == 0.0f) {  // the result does not depend on the destination: repeating the blend changes nothing
                    nmax = 1;
#pragma unroll
                    for (int s = 0; s < S; s++) n[s] = n[s] ? 1 : 0;
                }
#pragma unroll
                for (int s = 0; s < S; s++) two = two && (n[s] == 0 || n[s] == nmax);
                // warp-uniform choice (a divergent one would make the warp execute both variants)
                const bool opaque = ia == 0.0f;  // dst * 0 vanishes: the result does not depend on what a sample holds
                if (__all_sync(0xffffffffu, (uni || opaque) && two)) {  // one result per pixel: same colour in every sample (or an opaque source), blended nmax times or not at all
                    uint32_t c = col[0];
                    for (int32_t r = 0; r < nmax; r++) c = blend_over(c, src, ia, lut);
#pragma unroll
                    for (int s = 0; s < S; s++) col[s] = n[s] ? c : col[s];
                } else {
#pragma unroll
                    for (int s = 0; s < S; s++)
                        for (int32_t r = 0; r < n[s]; r++) col[s] = blend_over(col[s], src, ia, lut);
                }
            }
        }
        if (s_p >= end) break;   // (written before the barrier that precedes phase A: stable here)
        __syncthreads();         // phase B reads of cnt / tasks are done before the next group overwrites them
    }

    if (CLIP) {
#pragma unroll
        for (int k = 0; k < SW; k++) a.stencil[((size_t)tile * 256 + threadIdx.x) * SW + k] = stw[k];
    }
    bool differ = false;
#pragma unroll
    for (int s = 1; s < S; s++) differ = differ || col[s] != col[0];
    const uint32_t dmask = __ballot_sync(0xffffffffu, differ);
    const int tile_differs = __syncthreads_or(differ);
    if (differ) {
#pragma unroll
        for (int s = 0; s < S; s++) a.ms_image[mspix + s] = col[s];
    }
    if (tile_differs && lane == 0) a.ms_mask[tile * 8 + warp] = dmask;
    if (threadIdx.x == 0) a.tile_ms[tile] = tile_differs ? 1 : 0;
    if (inside) {
        uint32_t out = 0;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            uint32_t sum = 0;
#pragma unroll
            for (int s = 0; s < S; s++) sum += (col[s] >> (8 * k)) & 0xFF;
            out |= ((sum + S / 2) / S) << (8 * k);
        }
        a.image[pix] = out;
    }
}
// ----------------------------------------------------------------------------------------------------
// fine pass, analytic-coverage mode (sd.samples == 0): one colour per pixel, coverage = exact area.
//   A(pixel) = integral of the winding number over the pixel square = B + integral of V + integral of H, with the same
//   backdrop B (winding at C = (1/2, 1/2) fixed-point units inside the tile corner), the same L (x = 1/2 unit) and the same
//   binned edge lists as the MSAA pass:
//     V(y): crossings of L between C and y        -> per pixel row   (r + 1 - clamp(vc, r, r + 1)) - [vc < C.y]
//     H(x,y): crossings of the row between L and x -> the part of the edge right of L, clipped to the pixel row, adds the
//             area to the right of it inside each pixel it touches and its full height to every pixel further right;
//             stored as differences along the row so that one prefix sum per row yields all 16 areas (signed-area
//             accumulation as in font-rs / vello's fine stage).
//   coverage = min(|A|, 1) (NON_ZERO, strokes) or 1 - |A mod 2 - 1| (EVEN_ODD); the paint scaled by the coverage is
//   blended once.  Same definition as the oracle's area_brute / analytic_draw in oracle/vkvg_oracle.c (double precision,
//   whole surface, no tiles); float rounding here bounds the difference at ~1e-5 of a pixel for edges near the tile.
// A task is handled by one warp: lanes 0-15 take the even edges, lanes 16-31 the odd ones, one pixel row per lane.
// Areas are accumulated in 2^-20 fixed point: integer sums do not depend on the order in which the binning kernel's
// atomics happened to place the edges of a tile, so the output is deterministic.
// ----------------------------------------------------------------------------------------------------
#define FA_STRIDE 17
#define FA_ONE 1048576.0f
template <bool CAPTURE, bool CLIP> __global__ void __launch_bounds__(256) fine_analytic_k(FineArgs a) {
    const uint32_t tile  = a.tile_lo + blockIdx.x;  // (a band of tile rows when the surface has a read-back target, else all of them)
    if (a.counts->overflow) return;
    const uint32_t first = a.tile_first[tile], end = a.tile_end[tile];
    if (first == end) return;

    __shared__ float    lut[256];
    __shared__ int32_t  acc[FINE_SLOTS][2][16][FA_STRIDE];
    __shared__ FineTask tasks[FINE_SLOTS];
    __shared__ uint32_t s_p, s_k, s_nslots;

    lut[threadIdx.x] = (float)threadIdx.x / 255.0f;
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t tx = tile % a.sd.tiles_x, ty = tile / a.sd.tiles_x;
    int4 nh0 = make_int4(0, 0, 0, 0), nh1 = nh0;
    if (warp == 0 && lane < FINE_SLOTS && first + lane < end) { nh0 = a.hdr[2 * (first + lane)]; nh1 = a.hdr[2 * (first + lane) + 1]; }
    const uint32_t lx = (lane & 7) + 8 * (warp & 1), ly = (lane >> 3) + 4 * (warp >> 1);
    const uint32_t px = tx * VKB_TILE + lx, py = ty * VKB_TILE + ly;
    const bool     inside = px < a.sd.width && py < a.sd.height;
    const size_t   pix    = (size_t)py * a.sd.width + px;
    const int32_t  X0 = (int32_t)tx * VKB_TILE_FX, Y0 = (int32_t)ty * VKB_TILE_FX;
    const uint32_t band_y0 = a.sd.band_tiles ? (ty / a.sd.band_tiles) * a.sd.band_tiles * VKB_TILE : 0u;  // first pixel row of this tile's canvas

    uint32_t col  = (!a.dst_is_clear && inside) ? a.image[pix] : 0u;
    uint32_t stw[1] = {(CLIP && a.stencil_in) ? a.stencil[(size_t)tile * 256 + threadIdx.x] : 0u};  // one stencil byte per pixel in this mode
    int32_t  wacc = 0;
    if (threadIdx.x == 0) { s_p = first; s_k = 0; }
    __syncthreads();

    for (;;) {
        if (warp == 0) fine_build_group(lane, end, s_p, s_k, s_nslots, tasks, nh0, nh1);
        __syncthreads();
        const uint32_t nslots = s_nslots;
        if (warp == 0) {
            const uint32_t p = s_p + lane;
            if (lane < FINE_SLOTS && p < end) { nh0 = a.hdr[2 * p]; nh1 = a.hdr[2 * p + 1]; }
        }

        // ---- phase A: warp w accumulates the row differences of task w ----
        if (warp < nslots && tasks[warp].n > 0) {
            const int       n    = tasks[warp].n;
            const vkb_edge *ep   = a.tile_edges + tasks[warp].eoff;
            const int       half = (int)(lane >> 4), r = (int)(lane & 15);
            const float     rr   = (float)r;
            int32_t        *D    = &acc[warp][half][r][0];
#pragma unroll
            for (int c = 0; c < 16; c++) D[c] = 0;
            int32_t     base = 0;
            const float DL   = 1.0f / 512.0f;  // L and C sit half a fixed-point unit inside the tile corner
            for (int k = half; k < n; k += 2) {
                const int4    ev  = __ldg((const int4 *)(ep + k));
                const int32_t axi = ev.x - X0, bxi = ev.z - X0;
                float ax = (float)axi * (1.0f / 256.0f), ay = (float)(ev.y - Y0) * (1.0f / 256.0f);
                float bx = (float)bxi * (1.0f / 256.0f), by = (float)(ev.w - Y0) * (1.0f / 256.0f);
                const bool la = axi <= 0, lb = bxi <= 0;
                if (la != lb) {  // crosses L: V term, then keep the part right of L
                    const float vc = ay + (DL - ax) * __fdividef(by - ay, bx - ax);
                    const float sv = bx > ax ? -1.0f : 1.0f;
                    base += __float2int_rn(sv * ((rr + 1.0f - fminf(fmaxf(vc, rr), rr + 1.0f)) - (vc < DL ? 1.0f : 0.0f)) * FA_ONE);
                    if (la) { ax = DL; ay = vc; } else { bx = DL; by = vc; }
                } else if (la) continue;
                if (ay == by) continue;
                const bool  down = by > ay;
                const int   sgn = down ? 1 : -1;
                const float xt = down ? ax : bx, yt = down ? ay : by, xb = down ? bx : ax, yb = down ? by : ay;
                const float ys = fmaxf(yt, rr), ye = fminf(yb, rr + 1.0f);
                if (ye <= ys) continue;
                const float slope = __fdividef(xb - xt, yb - yt);
                const float xs = xt + (ys - yt) * slope, xe = xt + (ye - yt) * slope, h = ye - ys;
                const float xmin = fminf(xs, xe), xmax = fmaxf(xs, xe);
                if (xmin >= 16.0f) continue;
                const float rd = __fdividef(1.0f, fmaxf(xmax - xmin, 1e-20f));
                const int   c0 = max(0, (int)floorf(xmin)), c1 = min(15, (int)floorf(xmax));
                int         prev = 0;
                for (int c = c0; c <= c1; c++) {
                    const float hi = (float)(c + 1) - xmin, lo = (float)(c + 1) - xmax;
                    const float t0 = __saturatef(-lo * rd), t1 = __saturatef((hi - 1.0f) * rd);
                    const int   ar = __float2int_rn(h * (t1 + (1.0f - t0 - t1) * 0.5f * (fmaxf(lo, 0.0f) + fminf(hi, 1.0f))) * FA_ONE);
                    D[c] += sgn * (ar - prev);
                    prev = ar;
                }
                const int cn = c1 < c0 ? c0 : c1 + 1;
                if (cn < 16) D[cn] += sgn * (__float2int_rn(h * FA_ONE) - prev);
            }
            int32_t run = base;
#pragma unroll
            for (int c = 0; c < 16; c++) { run += D[c]; D[c] = run; }
        }
        __syncthreads();

        // ---- phase B: every thread owns one pixel; tasks in order ----
        for (uint32_t g = 0; g < nslots; g++) {
            const FineTask ft = tasks[g];
            if (ft.n > 0) wacc += acc[g][0][ly][lx] + acc[g][1][ly][lx];
            wacc += ft.bd * 1048576;
            if (!ft.last) continue;
            const float A = (float)wacc * (1.0f / FA_ONE);
            wacc = 0;
            if (CAPTURE) {
                if ((uint32_t)ft.draw == a.winding_draw && inside) a.winding_out[pix] = __float_as_int(A);
            }
            const vkb_paint pt   = ft.paint;
            const uint32_t  rule = pt.rule_pattern & 0xFF, pattern = (pt.rule_pattern >> 8) & 0xFF, op = (pt.rule_pattern >> 16) & 0xFF;
            float           cov;
            if (rule == VKB_RULE_EVEN_ODD) {
                const float t = A - 2.0f * floorf(A * 0.5f);
                cov = 1.0f - fabsf(t - 1.0f);
            } else cov = fminf(fabsf(A), 1.0f);
            if (CLIP && rule >= VKB_RULE_CLIP_EO) {
                if (rule <= VKB_RULE_CLIP_NZ) {  // inside where more than half of the pixel is covered by the clip path
                    const float t = A - 2.0f * floorf(A * 0.5f);
                    const float c = rule == VKB_RULE_CLIP_EO ? 1.0f - fabsf(t - 1.0f) : fminf(fabsf(A), 1.0f);
                    if (!(c > 0.5f)) stw[0] |= VKB_STENCIL_CLIP;
                } else stencil_word_op<1>(stw, rule, pt.color);
                continue;
            }
            if (CLIP && (stw[0] & VKB_STENCIL_CLIP)) cov = 0.0f;
            if (!__any_sync(0xffffffffu, cov > 0.0f)) continue;
            float src[4];
            eval_paint(pattern, a.grads + pt.gradient, a.gprep + (size_t)pt.gradient * VKB_GPREP_FLOATS, (float)a.sd.width, (float)a.sd.full_height, pt.color, pt.opacity, (float)px + 0.5f,
                       (float)(py + a.sd.origin_y - band_y0) + 0.5f, src, lut, a.surfpats, pt.gradient);
            if (cov > 0.0f) {
#pragma unroll
                for (int k = 0; k < 4; k++) src[k] *= cov;
                float ia = 1.0f - src[3];
                if (op == VKB_OP_SUB) ia = -ia;
                else if (op == VKB_OP_CLEAR) { src[0] = src[1] = src[2] = src[3] = 0.0f; ia = 1.0f - cov; }  // the covered part of the pixel is wiped
                col = blend_over(col, src, ia, lut);
            }
        }
        if (s_p >= end) break;
        __syncthreads();
    }
    if (CLIP) a.stencil[(size_t)tile * 256 + threadIdx.x] = stw[0];
    if (inside) a.image[pix] = col;
}
// ----------------------------------------------------------------------------------------------------
// fine pass, one WARP per tile (the variant launched for batches without clip state; fine_k above keeps the stencil and
// winding-capture variants).
//
// With a mean of 3-5 edges and ~80 covered pixels per path-tile (C2, C4, tiger) a block of 8 warps per tile spends its time
// on barriers and on every warp walking every task.  Here a tile belongs to one warp, which needs no barrier at all (and warps are
// persistent: each takes its next tile from a counter, so no warp slot waits for the slowest tile of a block):
//   * the tile's 256*S sample colours live in shared memory ([sample row][column]) for the whole list of path-tiles;
//   * winding: lanes are sample rows (row_edge above, unchanged: same exact predicates), one chunk of <= FW_CH (else FW_CH_LONG) edges at a
//     time, edges fetched one per lane and handed round with shuffles; a path-tile that fits one chunk never leaves the
//     registers, longer ones (and COUNT-rule draws with a translucent source, which blend |winding| times) accumulate
//     per-sample int32 windings in shared memory;
//   * the covered PIXELS of the path-tile are compacted into a queue (each lane pushes its share of the set bits), and the warp
//     blends them 32 at a time: the paint is evaluated once per pixel, the covered samples of the pixel are blended with the
//     reference's per-sample arithmetic (blend_over).  Lanes are pixels that need work, not pixels that might.
// Same results as fine_k bit for bit (tests/test_gpu_parity.py runs both on the same scenes).
// ----------------------------------------------------------------------------------------------------
// ----------------------------------------------------------------------------------------------------
// fine_warp_k, winding: (edge, sample row) WORK ITEMS.
// With one lane per sample row every lane walks every edge of the list, although a short edge (a flattened curve, a piece of a
// round join: the bulk of C3 / C4 / tiger) spans a handful of the tile's 16 * S sample rows: the C3 capture of the previous version
// showed 5 of 32 lanes active in the crossing code.  Here the 32 edges of a chunk are first looked at by one lane EACH (row range,
// per-edge constants), the (edge, row) pairs are numbered by a prefix sum, and the warp then works through them 32 at a time: every
// lane finds its pair (binary search over the prefix sums with shuffles), computes the exact first covered column with the same
// integer predicates as row_edge, and adds +-1 there to the row's packed per-column deltas in shared memory (one ATOMS per crossing).
// The V term (a constant per row) is evaluated by the row lanes only for the edges that cross L.  At the end of the list each row
// lane turns the deltas of its rows into coverage bits: a parity word (even-odd) or biased int16 walked column by column (non-zero, COUNT
// rule; lists of more than 16000 edges fold them into this warp's int32 plane every 16000 edges).  Lists of at most FW_CH edges do not
// come here: for them one lane per row with bit-sliced counters (plane_add) is cheaper than the fixed cost of numbering the pairs.
// Sample rows are numbered in order of y for this: y(r) = FW_YSTEP * r + FW_YOFF (the standard sample positions are evenly spaced
// in y), which turns "rows the edge spans" into two shifts.
// ----------------------------------------------------------------------------------------------------
// sample positions as packed nibbles (sample s = nibble s): no select chain, cheap enough to recompute anywhere
template <int S> struct SamplePack;
template <> struct SamplePack<1> { static constexpr unsigned long long X = 0x8ull, Y = 0x8ull; };
template <> struct SamplePack<2> { static constexpr unsigned long long X = 0x4Cull, Y = 0x4Cull; };
template <> struct SamplePack<4> { static constexpr unsigned long long X = 0xA2E6ull, Y = 0xEA62ull; };

#define FW_WSTRIDE 17   // words per sample row of the winding plane (odd: lanes = rows hit distinct banks)
#define FW_DL_STRIDE 8  // words per sample row of the delta plane: 16 x int16 (no padding: 8 blocks of 4 warps then fit the shared memory of an SM)
#define FW_CH 31        // lists of up to FW_CH edges take the one-lane-per-row path with bit-sliced counters, longer ones the work items (measured: 15 / 31 / 63 / 127 give C2 1.81 / 1.57 / 1.54 / 1.54 ms, C4 6.61 / 5.71 / 5.58 / 5.58 ms, C3 0.48 / 0.51 / 0.60 / 0.74 ms)
#define FW_HBIAS 0x4000u  // bias of the packed int16: room for 16383 crossings of either sign in one column of one row
#define FW_SEG_CHUNKS 500  // lists of more than 16000 edges are folded into the int32 plane every 500 chunks of 32
template <int S> struct FwRows;  // sample rows in order of y: y(r) = STEP * r + OFF (tile-relative 24.8), sample index of row r
template <> struct FwRows<1> { static constexpr int LOG = 8, OFF = 128; static __device__ __forceinline__ int smp(int) { return 0; } };
template <> struct FwRows<2> { static constexpr int LOG = 7, OFF = 64;  static __device__ __forceinline__ int smp(int r) { return 1 - (r & 1); } };
template <> struct FwRows<4> { static constexpr int LOG = 6, OFF = 32;  static __device__ __forceinline__ int smp(int r) { return r & 3; } };

// One chunk of <= 32 edges (this lane's edge: tile-relative a -> b, `valid` when the lane has one).  T = int32_t when every edge of the
// chunk lies within [-16384, 20479] of the tile origin (products < 7.6e8, sums < 1.6e9, as in row_edge), else long long.
// mode 0: parity words, 2: packed biased int16.  base[j] accumulates the V term of this lane's sample rows.
template <int S, int P, class T>
__device__ __forceinline__ void fw_chunk(uint32_t lane, bool valid, int32_t ax, int32_t ay, int32_t bx, int32_t by, const int32_t (&ry)[P], int mode,
                                         uint32_t *dl, int32_t (&base)[P]) {
    constexpr uint32_t FULL = 0xffffffffu;
    constexpr int      ROWS = 16 * S;
    const int32_t dx = bx - ax, dy = by - ay;
    const T       K  = (T)dy * ax - (T)dx * ay;  // m(y) = dx*y + K ;  2*E(L,y) = 2*m - dy ;  E(sample) = m - dy*sx
    // ---- V term: crossings of the vertical segment from C down to (L, y).  Only edges that cross L and lie on the far side of C ----
    {
        bool vact = false;
        if (valid && ((ax <= 0) != (bx <= 0))) {
            const bool    tie = dy == 0 || ((dx > 0) != (dy > 0));
            const int32_t rc  = dy - dx;  // 2*E(C) = 2*K - rc
            const bool belowC = dx > 0 ? (twice_gt<T>(K, rc) || (!twice_lt<T>(K, rc) && tie)) : (twice_lt<T>(K, rc) || (!twice_gt<T>(K, rc) && tie));
            vact = !belowC;  // (the predicate is monotone in y: once true at C it is true for every row, and the term vanishes)
        }
        uint32_t vmask = __ballot_sync(FULL, vact);
        while (vmask) {
            const int k = __ffs((int)vmask) - 1;
            vmask &= vmask - 1u;
            const int32_t edx = __shfl_sync(FULL, dx, k), edy = __shfl_sync(FULL, dy, k);
            const T       eK  = __shfl_sync(FULL, K, k);
            const bool    tie = edy == 0 || ((edx > 0) != (edy > 0));
#pragma unroll
            for (int j = 0; j < P; j++) {
                const T m = (T)edx * ry[j] + eK;
                if (edx > 0) base[j] -= (int)(twice_gt<T>(m, edy) || (!twice_lt<T>(m, edy) && tie));
                else base[j] += (int)(twice_lt<T>(m, edy) || (!twice_gt<T>(m, edy) && tie));
            }
        }
    }
    // ---- H term: one work item per (edge, spanned sample row) ----
    int ra = 0, n_i = 0;
    if (valid && dy != 0) {  // rows with min(ay, by) <= y(r) < max(ay, by)  (the "spans" test of row_edge)
        const int32_t lo = min(ay, by), hi = max(ay, by);
        const int32_t STEPM1 = (1 << FwRows<S>::LOG) - 1;
        ra = min(max((lo - FwRows<S>::OFF + STEPM1) >> FwRows<S>::LOG, 0), ROWS);
        const int rb = min(max((hi - FwRows<S>::OFF + STEPM1) >> FwRows<S>::LOG, 0), ROWS);
        n_i = rb - ra;
    }
    const int incl = warp_incl_scan(n_i), total = __shfl_sync(FULL, incl, 31), start = incl - n_i;
    for (int t0 = 0; t0 < total; t0 += 32) {
        const int t = t0 + (int)lane;
        int       i = 0;  // the edge whose items include t: the last one with start <= t
#pragma unroll
        for (int b = 16; b >= 1; b >>= 1) {
            const int s = __shfl_sync(FULL, start, i + b);
            if (s <= t) i += b;
        }
        const int     r   = __shfl_sync(FULL, ra, i) + (t - __shfl_sync(FULL, start, i));
        const int32_t edx = __shfl_sync(FULL, dx, i), edy = __shfl_sync(FULL, dy, i);
        const T       eK  = __shfl_sync(FULL, K, i);
        if (t < total) {
            const int     smp = FwRows<S>::smp(r);
            const int32_t y   = (r << FwRows<S>::LOG) + FwRows<S>::OFF;
            const int32_t rxo = 16 * (int32_t)((SamplePack<S>::X >> (4 * smp)) & 15);
            const T       m   = (T)edx * y + eK;
            const int32_t h   = edy > 0 ? (edy >> 1) : ((edy + 1) >> 1);
            // a crossing at or left of L is already counted by the backdrop and the V term (its two H contributions cancel)
            if (!(edy > 0 ? (m <= (T)h) : (m >= (T)h))) {
                T qq = m - (T)edy * rxo;  // first column c with 256*|dy|*c >= qq is the first one at or right of the crossing
                if (edy < 0) qq = -qq;
                int c0 = 0;
                if (qq > 0) {
                    const T D = (T)256 * (edy > 0 ? edy : -edy);
                    int     c = 0;
                    T       Dc = 0;
                    if (D * 8 < qq) { c = 8; Dc = D * 8; }
                    if (Dc + D * 4 < qq) { c += 4; Dc += D * 4; }
                    if (Dc + D * 2 < qq) { c += 2; Dc += D * 2; }
                    if (Dc + D < qq) { c += 1; }
                    c0 = c + 1;  // c = last column still left of the crossing
                }
                if (c0 < 16) {
                    uint32_t *row = dl + ((r & ~(S - 1)) | smp) * FW_DL_STRIDE;  // (storage order of the rows: pixel row * S + sample)
                    const uint32_t one = edy > 0 ? 1u : 0xffffffffu;
                    if (mode == 0) atomicXor(row, (0xFFFFu << c0) & 0xFFFFu);
                    else atomicAdd(row + (c0 >> 1), one << (16 * (c0 & 1)));
                }
            }
        }
    }
}
template <int S, int P>
__device__ __noinline__ void fw_chunk_far(uint32_t lane, bool valid, int32_t ax, int32_t ay, int32_t bx, int32_t by, int32_t ry0, int32_t ry1, int mode,
                                          uint32_t *dl, int32_t *base0, int32_t *base1) {
    int32_t ry[P], base[P];
    ry[0] = ry0; base[0] = *base0;
    if (P > 1) { ry[P - 1] = ry1; base[P - 1] = *base1; }
    fw_chunk<S, P, long long>(lane, valid, ax, ay, bx, by, ry, mode, dl, base);
    *base0 = base[0];
    if (P > 1) *base1 = base[P - 1];
}

// A path-tile with more than FW_CH edges: (edge, sample row) work items into the packed deltas of the rows, then the coverage bits of this
// lane's rows (bit 16 * j + c).  Out of line: the hot loop of fine_warp_k (short lists) keeps its size in the instruction cache.
// *cur holds edge `lane` of the list on entry and the first edges of the NEXT path-tile (eoff_n, ne_n) on return.
template <int S, int P>
__device__ __noinline__ uint32_t fw_long_list(uint32_t lane, const vkb_edge *tile_edges, uint32_t eoff, int n_e, uint32_t eoff_n, int ne_n, int4 *curp,
                                              int32_t X0, int32_t Y0, int32_t ry0, int32_t ry1, int32_t bd, int par, bool counted, uint32_t *dl, int32_t *Wg) {
    constexpr uint32_t FULL = 0xffffffffu;
    constexpr int      ROWS = 16 * S;
    int32_t ry[P];
    int     row[P];
    ry[0] = ry0; row[0] = (int)lane;
    if (P > 1) { ry[P - 1] = ry1; row[P - 1] = 32 + (int)lane; }
    int4       cur   = *curp;
    const int  mode  = par == 1 ? 0 : 2;  // 0: parity words, 2: packed biased int16
    const bool use_w = counted;           // the blend needs |winding| per sample: kept in this warp's int32 plane
    const bool multi = mode == 2 && n_e > 32 * FW_SEG_CHUNKS;  // the int16 deltas are folded into the int32 plane every FW_SEG_CHUNKS chunks
    const uint32_t bias = mode == 0 ? 0u : FW_HBIAS * 0x00010001u;
    const int      nw   = mode == 0 ? 1 : 8;
#pragma unroll
    for (int j = 0; j < P; j++)
        if (row[j] < ROWS)
            for (int w = 0; w < nw; w++) dl[row[j] * FW_DL_STRIDE + w] = bias;
    __syncwarp();
    int32_t base[P];
#pragma unroll
    for (int j = 0; j < P; j++) base[j] = 0;
    bool first_fold = true;
    for (int e0 = 0;; e0 += 32) {
        const int nn = max(0, min(32, n_e - e0));
        int4      nx = make_int4(0, 0, 0, 0);  // prefetch: the next 32 edges of this path-tile, else the first ones of the next
        if (e0 + 32 < n_e) {
            if ((int)lane < n_e - e0 - 32) nx = __ldg((const int4 *)(tile_edges + eoff + e0 + 32 + lane));
        } else if ((int)lane < ne_n) nx = __ldg((const int4 *)(tile_edges + eoff_n + lane));
        const bool    valid = (int)lane < nn;
        const int32_t ax = cur.x - X0, ay = cur.y - Y0, bx = cur.z - X0, by = cur.w - Y0;
        const bool    near = (uint32_t)(ax + 16384) < 36864u && (uint32_t)(ay + 16384) < 36864u && (uint32_t)(bx + 16384) < 36864u &&
                          (uint32_t)(by + 16384) < 36864u;
        if (__all_sync(FULL, near || !valid)) fw_chunk<S, P, int32_t>(lane, valid, ax, ay, bx, by, ry, mode, dl, base);
        else fw_chunk_far<S, P>(lane, valid, ax, ay, bx, by, ry[0], ry[P - 1], mode, dl, &base[0], &base[P - 1]);
        cur = nx;
        const bool last = e0 + 32 >= n_e;
        if (multi && (last || (e0 >> 5) % FW_SEG_CHUNKS == FW_SEG_CHUNKS - 1)) {  // fold the deltas so far into the int32 plane
            __syncwarp();
#pragma unroll
            for (int j = 0; j < P; j++) {
                if (row[j] < ROWS) {
                    uint32_t *dr  = dl + row[j] * FW_DL_STRIDE;
                    int32_t  *wr  = Wg + row[j] * FW_WSTRIDE;
                    for (int c = 0; c < 16; c++) {  // (per-column deltas; the prefix sums along the row are taken once, at the end)
                        const int32_t dlt = (int32_t)((dr[c >> 1] >> (16 * (c & 1))) & 0xFFFFu) - (int32_t)FW_HBIAS;
                        if (first_fold) wr[c] = dlt; else wr[c] += dlt;
                    }
                    for (int w = 0; w < 8; w++) dr[w] = bias;
                }
            }
            first_fold = false;
            __syncwarp();
        }
        if (last) break;
    }
    __syncwarp();
    *curp = cur;
    // ---- coverage bits of this lane's rows ----
    uint32_t m = 0;
#pragma unroll
    for (int j = 0; j < P; j++) {
        if (row[j] < ROWS) {
            const uint32_t *dr  = dl + row[j] * FW_DL_STRIDE;
            const int32_t   off = bd + base[j];
            uint32_t        bits;
            if (mode == 0) bits = (dr[0] ^ ((off & 1) ? 0xFFFFu : 0u)) & 0xFFFFu;
            else {
                int32_t *wr  = Wg + row[j] * FW_WSTRIDE;
                int32_t  run = off;
                bits = 0;
                for (int c = 0; c < 16; c++) {
                    int32_t w;
                    if (multi) run += wr[c];
                    else run += (int32_t)((dr[c >> 1] >> (16 * (c & 1))) & 0xFFFFu) - (int32_t)FW_HBIAS;
                    w = run;
                    if (use_w) wr[c] = w;
                    bits |= (w != 0 ? 1u : 0u) << c;
                }
            }
            m |= bits << (16 * j);
        }
    }
    return m;
}

#define FW_TILES 4      // warps per block, each working on a tile of its own; they only share the u8 -> float table
#define FW_BLOCKS_PER_SM 8
#define FW_WPLANE(S) (16 * (S) * FW_WSTRIDE)  // int32 words of one warp's winding plane (global scratch, see FineArgs::wscratch)
template <int S> struct alignas(16) FwShared {
    uint32_t col[16 * S * 16];        // sample colours, [sample row = ly * S + s][lx]
    uint32_t dl[16 * S * FW_DL_STRIDE];  // per sample row: packed per-column winding deltas of the current path-tile (see fw_chunk)
    uint16_t rowmask[16 * S];         // per sample row: bit lx = covered by the current path-tile
    uint8_t  queue[256];              // pixels (ly * 16 + lx) with at least one covered sample
    float    qxy[32];                 // pixel centre / surface size: 16 columns, then 16 rows (what the gradient evaluation starts from)
};


template <int S> __global__ void __launch_bounds__(32 * FW_TILES, FW_BLOCKS_PER_SM) fine_warp_k(FineArgs a, uint32_t n_tiles) {
    constexpr int ROWS = 16 * S;
    constexpr int P    = ROWS >= 64 ? 2 : 1;  // sample rows per lane
    constexpr uint32_t FULL = 0xffffffffu;
    __shared__ float       lut[256];
    __shared__ FwShared<S> sh_all[FW_TILES];
    for (uint32_t i = threadIdx.x; i < 256; i += 32 * FW_TILES) lut[i] = (float)i / 255.0f;
    __syncthreads();  // the only block-wide barrier
    if (a.counts->overflow) return;
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    FwShared<S>   &sh = sh_all[warp];
    // per-sample int32 windings of the current path-tile, only for COUNT-rule draws with a translucent source (which blend |winding| times)
    // and for lists so long that |winding| may exceed 255: global scratch of this warp (it stays in L1 / L2), not shared memory that every
    // resident warp would have to pay for
    int32_t *const Wg = a.wscratch + (size_t)(blockIdx.x * FW_TILES + warp) * FW_WPLANE(S);
    // Tiles differ a lot in the length of their lists: every warp takes the next tile from a counter when it has finished one, so
    // that no warp slot of the SM sits idle while the slowest tile of a block is still being walked (the grid is one wave).
    for (;;) {
    uint32_t tile = 0;
    if (lane == 0) tile = atomicAdd(a.tile_counter, 1u);
    tile = __shfl_sync(FULL, tile, 0);
    if (tile >= n_tiles) break;
    tile += a.tile_lo;
    const uint32_t first = a.tile_first[tile], end = a.tile_end[tile];
    if (first == end) continue;  // no draw of this batch touches the tile: the stored pixels stay as they are
    const uint32_t tx = tile % a.sd.tiles_x, ty = tile / a.sd.tiles_x;
    const int32_t  X0 = (int32_t)tx * VKB_TILE_FX, Y0 = (int32_t)ty * VKB_TILE_FX;
    const uint32_t band_y0 = a.sd.band_tiles ? (ty / a.sd.band_tiles) * a.sd.band_tiles * VKB_TILE : 0u;

    // ---- destination colours -> shared memory (pixel i * 32 + lane of the tile-major planes is pixel "thread" of fine_k) ----
    {
        const bool tms = !a.dst_is_clear && a.tile_ms[tile];
#pragma unroll 1
        for (uint32_t i = 0; i < 8; i++) {
            const uint32_t lx = (lane & 7) + 8 * (i & 1), ly = (lane >> 3) + 4 * (i >> 1);
            const uint32_t px = tx * VKB_TILE + lx, py = ty * VKB_TILE + ly;
            const bool     ms = tms && ((a.ms_mask[tile * 8 + i] >> lane) & 1u);
            if (ms) {
#pragma unroll
                for (int s = 0; s < S; s++) sh.col[(ly * S + s) * 16 + lx] = a.ms_image[((size_t)tile * 256 + i * 32 + lane) * S + s];
            } else {
                const uint32_t c = (!a.dst_is_clear && px < a.sd.width && py < a.sd.height) ? a.image[(size_t)py * a.sd.width + px] : 0u;
#pragma unroll
                for (int s = 0; s < S; s++) sh.col[(ly * S + s) * 16 + lx] = c;
            }
        }
    }
    // ---- this lane's sample rows ----
    int32_t ry[P], rxo[P];
    int     row[P];
#pragma unroll
    for (int j = 0; j < P; j++) {
        row[j] = j * 32 + (int)lane;
        const int r = min(row[j], ROWS - 1);
        ry[j]  = (r / S) * 256 + 16 * (int32_t)((SamplePack<S>::Y >> (4 * (r % S))) & 15);
        rxo[j] = 16 * (int32_t)((SamplePack<S>::X >> (4 * (r % S))) & 15);
        asm volatile("" : "+r"(ry[j]), "+r"(rxo[j]));  // keep them: the compiler otherwise re-derives both for every path-tile
    }
    sh.qxy[lane] = lane < 16 ? ((float)(tx * VKB_TILE + lane) + 0.5f) / (float)a.sd.width
                             : ((float)(ty * VKB_TILE + (lane - 16) + a.sd.origin_y - band_y0) + 0.5f) / (float)a.sd.full_height;
    const uint32_t grp = lane / S, sub = lane % S;  // pixel row group / which of its pixels this lane queues
    const uint32_t share = (S == 4 ? 0x11111111u : (S == 2 ? 0x55555555u : 0xffffffffu)) << sub;
    __syncwarp();

    for (uint32_t pb = first; pb < end; pb += 32) {
        // headers of the next 32 path-tiles, one per lane
        int4 hh0 = make_int4(0, 0, 0, 0), hh1 = hh0;
        if (pb + lane < end) { hh0 = __ldg(a.hdr + 2 * (size_t)(pb + lane)); hh1 = __ldg(a.hdr + 2 * (size_t)(pb + lane) + 1); }
        const int np = (int)min(32u, end - pb);
        int4      cur = make_int4(0, 0, 0, 0);  // edge (chunk start + lane) of the chunk about to be processed
        {
            const uint32_t eo = (uint32_t)__shfl_sync(FULL, hh0.z, 0);
            const int      ne = __shfl_sync(FULL, hh0.w, 0);
            if ((int)lane < ne) cur = __ldg((const int4 *)(a.tile_edges + eo + lane));
        }
        for (int q = 0; q < np; q++) {
            const int32_t   bd   = __shfl_sync(FULL, hh0.y, q);
            const uint32_t  eoff = (uint32_t)__shfl_sync(FULL, hh0.z, q);
            const int       n_e  = __shfl_sync(FULL, hh0.w, q);
            const vkb_paint pt   = vkb_paint{(uint32_t)__shfl_sync(FULL, hh1.x, q), (uint32_t)__shfl_sync(FULL, hh1.y, q),
                                             __int_as_float(__shfl_sync(FULL, hh1.z, q)), (uint32_t)__shfl_sync(FULL, hh1.w, q)};
            uint32_t        eoff_n = 0;
            int             ne_n   = 0;
            if (q + 1 < np) { eoff_n = (uint32_t)__shfl_sync(FULL, hh0.z, q + 1); ne_n = __shfl_sync(FULL, hh0.w, q + 1); }
            const uint32_t rule = pt.rule_pattern & 0xFF, pattern = (pt.rule_pattern >> 8) & 0xFF, op = (pt.rule_pattern >> 16) & 0xFF;
            // solid sources: one colour for the whole path-tile
            float    src[4], ia = 0.0f;
            bool     use_int = false;
            uint32_t s_lo = 0, s_hi = 0, IA = 0;
            if (pattern == VKB_PAT_SOLID) {
                eval_paint(pattern, nullptr, nullptr, 0.0f, 0.0f, pt.color, pt.opacity, 0.0f, 0.0f, src, lut);
                ia = 1.0f - src[3];
                if (op == VKB_OP_SUB) ia = -ia;
                else if (op == VKB_OP_CLEAR) { src[0] = src[1] = src[2] = src[3] = 0.0f; ia = 0.0f; }
                // exact multiples of 1/255 with no channel above alpha: the blend is done on packed integers (blend_int)
                const uint32_t A = pt.color >> 24;
                if (op == VKB_OP_CLEAR) use_int = true;  // (S = 0, IA = 0)
                else if (op == VKB_OP_OVER && pt.opacity == 1.0f && (pt.color & 0xFF) <= A && ((pt.color >> 8) & 0xFF) <= A && ((pt.color >> 16) & 0xFF) <= A) {
                    use_int = true;
                    s_lo = __byte_perm(pt.color, 0, 0x4240); s_hi = __byte_perm(pt.color, 0, 0x4341); IA = 255u - A;
                }
            }
            // COUNT rule (stroke triangles blended one by one): the number of blends matters unless the result does not depend on the destination
            const bool counted = rule == VKB_RULE_COUNT && !(pattern == VKB_PAT_SOLID && ia == 0.0f);
            const int      par   = rule == VKB_RULE_EVEN_ODD ? 1 : -1;  // covered <=> (winding & par) != 0
            uint32_t       m     = 0;                                   // bit j * 16 + c: sample row row[j], column c is covered
            const bool     skip  = rule >= VKB_RULE_CLIP_EO;            // stencil entries never reach this kernel (they bring a stencil plane with them)

            if (n_e <= FW_CH) {
                // ---- short lists: winding in bit-sliced counters, one lane per sample row (plane_add) ----
                const uint32_t wmax = (bd < 0 ? 0u - (uint32_t)bd : (uint32_t)bd) + (uint32_t)n_e;  // |winding| <= |backdrop| + one per edge
                const bool     use_w = counted || (rule != VKB_RULE_EVEN_ODD && wmax > 255u);
                // use_w (COUNT rule with a translucent source, or |winding| possibly above 255): the counters are folded into this warp's
                // int32 plane every 96 edges (partial sums stay within the signed 8 bits), which then holds the exact windings
                const int mode = par == 1 ? 0 : ((use_w || wmax > 15u) ? 2 : 1);
                uint32_t  pl[8];
#pragma unroll
                for (int p = 0; p < 8; p++) pl[p] = (!use_w && ((bd >> p) & 1)) ? FULL : 0u;
                bool first_fold = true;
                for (int e0 = 0;; e0 += 32) {
                    const int nn = max(0, min(32, n_e - e0));
                    int4      nx = make_int4(0, 0, 0, 0);  // prefetch: the next 32 edges of this path-tile, else the first ones of the next
                    if (e0 + 32 < n_e) {
                        if ((int)lane < n_e - e0 - 32) nx = __ldg((const int4 *)(a.tile_edges + eoff + e0 + 32 + lane));
                    } else if ((int)lane < ne_n) nx = __ldg((const int4 *)(a.tile_edges + eoff_n + lane));
                    // every lane classifies its own edge once; the loop below reads one bit per edge
                    const uint32_t nearmask = __ballot_sync(FULL, (uint32_t)(cur.x - X0 + 16384) < 36864u && (uint32_t)(cur.y - Y0 + 16384) < 36864u &&
                                                                      (uint32_t)(cur.z - X0 + 16384) < 36864u && (uint32_t)(cur.w - Y0 + 16384) < 36864u);
                    for (int k = 0; k < nn; k++) {
                        const int32_t ax = __shfl_sync(FULL, cur.x, k) - X0, ay = __shfl_sync(FULL, cur.y, k) - Y0;
                        const int32_t bx = __shfl_sync(FULL, cur.z, k) - X0, by = __shfl_sync(FULL, cur.w, k) - Y0;
                        const bool    crossL = (ax <= 0) != (bx <= 0);
                        // (binning is conservative: an edge that spans none of the tile's sample rows and does not cross L changes nothing)
                        const int32_t ylo = min(ay, by), yhi = max(ay, by);
                        if (!crossL && (((ylo - FwRows<S>::OFF + ((1 << FwRows<S>::LOG) - 1)) >> FwRows<S>::LOG) >= ((yhi - FwRows<S>::OFF + ((1 << FwRows<S>::LOG) - 1)) >> FwRows<S>::LOG) ||
                                        yhi <= FwRows<S>::OFF || ylo > ((ROWS - 1) << FwRows<S>::LOG) + FwRows<S>::OFF))
                            continue;
                        const unsigned long long mk = ((nearmask >> k) & 1u) ? row_edge_masks<P, int32_t>(ax, ay, bx, by, crossL, ry[0], ry[P - 1], rxo[0], rxo[P - 1])
                                                                               : row_edge_masks_far<P>(ax, ay, bx, by, crossL, ry[0], ry[P - 1], rxo[0], rxo[P - 1]);
                        const uint32_t hm = (uint32_t)mk, vb = (uint32_t)(mk >> 32);
                        plane_add(pl, hm, by < ay, mode);
                        if (crossL) plane_add(pl, ((vb & 1u) ? 0xFFFFu : 0u) | ((vb & 2u) ? 0xFFFF0000u : 0u), (vb & 4u) != 0, mode);
                    }
                    cur = nx;
                    const bool last = e0 + 32 >= n_e;
                    if (use_w && (last || (e0 >> 5) % 3 == 2)) {
#pragma unroll
                        for (int j = 0; j < P; j++) {
                            if (row[j] < ROWS) {
                                int32_t *wr = Wg + row[j] * FW_WSTRIDE;
#pragma unroll 4
                                for (int c = 0; c < 16; c++) {
                                    uint32_t v = 0;
#pragma unroll
                                    for (int p = 0; p < 8; p++) v |= ((pl[p] >> (16 * j + c)) & 1u) << p;
                                    const int32_t sv = (int32_t)(int8_t)v;
                                    if (first_fold) wr[c] = bd + sv; else wr[c] += sv;
                                }
                            }
                        }
                        first_fold = false;
#pragma unroll
                        for (int p = 0; p < 8; p++) pl[p] = 0u;
                    }
                    if (last) break;
                }
                if (!use_w) {
                    m = pl[0];
                    if (mode >= 1) m |= pl[1] | pl[2] | pl[3];
                    if (mode == 2) m |= pl[4] | pl[5] | pl[6] | pl[7];
                    m &= (row[0] < ROWS ? 0xFFFFu : 0u) | ((P > 1 && row[P - 1] < ROWS) ? 0xFFFF0000u : 0u);
                } else {  // each lane reads back the rows it accumulated itself
#pragma unroll
                    for (int j = 0; j < P; j++) {
                        if (row[j] < ROWS) {
                            const int32_t *wr   = Wg + row[j] * FW_WSTRIDE;
                            uint32_t       bits = 0;
#pragma unroll
                            for (int c = 0; c < 16; c++) bits |= ((wr[c] & par) != 0 ? 1u : 0u) << c;
                            m |= bits << (16 * j);
                        }
                    }
                }
            } else {  // long lists: (edge, sample row) work items, out of line (fw_long_list)
                m = fw_long_list<S, P>(lane, a.tile_edges, eoff, n_e, eoff_n, ne_n, &cur, X0, Y0, ry[0], ry[P - 1], bd, par, counted, sh.dl, Wg);
            }
            if (skip) continue;
#pragma unroll
            for (int j = 0; j < P; j++)
                if (row[j] < ROWS) sh.rowmask[row[j]] = (uint16_t)(m >> (16 * j));
            // ---- queue of the pixels with a covered sample: OR over the S sample rows of a pixel row, then every lane of the
            //      group pushes its share of the set bits ----
            uint32_t many = m;
            if (S >= 2) many |= __shfl_xor_sync(FULL, many, 1);
            if (S >= 4) many |= __shfl_xor_sync(FULL, many, 2);
            uint32_t       mine  = many & share;
            const uint32_t cnt   = (uint32_t)__popc(mine);
            const uint32_t incl  = warp_incl_scan(cnt);
            const uint32_t total = __shfl_sync(FULL, incl, 31);
            if (total == 0) continue;  // (warp uniform; nothing was written that a later path-tile does not overwrite)
            uint32_t off = incl - cnt;
            while (mine) {
                const uint32_t b = (uint32_t)__ffs((int)mine) - 1u;
                mine &= mine - 1u;
                sh.queue[off++] = (uint8_t)((grp + 8 * (b >> 4)) * 16 + (b & 15));
            }
            __syncwarp();  // (also orders the winding plane in global memory between the lanes of the warp)
            // ---- blend, 32 pixels at a time ----
            for (uint32_t i0 = 0; i0 < total; i0 += 32) {
                const bool     act = i0 + lane < total;
                const uint32_t pq  = act ? sh.queue[i0 + lane] : 0u;
                const uint32_t lx = pq & 15, ly = pq >> 4;
                int32_t        n[S], nmax = 0;
                if (counted) {
#pragma unroll
                    for (int s = 0; s < S; s++) n[s] = act ? abs(Wg[(ly * S + s) * FW_WSTRIDE + lx]) : 0;
                } else {  // the S row masks of a pixel row are adjacent: one load
                    unsigned long long rm;
                    if (S == 4) rm = *(const unsigned long long *)(sh.rowmask + ly * 4);
                    else if (S == 2) rm = *(const uint32_t *)(sh.rowmask + ly * 2);
                    else rm = sh.rowmask[ly];
                    rm = act ? rm >> lx : 0ull;
#pragma unroll
                    for (int s = 0; s < S; s++) n[s] = (int32_t)((uint32_t)(rm >> (16 * s)) & 1u);
                }
                if (pattern != VKB_PAT_SOLID) {
                    const uint32_t px = tx * VKB_TILE + lx, py = ty * VKB_TILE + ly;
                    eval_paint_q(pattern, a.grads + pt.gradient, a.gprep + (size_t)pt.gradient * VKB_GPREP_FLOATS, pt.color, pt.opacity, (float)px + 0.5f,
                                 (float)(py + a.sd.origin_y - band_y0) + 0.5f, sh.qxy[lx], sh.qxy[16 + ly], src, lut, a.surfpats, pt.gradient);
                    ia = 1.0f - src[3];
                    if (op == VKB_OP_SUB) ia = -ia;
                    else if (op == VKB_OP_CLEAR) { src[0] = src[1] = src[2] = src[3] = 0.0f; ia = 0.0f; }
                }
                const bool opaque = ia == 0.0f;  // the result does not depend on the destination: repeating the blend changes nothing
                uint32_t  *cp = sh.col + (ly * S) * 16 + lx;
                uint32_t   c[S];
                bool       uni = true, two = true;
#pragma unroll
                for (int s = 0; s < S; s++) {
                    if (counted && opaque) n[s] = n[s] ? 1 : 0;
                    nmax = max(nmax, n[s]);
                    c[s] = cp[16 * s];
                    uni  = uni && c[s] == c[0];
                }
                if (counted) {
#pragma unroll
                    for (int s = 0; s < S; s++) two = two && (n[s] == 0 || n[s] == nmax);
                }
                // one result per pixel when its samples hold one colour (or the source is opaque) and are blended equally often:
                // warp-uniform choice, a divergent one would execute both variants
                const bool one = __all_sync(FULL, !act || ((uni || opaque) && two));
                if (use_int) {  // (warp uniform) packed-integer blend: ten instructions per sample
                    if (one) {
                        uint32_t r = blend_int(c[0], s_lo, s_hi, IA);
                        if (counted)
                            for (int32_t k = 1; k < nmax; k++) r = blend_int(r, s_lo, s_hi, IA);
                        if (act) {
#pragma unroll
                            for (int s = 0; s < S; s++)
                                if (n[s]) cp[16 * s] = r;
                        }
                    } else {
#pragma unroll
                        for (int s = 0; s < S; s++) {
                            if (n[s] != 0) {
                                uint32_t r = blend_int(c[s], s_lo, s_hi, IA);
                                if (counted)
                                    for (int32_t k = 1; k < n[s]; k++) r = blend_int(r, s_lo, s_hi, IA);
                                cp[16 * s] = r;
                            }
                        }
                    }
                } else if (one) {
                    uint32_t r = blend_over(c[0], src, ia, lut);
                    if (counted)
                        for (int32_t k = 1; k < nmax; k++) r = blend_over(r, src, ia, lut);
                    if (act) {
#pragma unroll
                        for (int s = 0; s < S; s++)
                            if (n[s]) cp[16 * s] = r;
                    }
                } else {  // rolled: one copy of the fp32 blend in the instruction cache, which the gradient code fills
                    uint32_t nb = 0;
#pragma unroll
                    for (int s = 0; s < S; s++) nb |= (n[s] ? 1u : 0u) << s;
#pragma unroll 1
                    for (int s = 0; s < S; s++) {
                        if ((nb >> s) & 1u) {  // (a sample no lane covers costs one branch)
                            uint32_t r = blend_over(cp[16 * s], src, ia, lut);
                            if (counted) {
                                const int32_t reps = abs(Wg[(ly * S + s) * FW_WSTRIDE + lx]);
                                for (int32_t k = 1; k < (opaque ? 1 : reps); k++) r = blend_over(r, src, ia, lut);
                            }
                            cp[16 * s] = r;
                        }
                    }
                }
            }
            __syncwarp();  // the queue, the row masks and the winding plane are reused by the next path-tile
        }
    }

    // ---- resolve; pixels whose samples differ also go to the per-sample plane (same layout and flags as fine_k) ----
    __syncwarp();
    uint32_t my_mask = 0, any_mask = 0;
#pragma unroll 1
    for (uint32_t i = 0; i < 8; i++) {
        const uint32_t lx = (lane & 7) + 8 * (i & 1), ly = (lane >> 3) + 4 * (i >> 1);
        const uint32_t px = tx * VKB_TILE + lx, py = ty * VKB_TILE + ly;
        uint32_t       c[S];
        bool           differ = false;
#pragma unroll
        for (int s = 0; s < S; s++) {
            c[s]   = sh.col[(ly * S + s) * 16 + lx];
            differ = differ || c[s] != c[0];
        }
        const uint32_t dmask = __ballot_sync(FULL, differ);
        if (lane == i) my_mask = dmask;
        any_mask |= dmask;
        if (differ) {
#pragma unroll
            for (int s = 0; s < S; s++) a.ms_image[((size_t)tile * 256 + i * 32 + lane) * S + s] = c[s];
        }
        if (px < a.sd.width && py < a.sd.height) {
            uint32_t out = 0;
#pragma unroll
            for (int k = 0; k < 4; k++) {
                uint32_t sum = 0;
#pragma unroll
                for (int s = 0; s < S; s++) sum += (c[s] >> (8 * k)) & 0xFF;
                out |= ((sum + S / 2) / S) << (8 * k);
            }
            a.image[(size_t)py * a.sd.width + px] = out;
        }
    }
    if (any_mask && lane < 8) a.ms_mask[tile * 8 + lane] = my_mask;
    if (lane == 0) a.tile_ms[tile] = any_mask ? 1 : 0;
    __syncwarp();  // the next tile's colours overwrite this one's
    }
}

// Which kernel serves a batch without clip state: the warp-per-tile kernel needs many more tiles than the GPU has warp slots
// (148 SMs x 24 warps) to keep every scheduler busy, and a tile with a long list is walked by one warp from start to end; on
// small surfaces (tiger at 1024^2: 4096 tiles) the block-per-tile kernel finishes sooner.  vkvg_b200_set_fine_kernel /
// VKVG_B200_FINE=block|warp override the choice (A/B timing, kernel-equivalence test).
#define FW_MIN_TILES 16384u
static int g_fine_mode = [] {
    const char *e = getenv("VKVG_B200_FINE");
    return (e && e[0] == 'b') ? 1 : ((e && e[0] == 'w') ? 2 : 0);
}();
void vkb_fine_set_mode(int mode) { g_fine_mode = (mode == 1 || mode == 2) ? mode : 0; }
int  vkb_fine_get_mode() { return g_fine_mode; }
// words of winding-plane scratch the warp-per-tile kernel needs (FineArgs::wscratch): one plane per warp of the (one-wave) grid
size_t vkb_fine_wscratch_words(uint32_t samples) { return (size_t)148 * FW_BLOCKS_PER_SM * FW_TILES * FW_WPLANE(samples > 4 ? 4 : (samples ? samples : 1)); }
template <int S> static void launch_fine_warp(const FineArgs &a, uint32_t tiles, cudaStream_t s) {
    // one wave of persistent warps: as many blocks as are resident at once (registers allow FW_BLOCKS_PER_SM, shared memory may allow fewer)
    static int per_sm = 0;
    if (!per_sm) {
        int n = 0;
        if (const char *e = getenv("VKVG_B200_FW_CARVEOUT")) cudaFuncSetAttribute(fine_warp_k<S>, cudaFuncAttributePreferredSharedMemoryCarveout, atoi(e));  // (experiments)
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, fine_warp_k<S>, 32 * FW_TILES, 0) != cudaSuccess || n < 1) { cudaGetLastError(); n = 4; }
        per_sm = n < FW_BLOCKS_PER_SM ? n : FW_BLOCKS_PER_SM;
        if (const char *e = getenv("VKVG_B200_FW_BLOCKS")) per_sm = atoi(e) > 0 && atoi(e) < per_sm ? atoi(e) : per_sm;
    }
    const uint32_t blocks = vkb_div_up(tiles, FW_TILES), wave = 148u * (uint32_t)per_sm;
    fine_warp_k<S><<<blocks < wave ? blocks : wave, 32 * FW_TILES, 0, s>>>(a, tiles);
}
template <int S> static void launch_fine_s(const FineArgs &a, uint32_t tiles, cudaStream_t s) {
    const bool cap = a.winding_out != nullptr, clip = a.stencil != nullptr;
    if constexpr (S <= 4) {
        if (!cap && !clip && (g_fine_mode == 2 || (g_fine_mode == 0 && tiles >= FW_MIN_TILES))) { launch_fine_warp<S>(a, tiles, s); return; }
    }
    if (cap) { if (clip) fine_k<S, true, true><<<tiles, 256, 0, s>>>(a); else fine_k<S, true, false><<<tiles, 256, 0, s>>>(a); }
    else { if (clip) fine_k<S, false, true><<<tiles, 256, 0, s>>>(a); else fine_k<S, false, false><<<tiles, 256, 0, s>>>(a); }
}
void vkb_launch_fine(const FineArgs &a, cudaStream_t s) {
    uint32_t tiles = a.tile_hi - a.tile_lo;   // (the whole surface, or one band of tile rows: FineArgs::tile_lo / tile_hi)
    if (!tiles) return;
    const bool cap = a.winding_out != nullptr, clip = a.stencil != nullptr;
    switch (a.sd.samples) {
    case 0:
        if (cap) { if (clip) fine_analytic_k<true, true><<<tiles, 256, 0, s>>>(a); else fine_analytic_k<true, false><<<tiles, 256, 0, s>>>(a); }
        else { if (clip) fine_analytic_k<false, true><<<tiles, 256, 0, s>>>(a); else fine_analytic_k<false, false><<<tiles, 256, 0, s>>>(a); }
        break;
    case 1: launch_fine_s<1>(a, tiles, s); break;
    case 2: launch_fine_s<2>(a, tiles, s); break;
    case 4: launch_fine_s<4>(a, tiles, s); break;
    case 8: launch_fine_s<8>(a, tiles, s); break;
    case 16: launch_fine_s<16>(a, tiles, s); break;
    default: return;
    }
    VKB_LAUNCHED();
}

// vkvg_surface_write_to_png / _to_memory un-premultiply in double with truncation (src/vkvg_surface.c:371-382)
__global__ void unpremultiply_k(const uint32_t *image, uint64_t n, uint32_t *out) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t p = image[i], a = p >> 24, r = 0;
    double   alpha = (double)a / 255.f;
#pragma unroll
    for (int k = 0; k < 3; k++) {
        double v = (double)((p >> (8 * k)) & 0xFF) / alpha;
        uint32_t q = (v == v && v < 2147483648.0) ? ((uint32_t)(int32_t)v & 0xFF) : 0u;
        r |= q << (8 * k);
    }
    out[i] = r | (a << 24);
}
void vkb_launch_unpremultiply(const uint32_t *image, uint64_t n_pixels, uint32_t *out, cudaStream_t s) {
    if (!n_pixels) return;
    unpremultiply_k<<<vkb_div_up(n_pixels, 256), 256, 0, s>>>(image, n_pixels, out);
    VKB_LAUNCHED();
}
