// Stroke expansion kernels: flattened points -> stroke triangles (vertices + indices).
//
// Replaces the serial loop of the reference (_stroke_preserve src/vkvg_context.c:822-948, _build_vb_step
// src/vkvg_context_internal.c:924-1163, _draw_stoke_cap :1165-1239, _draw_dashed_segment :1240-1264).
// The reference's output order depends on a running vertex counter and, for dashes, on a carried dash
// phase.  Here every point of every stroked sub-path is one work item: a count pass reproduces every
// branch (including the float `while` loops that size round joins / caps), a prefix scan turns counts into
// vertex / index offsets, and an emit pass writes the same vertices in the same order.  Because vertex
// numbering is contiguous, the reference's forward references ("closing quad of a join uses the next
// join's first two vertices") are simply base + own_count (+1).  The dash phase comes from a
// double-precision prefix scan of the float segment lengths instead of a carried float.
// The emit pass (stroke_emit_k) also runs the vertex stage - what the rasteriser reads are the snapped integer vertices - and writes a
// block's contiguous output ranges through shared memory with 16-byte stores; the float vertices exist only for geometry captures.
#include "pipeline.h"
#include <float.h>
#include <string.h>

#define PIF 3.14159265358979323846f
#define PIF_2 1.57079632679489661923f
#define EQUF(a, b) (fabsf((a) - (b)) <= FLT_EPSILON)

struct v2 {
    float x, y;
};
__device__ __forceinline__ v2    mk(float x, float y) { v2 r = {x, y}; return r; }
__device__ __forceinline__ v2    v2add(v2 a, v2 b) { return mk(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ v2    v2sub(v2 a, v2 b) { return mk(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ v2    v2mul(v2 a, float m) { return mk(a.x * m, a.y * m); }
__device__ __forceinline__ v2    v2div(v2 a, float m) { return mk(a.x / m, a.y / m); }
__device__ __forceinline__ v2    v2perp(v2 a) { return mk(a.y, -a.x); }
__device__ __forceinline__ float v2len(v2 a) { return sqrtf(a.x * a.x + a.y * a.y); }
__device__ __forceinline__ v2    v2norm(v2 a) { float m = sqrtf(a.x * a.x + a.y * a.y); return mk(a.x / m, a.y / m); }
__device__ __forceinline__ v2    ld(const float2 *p, uint32_t i) { float2 f = p[i]; return mk(f.x, f.y); }

// MODE 0: count only.  MODE 1: float vertices + indices straight to global memory (the first emitter, kept as VKVG_B200_STROKE=legacy).
// MODE 2: every vertex also goes through the vertex stage here (vs_snap: what the rasteriser reads are the snapped integers) and vertices and
// indices are written wherever sn / idx point - global memory, or the block's staging area in shared memory, which holds the contiguous
// output range of the block's 128 items from the 16-byte aligned global positions voff / ioff on (stroke_emit_k).
template <int MODE> struct Out {
    float2  *v;
    uint32_t *idx;
    uint32_t vbase, ibase;  // absolute offsets of this item
    uint32_t nv, ni;
    int2    *sn;
    uint32_t voff, ioff;
    float    m[6], W, H;
    int32_t  yoff;
    __device__ __forceinline__ uint32_t cur() const { return vbase + nv; }  // == reference's (vertCount - curVertOffset)
    __device__ __forceinline__ void     vert(v2 p) {
        if (MODE == 1) v[vbase + nv] = make_float2(p.x, p.y);
        if (MODE == 2) {
            if (v) v[vbase + nv] = make_float2(p.x, p.y);  // (geometry captures)
            int32_t x, y;
            vs_snap(m, W, H, p.x, p.y, x, y);
            sn[vbase + nv - voff] = make_int2(x, y + yoff);
        }
        nv++;
    }
    __device__ __forceinline__ void tri(uint32_t a, uint32_t b, uint32_t c) {
        if (MODE == 1) {
            idx[ibase + ni]     = a;
            idx[ibase + ni + 1] = b;
            idx[ibase + ni + 2] = c;
        }
        if (MODE == 2) {
            uint32_t *t = idx + (ibase + ni - ioff);
            t[0] = a; t[1] = b; t[2] = c;
        }
        ni += 3;
    }
    __device__ __forceinline__ void rect(uint32_t i) {  // _add_tri_indices_for_rect, internal.c:341-354
        tri(i, i + 2, i + 1);
        tri(i + 1, i + 2, i + 3);
    }
};

struct StrokeParams {
    float    hw, lhMax, arcStep;
    uint32_t join, cap;
};

// one join, internal.c:924-1163.  Returns the reference's `inverse` flag.
template <int MODE> __device__ bool build_join(Out<MODE> &o, const StrokeParams &sp, v2 pL, v2 p0, v2 pR, bool isCurve) {
    v2    v0 = v2sub(p0, pL), v1 = v2sub(pR, p0);
    float length_v0 = v2len(v0), length_v1 = v2len(v1);
    if (length_v0 < FLT_EPSILON || length_v1 < FLT_EPSILON) return false;
    v2    v0n = v2div(v0, length_v0), v1n = v2div(v1, length_v1);
    float dot = (v0n.x * v1n.x) + (v0n.y * v1n.y);
    float det = v0n.x * v1n.y - v0n.y * v1n.x;
    if (EQUF(dot, 1.0f)) return false;
    uint32_t idx = o.cur();
    if (EQUF(dot, -1.0f)) {  // cusp
        v2 vPerp = v2mul(v2perp(v0n), sp.hw);
        o.vert(v2add(p0, vPerp));
        o.vert(v2sub(p0, vPerp));
        o.tri(idx, idx + 1, idx + 2);
        o.tri(idx, idx + 2, idx + 3);
        return true;
    }
    v2    bisec_n = v2norm(v2add(v0n, v1n));
    float alpha   = acosf(dot);
    if (det < 0) alpha = -alpha;
    float halfAlpha    = alpha / 2.f;
    float cosHalfAlpha = cosf(halfAlpha);
    float lh           = sp.hw / cosHalfAlpha;
    v2    bisec_n_perp = v2perp(bisec_n);
    float rlh          = lh;
    if (dot < 0.f) rlh = fminf(rlh, fminf(length_v0, length_v1));
    v2 bisec = v2mul(bisec_n_perp, rlh);
    v2 in_pos, out_pos;
    if (rlh < lh) {
        v2    vnPerp  = length_v0 < length_v1 ? v2perp(v1n) : v2perp(v0n);
        v2    vHwPerp = v2mul(vnPerp, sp.hw);
        float lbc     = cosHalfAlpha * rlh;  // a double temporary in the reference, but computed and consumed as float
        if (det < 0.f) {
            in_pos  = v2add(v2add(v2mul(vnPerp, -lbc), v2add(p0, bisec)), vHwPerp);
            out_pos = v2sub(p0, v2mul(bisec_n_perp, lh));
        } else {
            in_pos  = v2sub(v2add(v2mul(vnPerp, lbc), v2sub(p0, bisec)), vHwPerp);
            out_pos = v2add(p0, v2mul(bisec_n_perp, lh));
        }
    } else {
        if (det < 0.0f) { in_pos = v2add(p0, bisec); out_pos = v2sub(p0, bisec); }
        else { in_pos = v2sub(p0, bisec); out_pos = v2add(p0, bisec); }
    }
    uint32_t join = sp.join;
    if (isCurve) join = dot < 0.8f ? 1u /*ROUND*/ : 0u /*MITER*/;
    if (join == 0) {  // VKVG_LINE_JOIN_MITER
        if (lh > sp.lhMax) {
            float x         = (lh - sp.lhMax) * cosHalfAlpha;
            v2    bisecPerp = v2mul(bisec_n, x);
            bisec           = v2mul(bisec_n_perp, sp.lhMax);
            if (det < 0) {
                o.vert(in_pos);
                v2 p = v2sub(p0, bisec);
                o.vert(v2sub(p, bisecPerp));
                o.vert(v2add(p, bisecPerp));
                o.tri(idx, idx + 2, idx + 1);
                o.tri(idx + 2, idx + 4, idx);
                o.tri(idx, idx + 3, idx + 4);
                return true;
            } else {
                v2 p = v2add(p0, bisec);
                o.vert(v2sub(p, bisecPerp));
                o.vert(in_pos);
                o.vert(v2add(p, bisecPerp));
                o.tri(idx, idx + 2, idx + 1);
                o.tri(idx + 2, idx + 3, idx + 1);
                o.tri(idx + 1, idx + 3, idx + 4);
                return false;
            }
        } else {
            if (det < 0) { o.vert(in_pos); o.vert(out_pos); }
            else { o.vert(out_pos); o.vert(in_pos); }
            o.rect(idx);
            return false;
        }
    } else {
        v2 vp = v2perp(v0n);
        if (det < 0) {
            o.vert((dot < 0 && rlh < lh) ? in_pos : v2add(p0, bisec));
            o.vert(v2sub(p0, v2mul(vp, sp.hw)));
        } else {
            o.vert(v2add(p0, v2mul(vp, sp.hw)));
            o.vert((dot < 0 && rlh < lh) ? in_pos : v2sub(p0, bisec));
        }
        if (join == 2) {  // BEVEL
            if (det < 0) { o.tri(idx, idx + 2, idx + 1); o.tri(idx + 2, idx + 4, idx + 0); o.tri(idx, idx + 3, idx + 4); }
            else { o.tri(idx, idx + 2, idx + 1); o.tri(idx + 2, idx + 3, idx + 1); o.tri(idx + 1, idx + 3, idx + 4); }
        } else if (join == 1) {  // ROUND
            float a = acosf(vp.x);
            if (vp.y < 0) a = -a;
            if (det < 0) {
                a += PIF;
                float a1 = a + alpha;
                a -= sp.arcStep;
                while (a > a1) { o.vert(mk(cosf(a) * sp.hw + p0.x, sinf(a) * sp.hw + p0.y)); a -= sp.arcStep; }
            } else {
                float a1 = a + alpha;
                a += sp.arcStep;
                while (a < a1) { o.vert(mk(cosf(a) * sp.hw + p0.x, sinf(a) * sp.hw + p0.y)); a += sp.arcStep; }
            }
            uint32_t p0Idx = o.cur();
            o.tri(idx, idx + 2, idx + 1);
            if (det < 0) {
                for (uint32_t p = idx + 2; p < p0Idx; p++) o.tri(p, p + 1, idx);
                o.tri(p0Idx, p0Idx + 2, idx);
                o.tri(idx, p0Idx + 1, p0Idx + 2);
            } else {
                for (uint32_t p = idx + 2; p < p0Idx; p++) o.tri(p, p + 1, idx + 1);
                o.tri(p0Idx, p0Idx + 1, idx + 1);
                o.tri(idx + 1, p0Idx + 1, p0Idx + 2);
            }
        }
        vp = v2mul(v2perp(v1n), sp.hw);
        o.vert(det < 0 ? v2sub(p0, vp) : v2add(p0, vp));
    }
    return (det < 0);
}

// caps, internal.c:1165-1239
template <int MODE> __device__ void draw_cap(Out<MODE> &o, const StrokeParams &sp, v2 p0, v2 n, bool isStart) {
    uint32_t firstIdx = o.cur();
    if (isStart) {
        v2 vhw = v2mul(n, sp.hw);
        if (sp.cap == 2) p0 = v2sub(p0, vhw);  // SQUARE
        vhw = v2perp(vhw);
        if (sp.cap == 1) {  // ROUND
            float a = acosf(n.x) + PIF_2;
            if (n.y < 0) a = PIF - a;
            float a1 = a + PIF;
            a += sp.arcStep;
            while (a < a1) { o.vert(mk(cosf(a) * sp.hw + p0.x, sinf(a) * sp.hw + p0.y)); a += sp.arcStep; }
            uint32_t p0Idx = o.cur();
            for (uint32_t p = firstIdx; p < p0Idx; p++) o.tri(p0Idx + 1, p, p + 1);
            firstIdx = p0Idx;
        }
        o.vert(v2add(p0, vhw));
        o.vert(v2sub(p0, vhw));
        o.rect(firstIdx);
    } else {
        v2 vhw = v2mul(n, sp.hw);
        if (sp.cap == 2) p0 = v2add(p0, vhw);
        vhw = v2perp(vhw);
        o.vert(v2add(p0, vhw));
        o.vert(v2sub(p0, vhw));
        firstIdx = o.cur();
        if (sp.cap == 1) {
            float a = acosf(n.x) + PIF_2;
            if (n.y < 0) a = PIF - a;
            float a1 = a - PIF;
            a -= sp.arcStep;
            while (a > a1) { o.vert(mk(cosf(a) * sp.hw + p0.x, sinf(a) * sp.hw + p0.y)); a -= sp.arcStep; }
            uint32_t p0Idx = o.cur() - 1;
            for (uint32_t p = firstIdx - 1; p < p0Idx; p++) o.tri(p + 1, p, firstIdx - 2);
        }
    }
}

// ---- dash bookkeeping in exact (double) arc length --------------------------------------------------
struct DashPat {
    float  d[VKB_MAX_DASHES];
    double pre[VKB_MAX_DASHES + 1];  // pre[j] = d[0]+..+d[j-1]; pre[n] = total
    int    n;
    double off0;  // fmodf(dashOffset, total), vkvg_context.c:864-866
};
// number of dash boundaries strictly before arc length c (boundary m sits at off0 + (m/n)*tot + pre[m%n])
__device__ __forceinline__ long long dash_count_before(const DashPat &dp, double c) {
    double t = c - dp.off0;
    if (!(t > 0.0)) return 0;
    double    tot = dp.pre[dp.n];
    long long P   = (long long)floor(t / tot);
    double    r   = t - (double)P * tot;
    if (r < 0.0) { P--; r += tot; }
    if (r >= tot) { P++; r -= tot; }
    int j = 0;
    while (j < dp.n && dp.pre[j] < r) j++;
    return P * dp.n + j;
}
__device__ __forceinline__ double dash_boundary_pos(const DashPat &dp, long long m) {
    long long P = m / dp.n;
    int       j = (int)(m - P * dp.n);
    return dp.off0 + (double)P * dp.pre[dp.n] + dp.pre[j];
}

// job table entry for strokes
struct StrokeJob {
    uint32_t draw, first_point, n_points, flags;  // flags: VKB_SP_CLOSED
    uint32_t item_base;                           // first work item (== point) of this job in the global item space
};

// what every stroke kernel needs to find an item's job, sub-path and stroke state
struct ItemArgs {
    const float2      *pts;
    const uint8_t     *ptflags;
    const vkb_draw    *draws;
    const vkb_stroke  *strokes;
    const float       *dash_table;
    const uint32_t    *job_draw, *job_sp, *job_base;
    uint32_t           n_jobs;
    const uint32_t    *sp_first, *sp_count;
    const vkb_subpath *sps;
    const double      *cum;
};
static ItemArgs item_args(const StrokeArgs &a) {
    return ItemArgs{a.pts, a.ptflags, a.draws, a.strokes, a.dash_table, a.job_draw, a.job_sp, a.job_base, a.n_jobs, a.sp_first, a.sp_count, a.sps, a.cum};
}
__device__ __forceinline__ uint32_t item_job(const ItemArgs &a, uint32_t item) {  // last j with job_base[j] <= item
    uint32_t lo = 0, hi = a.n_jobs;
    while (hi - lo > 1) {
        uint32_t mid = (lo + hi) >> 1;
        if (a.job_base[mid] <= item) lo = mid; else hi = mid;
    }
    return lo;
}

// one work item = one point of a stroked sub-path (job j): its join or cap, or - dashed - its segment with the dash caps that fall on it.
// The same code counts (MODE 0) and emits, so the offsets the scan made of the counts are exactly where the emitting pass writes.
template <int MODE>
__device__ __forceinline__ void stroke_item(const ItemArgs &a, uint32_t item, uint32_t j, Out<MODE> &o, uint32_t *job_inverse) {
    const uint32_t k = item - a.job_base[j];
    const uint32_t s = a.job_sp[j];
    const uint32_t first = a.sp_first[s], n = a.sp_count[s];
    const bool     closed = a.sps[s].flags & VKB_SP_CLOSED;
    const vkb_stroke &d = a.strokes[a.draws[a.job_draw[j]].xform_stroke >> 16];
    StrokeParams   sp = {d.hw, d.lhMax, d.arcStep, d.join, d.cap};
    const uint8_t *ptflags = a.ptflags;
    const float2  *P = a.pts + first;
    if (n < 2) return;
    if (d.dash_count == 0) {
        if (closed) {
            // join at every point; the one at the last point closes the loop (vkvg_context.c:917-932)
            uint32_t iL = k == 0 ? n - 1 : k - 1, iR = k == n - 1 ? 0 : k + 1;
            bool     inv = build_join(o, sp, ld(P, iL), ld(P, k), ld(P, iR), k == n - 1 ? false : (ptflags[first + k] != 0));
            if (MODE && k == n - 1) job_inverse[j] = inv;
        } else if (k == 0) {
            draw_cap(o, sp, ld(P, 0), v2norm(v2sub(ld(P, 1), ld(P, 0))), true);  // vkvg_context.c:871-873
        } else if (k == n - 1) {
            draw_cap(o, sp, ld(P, k), v2norm(v2sub(ld(P, k), ld(P, k - 1))), false);  // :934-935
        } else {
            build_join(o, sp, ld(P, k - 1), ld(P, k), ld(P, k + 1), ptflags[first + k] != 0);
        }
    } else {
        // dashed: item k owns segment k -> k+1 (the closing segment for k == n-1 of a closed path) and,
        // if it is the last segment, the tail cap (vkvg_context.c:898-916)
        DashPat dp;
        dp.n      = (int)d.dash_count;
        dp.pre[0] = 0.0;
        for (int i = 0; i < dp.n; i++) {
            dp.d[i]       = a.dash_table[d.dash_first + i];
            dp.pre[i + 1] = dp.pre[i] + (double)dp.d[i];
        }
        float totf = 0.f;
        for (int i = 0; i < dp.n; i++) totf += dp.d[i];  // float accumulation as the reference, :858-859
        dp.off0 = (double)fmodf(d.dash_offset, totf);
        const double   c0       = a.cum[a.job_base[j]];
        const bool     has_seg  = (k + 1 < n) || closed;
        const uint32_t last_seg = closed ? n - 1 : n - 2;
        if (has_seg) {
            const double ck = a.cum[item] - c0, ck1 = a.cum[item + 1] - c0;
            long long    m0 = k == 0 ? 0 : dash_count_before(dp, ck);
            long long    m1 = dash_count_before(dp, ck1);
            uint32_t     iL = k == 0 ? n - 1 : k - 1, iR = k == n - 1 ? 0 : k + 1;  // str.iL = lastPathPointIdx, :867
            v2           p = ld(P, k), pR = ld(P, iR);
            if (m0 & 1)  // inside a dash at the segment start: !dashOn, internal.c:1245
                build_join(o, sp, ld(P, iL), p, pR, k == n - 1 ? false : (ptflags[first + k] != 0));
            v2 dvec = v2sub(pR, p);
            v2 nrm  = v2norm(dvec);
            for (long long m = m0; m < m1; m++) {
                float off = (float)(dash_boundary_pos(dp, m) - ck);
                draw_cap(o, sp, v2add(p, v2mul(nrm, off)), nrm, (m & 1) == 0);
            }
            if (k == last_seg && (m1 & 1)) {
                int   cur  = (int)(m1 % dp.n);
                int   prev = cur - 1 < 0 ? dp.n - 1 : cur - 1;  // the reference reads dashes[-1] here for odd counts (UB)
                float curOff = (float)(dash_boundary_pos(dp, m1) - ck1);
                float mlen   = fminf(dp.d[prev] - curOff, dp.d[cur]);
                draw_cap(o, sp, v2sub(pR, v2mul(nrm, mlen)), nrm, false);
            }
        }
    }
}

// counting pass (MODE 0) and the first emitter (MODE 1: one thread writes its item's vertices and indices straight to global memory)
template <int MODE>
__global__ void __launch_bounds__(128)
stroke_items_k(ItemArgs a, const vkb_counts *C, unsigned long long *counts, const unsigned long long *offsets, float2 *verts, uint32_t *inds, uint32_t *job_inverse) {
    if (C->overflow) return;
    const uint32_t n_items = C->n[VKC_SITEMS];
    // grid-stride over the LIVE items: the grid is sized from the capacity of the item space (a guess for the first flush, what earlier
    // flushes needed afterwards) and capped, see VKB_STROKE_GRID
    for (uint32_t item = blockIdx.x * blockDim.x + threadIdx.x; item < n_items; item += gridDim.x * blockDim.x) {
        Out<MODE> o;
        o.v = verts; o.idx = inds; o.nv = 0; o.ni = 0;
        if (MODE) {
            unsigned long long off = offsets[item];
            o.vbase = (uint32_t)(off & 0xffffffffull);
            o.ibase = (uint32_t)(off >> 32);
        } else
            o.vbase = o.ibase = 0;
        stroke_item<MODE>(a, item, item_job(a, item), o, job_inverse);
        if (!MODE) counts[item] = (unsigned long long)o.nv | ((unsigned long long)o.ni << 32);
    }
}

// The emitter.  A block takes 128 consecutive items; the scan made their outputs one contiguous range of the vertex array and one of the
// index array, so the block builds both ranges in shared memory - every vertex already through the vertex stage (the float vertices are
// only stored for geometry captures) - and then copies them out with 16-byte stores, a warp writing 512 contiguous bytes at a time,
// instead of 128 threads each walking its own few vertices.  Slot 0 of a staging array corresponds to the 16-byte aligned global
// element at or below the range's first one, so shared and global addresses are aligned alike.  A block whose ranges do not fit (round
// joins of a wide stroke: a hundred vertices per item) writes straight to global memory, as does every block when direct != 0.
#define SE_BLOCK 128
#define SE_VERTS 1536  // snapped vertices (int2) a block can stage: 12 KB
#define SE_INDS  4608  // indices: 18 KB
__global__ void __launch_bounds__(SE_BLOCK)
stroke_emit_k(ItemArgs a, const vkb_counts *C, const unsigned long long *offsets, const vkb_xform *xforms, SurfaceDesc sd, int direct, float2 *verts, int2 *snapped,
              uint32_t *inds, uint32_t *job_inverse) {
    if (C->overflow) return;
    __shared__ __align__(16) int2     s_v[SE_VERTS];
    __shared__ __align__(16) uint32_t s_i[SE_INDS];
    const uint32_t n_items = C->n[VKC_SITEMS], tot_v = C->n[VKC_VERTS], tot_i = C->n[VKC_INDS];
    for (uint32_t i0 = blockIdx.x * SE_BLOCK; i0 < n_items; i0 += gridDim.x * SE_BLOCK) {   // (block-uniform: the barriers below are safe)
        const uint32_t           i1 = min(i0 + SE_BLOCK, n_items);
        const unsigned long long o0 = offsets[i0], o1 = i1 < n_items ? offsets[i1] : ((unsigned long long)tot_v | ((unsigned long long)tot_i << 32));
        const uint32_t v0 = (uint32_t)(o0 & 0xffffffffull), v1 = (uint32_t)(o1 & 0xffffffffull), x0 = (uint32_t)(o0 >> 32), x1 = (uint32_t)(o1 >> 32);
        const uint32_t voff = v0 & ~1u, ioff = x0 & ~3u;
        // (joins of two vertices and a quad - a tiger's miter joins - gain nothing from the detour and pay its two barriers)
        const bool     staged = !direct && v1 - v0 > 3u * SE_BLOCK && v1 - voff <= SE_VERTS && x1 - ioff <= SE_INDS;
        const uint32_t item = i0 + threadIdx.x;
        if (item < i1) {
            const uint32_t   j  = item_job(a, item);
            const vkb_xform &xf = xforms[a.draws[a.job_draw[j]].xform_stroke & 0xFFFF];
            Out<2> o;
            o.v = verts; o.nv = 0; o.ni = 0;
            const unsigned long long off = offsets[item];
            o.vbase = (uint32_t)(off & 0xffffffffull);
            o.ibase = (uint32_t)(off >> 32);
            o.sn = staged ? s_v : snapped; o.voff = staged ? voff : 0u;
            o.idx = staged ? s_i : inds;   o.ioff = staged ? ioff : 0u;
#pragma unroll
            for (int q = 0; q < 6; q++) o.m[q] = xf.mat[q];
            o.W = (float)sd.width; o.H = (float)sd.full_height;
            o.yoff = (int32_t)(xf.band * sd.band_tiles) * VKB_TILE_FX - (int32_t)sd.origin_y * 256;
            stroke_item<2>(a, item, j, o, job_inverse);
        }
        if (staged) {
            __syncthreads();
            for (uint32_t q = threadIdx.x; 2 * q < v1 - voff; q += SE_BLOCK) {   // vertices, two to a 16-byte store
                const uint32_t g = voff + 2 * q;
                if (g >= v0 && g + 2 <= v1) *reinterpret_cast<int4 *>(snapped + g) = *reinterpret_cast<const int4 *>(s_v + 2 * q);
                else {
                    if (g >= v0 && g < v1) snapped[g] = s_v[2 * q];
                    if (g + 1 >= v0 && g + 1 < v1) snapped[g + 1] = s_v[2 * q + 1];
                }
            }
            for (uint32_t q = threadIdx.x; 4 * q < x1 - ioff; q += SE_BLOCK) {   // indices, four to a 16-byte store
                const uint32_t g = ioff + 4 * q;
                if (g >= x0 && g + 4 <= x1) *reinterpret_cast<uint4 *>(inds + g) = *reinterpret_cast<const uint4 *>(s_i + 4 * q);
                else {
#pragma unroll
                    for (uint32_t e = 0; e < 4; e++)
                        if (g + e >= x0 && g + e < x1) inds[g + e] = s_i[4 * q + e];
                }
            }
            __syncthreads();   // (the next chunk of a grid-stride block reuses the staging area)
        }
    }
}

// float segment lengths for the dash phase scan (one per stroke item; 0 where there is no segment)
__global__ void stroke_seglen_k(const float2 *pts, const uint32_t *job_sp, const uint32_t *job_base, uint32_t n_jobs, const uint32_t *sp_first,
                                const uint32_t *sp_count, const vkb_subpath *sps, const vkb_counts *C, float *seglen) {
    if (C->overflow) return;
    const uint32_t n_items = C->n[VKC_SITEMS];
    for (uint32_t item = blockIdx.x * blockDim.x + threadIdx.x; item <= n_items; item += gridDim.x * blockDim.x) {
        if (item == n_items) { seglen[item] = 0.f; break; }
        uint32_t lo = 0, hi = n_jobs;
        while (hi - lo > 1) {
            uint32_t mid = (lo + hi) >> 1;
            if (job_base[mid] <= item) lo = mid; else hi = mid;
        }
        uint32_t k = item - job_base[lo], s = job_sp[lo], first = sp_first[s], n = sp_count[s];
        float    L = 0.f;
        if (n >= 2) {
            uint32_t iR = k + 1 < n ? k + 1 : ((sps[s].flags & VKB_SP_CLOSED) ? 0 : n);
            if (iR < n) L = v2len(v2sub(ld(pts + first, iR), ld(pts + first, k)));
        }
        seglen[item] = L;
    }
}

// closed, undashed sub-paths: redirect the forward references of the closing quad to the first two
// vertices of the sub-path (vkvg_context.c:921-931)
__global__ void stroke_patch_closed_k(const vkb_draw *draws, const vkb_stroke *strokes, const uint32_t *job_draw, const uint32_t *job_sp, const uint32_t *job_base,
                                      uint32_t n_jobs, const vkb_subpath *sps, const uint32_t *sp_count, const unsigned long long *offsets,
                                      const vkb_counts *C, const uint32_t *job_inverse, uint32_t *inds) {
    uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_jobs || C->overflow) return;
    const uint32_t           n_items = C->n[VKC_SITEMS];
    const unsigned long long total   = (unsigned long long)C->n[VKC_VERTS] | ((unsigned long long)C->n[VKC_INDS] << 32);
    uint32_t s = job_sp[j];
    if (!(sps[s].flags & VKB_SP_CLOSED) || strokes[draws[job_draw[j]].xform_stroke >> 16].dash_count != 0 || sp_count[s] < 2) return;
    if ((j + 1 < n_jobs ? job_base[j + 1] : n_items) == job_base[j]) return;  // a job without work items (culled: it cannot touch the surface)
    unsigned long long a = offsets[job_base[j]];
    uint32_t           e = job_base[j] + sp_count[s];
    unsigned long long b = e < n_items ? offsets[e] : total;
    uint32_t           ia = (uint32_t)(a >> 32), ib = (uint32_t)(b >> 32), ii = (uint32_t)(a & 0xffffffffull);
    if (ib - ia < 6) return;
    uint32_t *t = inds + ib - 6;
    if (job_inverse[j]) { t[1] = ii + 1; t[4] = ii + 1; t[5] = ii; }
    else { t[1] = ii; t[4] = ii; t[5] = ii + 1; }
}

// a.n_items is the CAPACITY of the item space (grids are sized for it, up to VKB_STROKE_GRID blocks that stride); the item count is a.C->n[VKC_SITEMS]
#define VKB_STROKE_GRID (148u * 16u * 8u)
void vkb_launch_stroke_seglen(const StrokeArgs &a, float *seglen, cudaStream_t s) {
    stroke_seglen_k<<<min(vkb_div_up(a.n_items + 1, 256), VKB_STROKE_GRID), 256, 0, s>>>(a.pts, a.job_sp, a.job_base, a.n_jobs, a.sp_first, a.sp_count, a.sps, a.C, seglen);
    VKB_LAUNCHED();
}
void vkb_launch_stroke_count(const StrokeArgs &a, unsigned long long *counts, cudaStream_t s) {
    stroke_items_k<0><<<min(vkb_div_up(a.n_items, 128), VKB_STROKE_GRID), 128, 0, s>>>(item_args(a), a.C, counts, nullptr, nullptr, nullptr, nullptr);
    VKB_LAUNCHED();
}
static void launch_patch_closed(const StrokeArgs &a, const unsigned long long *offsets, uint32_t *inds, uint32_t *job_inverse, cudaStream_t s) {
    stroke_patch_closed_k<<<vkb_div_up(a.n_jobs, 128), 128, 0, s>>>(a.draws, a.strokes, a.job_draw, a.job_sp, a.job_base, a.n_jobs, a.sps, a.sp_count, offsets,
                                                                   a.C, job_inverse, inds);
    VKB_LAUNCHED();
}
// VKVG_B200_STROKE=legacy: the first emitter (float vertices, snapped by snap_verts_k afterwards); =direct: the fused vertex stage without the staging area
int vkb_stroke_emit_mode() {
    static int mode = -1;
    if (mode < 0) {
        const char *e = getenv("VKVG_B200_STROKE");
        mode = (e && !strcmp(e, "legacy")) ? 1 : (e && !strcmp(e, "direct")) ? 2 : 0;
    }
    return mode;
}
void vkb_launch_stroke_emit(const StrokeArgs &a, const unsigned long long *offsets, float2 *verts, uint32_t *inds, uint32_t *job_inverse, cudaStream_t s) {
    stroke_items_k<1><<<min(vkb_div_up(a.n_items, 128), VKB_STROKE_GRID), 128, 0, s>>>(item_args(a), a.C, nullptr, offsets, verts, inds, job_inverse);
    VKB_LAUNCHED();
    launch_patch_closed(a, offsets, inds, job_inverse, s);
}
void vkb_launch_stroke_emit_snapped(const StrokeArgs &a, const unsigned long long *offsets, const vkb_xform *xforms, const SurfaceDesc &sd, float2 *verts, int2 *snapped,
                                    uint32_t *inds, uint32_t *job_inverse, cudaStream_t s) {
    stroke_emit_k<<<min(vkb_div_up(a.n_items, SE_BLOCK), VKB_STROKE_GRID), SE_BLOCK, 0, s>>>(item_args(a), a.C, offsets, xforms, sd, vkb_stroke_emit_mode() == 2 ? 1 : 0, verts,
                                                                                          snapped, inds, job_inverse);
    VKB_LAUNCHED();
    launch_patch_closed(a, offsets, inds, job_inverse, s);
}
