// The packed command stream decoded ON THE DEVICE (vkvg_b200_submit).
//
// vkvg_b200_replay turns a stream into path elements, sub-paths, draws and side tables by issuing the vkvg_* calls one by one on the
// host: 5.5 ms of single-threaded recording for C2's 400 k commands while the GPU waits (the whole frame renders in 2 ms).  Here the
// stream itself is uploaded - cmds[i] = op | n_args << 8, args - and the same tables are built by kernels:
//   1. one exclusive scan over the commands of NF counters at once (arguments, elements, payload floats, sub-paths, draws, setters of
//      every state class ...): where each command's output goes, and - because setters are COUNTED - which setter of a class is the
//      latest one before any command (the k-th setter's command index is scattered into a list);
//   2. one thread per element / sub-path / draw / gradient writes its record (the records vkvg_api.cpp's recorder would have written);
//      the CTM commands, whose products must round exactly as on the host, are walked by one thread (there are few);
//   3. a census (counts the pipeline's host side sizes its launches from, the state the context is left in) is read back: one 256-byte
//      copy, the only round trip.
// The decoder accepts the REGULAR case of a subset of the commands (polygons, polylines, cubics, fills and strokes with solid or
// gradient sources, translations, canvases): every point kept, every sub-path of at least two points, paths ended by a draw.  Anything
// else - an op outside the subset, a NaN or repeated point (which the reference drops, shifting every later index), a close_path that
// the reference ignores - sets `irregular`, nothing is committed, and the host decodes the stream the old way.  Same results either
// way: tests/test_gpu_submit.py compares the two bit for bit.
#include "pipeline.h"
#include "../../include/vkvg_b200.h"  // (the op codes)
#include <float.h>

enum {  // counters of the command scan (all sums)
    F_ARGS, F_EL, F_DAT, F_SP, F_DRAW, F_CUBIC, F_NC, F_PB, F_SRC, F_RULE, F_LW, F_CAP, F_JOIN, F_MITER, F_DASH, F_DASHF, F_OPAC, F_XF, F_GRAD, F_SDRAW, NF
};
// counters that also have a LIST: list f holds, in order, the command indices that contributed to counter f
static __device__ __constant__ const int k_list_fields[] = {F_SP, F_DRAW, F_NC, F_PB, F_SRC, F_RULE, F_LW, F_CAP, F_JOIN, F_MITER, F_DASH, F_OPAC, F_XF, F_GRAD};
#define VKD_N_LISTS 14
#define VKD_EQUF(a, b) (fabsf((a) - (b)) <= FLT_EPSILON)

__device__ __forceinline__ bool op_is_point(uint32_t op) {
    return op == VKVG_B200_OP_MOVE_TO || op == VKVG_B200_OP_LINE_TO || op == VKVG_B200_OP_CURVE_TO || op == VKVG_B200_OP_POLYLINE;
}
// what one command adds to every counter; returns false for a command outside the decodable subset (or with a wrong argument count)
__device__ __forceinline__ bool cmd_contrib(uint32_t cmd, uint32_t (&c)[NF]) {
    const uint32_t op = cmd & 0xFF, na = cmd >> 8;
#pragma unroll
    for (int f = 0; f < NF; f++) c[f] = 0;
    c[F_ARGS] = na;
    c[F_NC]   = 1;  // every command but line_to / curve_to ends the sub-path under construction
    bool ok = true;
    switch (op) {
    case VKVG_B200_OP_MOVE_TO: c[F_EL] = 1; c[F_DAT] = 2; c[F_SP] = 1; ok = na == 2; break;
    case VKVG_B200_OP_LINE_TO: c[F_EL] = 1; c[F_DAT] = 2; c[F_NC] = 0; ok = na == 2; break;
    case VKVG_B200_OP_CURVE_TO: c[F_EL] = 1; c[F_DAT] = 9; c[F_CUBIC] = 1; c[F_NC] = 0; ok = na == 6; break;
    case VKVG_B200_OP_POLYLINE: c[F_EL] = na / 2; c[F_DAT] = na; c[F_SP] = 1; ok = na >= 4 && (na & 1) == 0; break;
    case VKVG_B200_OP_CLOSE_PATH: ok = na == 0; break;
    case VKVG_B200_OP_NEW_PATH: c[F_PB] = 1; ok = na == 0; break;
    case VKVG_B200_OP_FILL: c[F_PB] = 1; c[F_DRAW] = 1; ok = na == 0; break;
    case VKVG_B200_OP_FILL_PRESERVE: c[F_DRAW] = 1; ok = na == 0; break;
    case VKVG_B200_OP_STROKE: c[F_PB] = 1; c[F_DRAW] = 1; c[F_SDRAW] = 1; ok = na == 0; break;
    case VKVG_B200_OP_STROKE_PRESERVE: c[F_DRAW] = 1; c[F_SDRAW] = 1; ok = na == 0; break;
    case VKVG_B200_OP_SET_SOURCE_RGBA: c[F_SRC] = 1; ok = na == 4; break;
    case VKVG_B200_OP_SET_SOURCE_LINEAR: c[F_SRC] = 1; c[F_GRAD] = 1; ok = na >= 4 + 10 && na <= 4 + 80 && (na - 4) % 5 == 0; break;
    case VKVG_B200_OP_SET_SOURCE_RADIAL: c[F_SRC] = 1; c[F_GRAD] = 1; ok = na >= 6 + 10 && na <= 6 + 80 && (na - 6) % 5 == 0; break;
    case VKVG_B200_OP_SET_FILL_RULE: c[F_RULE] = 1; ok = na == 1; break;
    case VKVG_B200_OP_SET_LINE_WIDTH: c[F_LW] = 1; ok = na == 1; break;
    case VKVG_B200_OP_SET_LINE_CAP: c[F_CAP] = 1; ok = na == 1; break;
    case VKVG_B200_OP_SET_LINE_JOIN: c[F_JOIN] = 1; ok = na == 1; break;
    case VKVG_B200_OP_SET_MITER_LIMIT: c[F_MITER] = 1; ok = na == 1; break;
    case VKVG_B200_OP_SET_DASH: c[F_DASH] = 1; c[F_DASHF] = na ? na - 1 : 0; ok = na >= 1 && na <= 1 + VKB_MAX_DASHES; break;
    case VKVG_B200_OP_SET_OPACITY: c[F_OPAC] = 1; ok = na == 1; break;
    case VKVG_B200_OP_IDENTITY_MATRIX: c[F_XF] = 1; ok = na == 0; break;
    case VKVG_B200_OP_TRANSLATE: c[F_XF] = 1; ok = na == 2; break;
    case VKVG_B200_OP_SET_CANVAS: c[F_XF] = 1; ok = na == 1; break;
    default: ok = false; break;
    }
    return ok;
}

// ---- the scan: S[i][f] = sum of counter f over commands 0 .. i-1; row n_cmds = the totals ----
#define VKD_BLOCK 256
#define VKD_ITEMS 4
#define VKD_CHUNK (VKD_BLOCK * VKD_ITEMS)
__global__ void __launch_bounds__(VKD_BLOCK) vkd_scan_reduce_k(const uint32_t *cmds, uint32_t n, uint32_t *blocksum, uint32_t *irregular) {
    uint32_t acc[NF];
#pragma unroll
    for (int f = 0; f < NF; f++) acc[f] = 0;
    const uint32_t base = blockIdx.x * VKD_CHUNK + threadIdx.x * VKD_ITEMS;
    bool bad = false;
    for (int k = 0; k < VKD_ITEMS; k++)
        if (base + k < n) {
            uint32_t c[NF];
            bad |= !cmd_contrib(cmds[base + k], c);
#pragma unroll
            for (int f = 0; f < NF; f++) acc[f] += c[f];
        }
    if (bad) atomicOr(irregular, 1u);
    __shared__ uint32_t red[NF][VKD_BLOCK / 32];
#pragma unroll
    for (int f = 0; f < NF; f++) {
        uint32_t v = acc[f];
        for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if ((threadIdx.x & 31) == 0) red[f][threadIdx.x >> 5] = v;
    }
    __syncthreads();
    if (threadIdx.x < NF) {
        uint32_t v = 0;
        for (int w = 0; w < VKD_BLOCK / 32; w++) v += red[threadIdx.x][w];
        blocksum[(size_t)blockIdx.x * NF + threadIdx.x] = v;
    }
}
__global__ void vkd_scan_sums_k(uint32_t *blocksum, uint32_t n_blocks, uint32_t *totals_row) {
    const uint32_t f = threadIdx.x;
    if (f >= NF) return;
    uint32_t run = 0;
    for (uint32_t b = 0; b < n_blocks; b++) {
        const uint32_t v = blocksum[(size_t)b * NF + f];
        blocksum[(size_t)b * NF + f] = run;
        run += v;
    }
    totals_row[f] = run;
}
__global__ void __launch_bounds__(VKD_BLOCK) vkd_scan_apply_k(const uint32_t *cmds, uint32_t n, const uint32_t *blocksum, uint32_t *S) {
    const uint32_t base = blockIdx.x * VKD_CHUNK + threadIdx.x * VKD_ITEMS;
    uint32_t       c[VKD_ITEMS][NF], acc[NF];
#pragma unroll
    for (int f = 0; f < NF; f++) acc[f] = 0;
    for (int k = 0; k < VKD_ITEMS; k++) {
        if (base + k < n) cmd_contrib(cmds[base + k], c[k]);
        else {
#pragma unroll
            for (int f = 0; f < NF; f++) c[k][f] = 0;
        }
#pragma unroll
        for (int f = 0; f < NF; f++) acc[f] += c[k][f];
    }
#pragma unroll
    for (int f = 0; f < NF; f++) {
        uint32_t tot;
        uint32_t e = block_excl_scan<uint32_t, VKD_BLOCK>(acc[f], tot) + blocksum[(size_t)blockIdx.x * NF + f];
        for (int k = 0; k < VKD_ITEMS; k++) {
            if (base + k < n) S[(size_t)(base + k) * NF + f] = e;
            e += c[k][f];
        }
    }
}

struct DecodeBufs {
    const uint32_t *cmds;
    const float    *args;
    uint32_t        n_cmds;
    uint32_t       *S;       // (n_cmds + 1) x NF
    uint32_t       *lists;   // the VKD_N_LISTS lists, back to back (list_base below)
    uint32_t       *irregular;
    // outputs: the batch as vkvg_api.cpp's recorder would have left it (vkb_types.h)
    uint32_t       *elem_hdr;
    float          *elem_data;
    vkb_subpath    *subpaths;
    vkb_draw       *draws;
    vkb_xform      *xforms;     // entry 0 = the state at entry, entry k + 1 = after the k-th CTM / canvas command
    float          *xf_scale;   // 2 floats per xform entry: vkvg_matrix_get_scale of its matrix
    vkb_stroke     *strokes;    // one per stroke draw
    vkb_gradient   *grads;      // one per gradient setter
    float          *dashes;
    vkb_decode_census *census;
    uint32_t       *sp_null;    // per sub-path: curve_to commands the reference skips (zeroed before vkd_elems_k)
};
typedef vkb_decode_init DecodeInit;  // the context's state when the stream starts
__device__ __forceinline__ const uint32_t *tot_row(const DecodeBufs &b) { return b.S + (size_t)b.n_cmds * NF; }
__device__ __forceinline__ uint32_t list_base(const DecodeBufs &b, int f) {  // where list f starts inside b.lists
    uint32_t o = 0;
    for (int k = 0; k < VKD_N_LISTS && k_list_fields[k] != f; k++) o += tot_row(b)[k_list_fields[k]];
    return o;
}
__device__ __forceinline__ uint32_t Sv(const DecodeBufs &b, uint32_t i, int f) { return b.S[(size_t)i * NF + f]; }

__global__ void __launch_bounds__(256) vkd_lists_k(DecodeBufs b) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= b.n_cmds) return;
    uint32_t c[NF];
    cmd_contrib(b.cmds[i], c);
    uint32_t o = 0;
    const uint32_t *T = tot_row(b);
    for (int k = 0; k < VKD_N_LISTS; k++) {
        const int f = k_list_fields[k];
        if (c[f]) b.lists[o + Sv(b, i, f)] = i;
        o += T[f];
    }
}
// the command of class f in force at command i (the last one before it), or -1
__device__ __forceinline__ int latest(const DecodeBufs &b, uint32_t i, int f) {
    const uint32_t k = Sv(b, i, f);
    return k ? (int)b.lists[list_base(b, f) + k - 1] : -1;
}
__device__ __forceinline__ void end_point_of(const DecodeBufs &b, uint32_t i, float &x, float &y) {  // last two arguments of point command i
    const uint32_t a = Sv(b, i, F_ARGS), na = b.cmds[i] >> 8;
    x = b.args[a + na - 2]; y = b.args[a + na - 1];
}

// ---- elements ----
__global__ void __launch_bounds__(256) vkd_elems_k(DecodeBufs b) {
    const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t n_el = tot_row(b)[F_EL];
    if (e >= n_el) return;
    uint32_t lo = 0, hi = b.n_cmds;  // the command that produces element e: the last one whose element offset is <= e
    while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if (Sv(b, mid, F_EL) <= e) lo = mid; else hi = mid;
    }
    const uint32_t i = lo, op = b.cmds[i] & 0xFF, a = Sv(b, i, F_ARGS), k = e - Sv(b, i, F_EL);
    uint32_t       d = Sv(b, i, F_DAT);
    bool           bad = false;
    if (op == VKVG_B200_OP_CURVE_TO) {
        // _curve_to (src/vkvg_context.c:541-566): needs a current point, which the previous command must have produced
        if (i == 0 || !op_is_point(b.cmds[i - 1] & 0xFF)) { atomicOr(b.irregular, 2u); return; }
        float cx, cy;
        end_point_of(b, i - 1, cx, cy);
        const float *p = b.args + a;
        bad = !(isfinite(p[0]) && isfinite(p[1]) && isfinite(p[2]) && isfinite(p[3]) && isfinite(p[4]) && isfinite(p[5]));
        if (VKD_EQUF(p[0], p[2]) && VKD_EQUF(p[2], p[4]) && VKD_EQUF(p[1], p[3]) && VKD_EQUF(p[3], p[5]) && VKD_EQUF(cx, p[0]) && VKD_EQUF(cy, p[1])) {
            // a curve that goes nowhere: _curve_to returns without adding anything and without moving the current point.  Its slot becomes an
            // element without points (and is taken off the sub-path's point count); only when the point it "ends" on is bit for bit the
            // current one, so that whatever follows starts from the same floats as on the host
            if (p[4] != cx || p[5] != cy) bad = true;
            else {
                b.elem_data[d] = cx; b.elem_data[d + 1] = cy;
                b.elem_hdr[e]  = VKB_EL_NONE | (d << VKB_EL_PAYLOAD_SHIFT);
                atomicAdd(&b.sp_null[Sv(b, i, F_SP) - 1], 1u);
                return;
            }
        }
        const uint32_t xf = Sv(b, i, F_XF);
        const float    tol = fabsf(0.25f / fmaxf(b.xf_scale[2 * xf], b.xf_scale[2 * xf + 1]));
        float *o = b.elem_data + d;
        o[0] = cx; o[1] = cy;
        for (int q = 0; q < 6; q++) o[2 + q] = p[q];
        o[8] = tol;
        b.elem_hdr[e] = VKB_EL_CUBIC | VKB_EL_CURVED | (d << VKB_EL_PAYLOAD_SHIFT);
    } else {
        const float x = b.args[a + 2 * k], y = b.args[a + 2 * k + 1];
        d += 2 * k;
        bad = isnan(x) || isnan(y);  // _add_point drops these
        // _line_to drops a point equal to the current one (the first point of a polyline / a move_to is never compared)
        if (op == VKVG_B200_OP_LINE_TO) {
            if (i == 0 || !op_is_point(b.cmds[i - 1] & 0xFF)) { atomicOr(b.irregular, 2u); return; }  // line_to without a current point acts as a move_to
            float cx, cy;
            end_point_of(b, i - 1, cx, cy);
            bad |= VKD_EQUF(cx, x) && VKD_EQUF(cy, y);
        } else if (op == VKVG_B200_OP_POLYLINE && k > 0) {
            bad |= VKD_EQUF(b.args[a + 2 * k - 2], x) && VKD_EQUF(b.args[a + 2 * k - 1], y);
        }
        b.elem_data[d] = x; b.elem_data[d + 1] = y;
        b.elem_hdr[e]  = VKB_EL_POINT | (d << VKB_EL_PAYLOAD_SHIFT);
    }
    if (bad) atomicOr(b.irregular, 4u);
}

// ---- sub-paths: one per move_to / polyline, ended by the first later command that is not a line_to / curve_to ----
__global__ void __launch_bounds__(256) vkd_subpaths_k(DecodeBufs b) {
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t *T = tot_row(b);
    if (s >= T[F_SP]) return;
    const uint32_t i = b.lists[list_base(b, F_SP) + s];
    const uint32_t nc = i + 1 < b.n_cmds ? Sv(b, i + 1, F_NC) : T[F_NC];   // non-continuation commands before i + 1
    if (nc >= T[F_NC]) { atomicOr(b.irregular, 8u); return; }              // the stream ends inside the sub-path: the path would stay open
    const uint32_t j = b.lists[list_base(b, F_NC) + nc];                   // the command that ends it
    const uint32_t first = Sv(b, i, F_EL), n = Sv(b, j, F_EL) - first;
    const uint32_t cubics = Sv(b, j, F_CUBIC) - Sv(b, i, F_CUBIC);
    const uint32_t sp_points = n + 2 * cubics - 3 * b.sp_null[s];  // the host's running count: a cubic stands for at least three points, a skipped one for none
    uint32_t       flags = 0;
    bool           bad = sp_points < 2;         // a lone move_to is dropped by _finish_path (and its element with it)
    if ((b.cmds[j] & 0xFF) == VKVG_B200_OP_CLOSE_PATH) {  // vkvg_close_path, src/vkvg_context.c:350-373
        if (sp_points < 3) bad = true;          // ignored by the host: the sub-path stays open
        else {
            flags = VKB_SP_CLOSED;
            float fx = b.args[Sv(b, i, F_ARGS)], fy = b.args[Sv(b, i, F_ARGS) + 1], lx, ly;
            end_point_of(b, j - 1, lx, ly);
            if (VKD_EQUF(lx, fx) && VKD_EQUF(ly, fy)) {
                if (sp_points < 4) bad = true;
                else flags |= VKB_SP_DROP_LAST;
            }
        }
    }
    if ((b.cmds[j] & 0xFF) == VKVG_B200_OP_NEW_PATH) bad = true;  // vkvg_new_path drops the elements of an unfinished sub-path
    b.subpaths[s] = vkb_subpath{first, n, flags, 0};
    if (n > VKB_SP_LONG) atomicMax(&b.census->max_sp_elems, n);
    if (bad) atomicOr(b.irregular, 16u);
}

// ---- the CTM / canvas commands, in order (their products must round as on the host: one thread) ----
__device__ __forceinline__ void xf_scale_of(const float *m, float *out) {  // vkvg_matrix_get_scale: double sqrt, float store
    out[0] = (float)sqrt((double)(m[0] * m[0] + m[2] * m[2]));
    out[1] = (float)sqrt((double)(m[1] * m[1] + m[3] * m[3]));
}
__global__ void vkd_xforms_k(DecodeBufs b, DecodeInit init) {
    if (threadIdx.x || blockIdx.x) return;
    float    m[6];   // xx yx xy yy x0 y0
    uint32_t band = init.band;
    for (int q = 0; q < 6; q++) m[q] = init.mat[q];
    const uint32_t n = tot_row(b)[F_XF], base = list_base(b, F_XF);
    for (uint32_t k = 0;; k++) {
        vkb_xform x;
        for (int q = 0; q < 6; q++) x.mat[q] = m[q];
        x.band = band; x.pad = 0;
        b.xforms[k] = x;
        xf_scale_of(m, b.xf_scale + 2 * k);
        if (k == n) break;
        const uint32_t i = b.lists[base + k], op = b.cmds[i] & 0xFF, a = Sv(b, i, F_ARGS);
        if (op == VKVG_B200_OP_IDENTITY_MATRIX) { m[0] = 1; m[1] = 0; m[2] = 0; m[3] = 1; m[4] = 0; m[5] = 0; }
        else if (op == VKVG_B200_OP_TRANSLATE) {
            // vkvg_matrix_translate: result = T x m with T = (1 0 0 1 tx ty), every product and sum as vkvg_matrix_multiply writes them
            const float tx = b.args[a], ty = b.args[a + 1];
            const float axx = 1.f, ayx = 0.f, axy = 0.f, ayy = 1.f;
            float r[6];
            r[0] = axx * m[0] + ayx * m[2];
            r[1] = axx * m[1] + ayx * m[3];
            r[2] = axy * m[0] + ayy * m[2];
            r[3] = axy * m[1] + ayy * m[3];
            r[4] = tx * m[0] + ty * m[2] + m[4];
            r[5] = tx * m[1] + ty * m[3] + m[5];
            for (int q = 0; q < 6; q++) m[q] = r[q];
        } else {  // SET_CANVAS
            const float v = b.args[a];
            if (!(v >= 0.0f && v <= 16777216.0f)) atomicOr(b.irregular, 32u);
            else band = (uint32_t)v;
        }
    }
}

// ---- gradients: vkvg_pattern_create_linear / _radial + add_color_stop + _update_cur_pattern (src/vkvg_pattern.c:95-167,
//      src/vkvg_context_internal.c:774-826) through the CTM in force at the setter ----
__device__ __forceinline__ void xf_point(const float *m, float &x, float &y) {
    const float nx = (m[0] * x + m[2] * y), ny = (m[1] * x + m[3] * y);
    x = nx + m[4]; y = ny + m[5];
}
__device__ __forceinline__ void xf_distance(const float *m, float &x, float &y) {
    const float nx = (m[0] * x + m[2] * y), ny = (m[1] * x + m[3] * y);
    x = nx; y = ny;
}
__global__ void __launch_bounds__(128) vkd_grads_k(DecodeBufs b) {
    const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= tot_row(b)[F_GRAD]) return;
    const uint32_t i = b.lists[list_base(b, F_GRAD) + g], op = b.cmds[i] & 0xFF, a = Sv(b, i, F_ARGS), na = b.cmds[i] >> 8;
    const float   *p = b.args + a;
    const float   *m = b.xforms[Sv(b, i, F_XF)].mat;
    vkb_gradient   G;
    memset(&G, 0, sizeof G);
    const int np = op == VKVG_B200_OP_SET_SOURCE_LINEAR ? 4 : 6;
    if (np == 4) { G.cp[0][0] = p[0]; G.cp[0][1] = p[1]; G.cp[0][2] = p[2]; G.cp[0][3] = p[3]; }
    else {  // vkvg_pattern_edit_radial
        float cx0 = p[0], cy0 = p[1], r0 = p[2];
        const float cx1 = p[3], cy1 = p[4], r1 = p[5];
        float c0x = cx0, c0y = cy0;
        if (r0 > r1 - 1.0f) r0 = r1 - 1.0f;
        const float ux = c0x - cx1, uy = c0y - cy1;
        const float l  = sqrtf(ux * ux + uy * uy);
        if (l + r0 + 1.0f >= r1) {
            const float vx = ux / l, vy = uy / l, mm = r1 - r0 - 1.0f;
            c0x = cx1 + vx * mm; c0y = cy1 + vy * mm;
        }
        G.cp[0][0] = c0x; G.cp[0][1] = c0y; G.cp[0][2] = r0; G.cp[0][3] = 0;
        G.cp[1][0] = cx1; G.cp[1][1] = cy1; G.cp[1][2] = r1; G.cp[1][3] = 0;
    }
    const uint32_t ns = (na - np) / 5;
    const float   *s = p + np;
    for (uint32_t j = 0; j < ns && j < 16; j++) {
        const float off = s[5 * j], r = s[5 * j + 1], gg = s[5 * j + 2], bb = s[5 * j + 3], al = s[5 * j + 4];
        G.colors[j][0] = al * r; G.colors[j][1] = al * gg; G.colors[j][2] = al * bb; G.colors[j][3] = al;
        G.stops[j] = off;
    }
    G.count = ns < 16 ? ns : 16;
    xf_point(m, G.cp[0][0], G.cp[0][1]);
    if (np == 4) xf_point(m, G.cp[0][2], G.cp[0][3]);
    else {
        xf_point(m, G.cp[1][0], G.cp[1][1]);
        xf_distance(m, G.cp[0][2], G.cp[0][3]);
        xf_distance(m, G.cp[1][2], G.cp[0][3]);
    }
    b.grads[g] = G;
}
__global__ void __launch_bounds__(128) vkd_dashes_k(DecodeBufs b) {
    const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= tot_row(b)[F_DASH]) return;
    const uint32_t i = b.lists[list_base(b, F_DASH) + q], a = Sv(b, i, F_ARGS), na = b.cmds[i] >> 8, o = Sv(b, i, F_DASHF);
    for (uint32_t k = 0; k + 1 < na; k++) b.dashes[VKB_MAX_DASHES + o + k] = b.args[a + 1 + k];   // (the first VKB_MAX_DASHES slots hold the pattern in force at entry)
}

// ---- draws ----
__device__ __forceinline__ uint32_t rgbaf_dev(float r, float g, float bl, float a) {  // CreateRgbaf, src/vkvg_context_internal.h:60-62
    return (((uint32_t)(a * 255.0f) & 0xFF) << 24) | (((uint32_t)(bl * a * 255.0f) & 0xFF) << 16) | (((uint32_t)(g * a * 255.0f) & 0xFF) << 8) |
           ((uint32_t)(r * a * 255.0f) & 0xFF);
}
__device__ __forceinline__ float arc_step_dev(const float *scale, float radius) {  // _get_arc_step, src/vkvg_context_internal.c:245-252
    const float PIF = 3.14159265358979323846f;
    const float r = radius * fabsf(fmaxf(scale[0], scale[1]));
    if (r < 30.0f) return fminf(PIF / 3.f, PIF / r);
    return fminf(PIF / 3.f, PIF / (r * 0.4f));
}
__global__ void __launch_bounds__(128) vkd_draws_k(DecodeBufs b, DecodeInit init) {
    const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t *T = tot_row(b);
    if (q >= T[F_DRAW]) return;
    const uint32_t i = b.lists[list_base(b, F_DRAW) + q], op = b.cmds[i] & 0xFF;
    const bool     stroke = op == VKVG_B200_OP_STROKE || op == VKVG_B200_OP_STROKE_PRESERVE;
    const int      pb = latest(b, i, F_PB);
    const uint32_t first_sp = pb >= 0 ? Sv(b, (uint32_t)pb, F_SP) : 0u, n_sp = Sv(b, i, F_SP) - first_sp;
    auto arg1 = [&](int f, float dflt) { const int c = latest(b, i, f); return c >= 0 ? b.args[Sv(b, (uint32_t)c, F_ARGS)] : dflt; };
    vkb_draw d;
    d.kind = stroke ? VKB_DRAW_STROKE : VKB_DRAW_FILL;
    uint32_t rule = init.rule;
    {
        const int c = latest(b, i, F_RULE);
        if (c >= 0) {
            const float v = b.args[Sv(b, (uint32_t)c, F_ARGS)];
            if (!(v >= 0.0f && v <= 1.0f)) atomicOr(b.irregular, 64u);
            rule = (uint32_t)(int)v;   // vkvg_fill_rule_t: 0 even-odd, 1 non-zero
        }
    }
    uint32_t pat = init.pattern, color = init.color, grad = T[F_GRAD];  // (the gradient in force at entry sits behind the stream's own)
    {
        const int c = latest(b, i, F_SRC);
        if (c >= 0) {
            const uint32_t sop = b.cmds[c] & 0xFF, a = Sv(b, (uint32_t)c, F_ARGS);
            if (sop == VKVG_B200_OP_SET_SOURCE_RGBA) { pat = VKB_PAT_SOLID; color = rgbaf_dev(b.args[a], b.args[a + 1], b.args[a + 2], b.args[a + 3]); }
            else {
                pat  = sop == VKVG_B200_OP_SET_SOURCE_LINEAR ? VKB_PAT_LINEAR : VKB_PAT_RADIAL;
                grad = Sv(b, (uint32_t)c, F_GRAD);
                // (curColor keeps the last solid colour on the host; a gradient draw does not read it)
                const int cs = c;  (void)cs;
            }
        }
    }
    d.rule_pattern = (stroke ? VKB_RULE_COUNT : (rule == 0 ? VKB_RULE_EVEN_ODD : VKB_RULE_NON_ZERO)) | (pat << 8) | (init.bop << 16);
    d.first_subpath = first_sp; d.n_subpaths = n_sp;
    d.color = color; d.opacity = arg1(F_OPAC, init.opacity); d.gradient = grad;
    const uint32_t xf = Sv(b, i, F_XF);
    d.xform_stroke = xf;
    if (stroke) {
        const uint32_t sidx = Sv(b, i, F_SDRAW);
        vkb_stroke st;
        memset(&st, 0, sizeof st);
        const float lw = arg1(F_LW, init.lw), miter = arg1(F_MITER, init.miter);
        st.hw = lw * 0.5f; st.lhMax = miter * lw; st.arcStep = arc_step_dev(b.xf_scale + 2 * xf, st.hw);
        const float jv = arg1(F_JOIN, (float)init.join), cv = arg1(F_CAP, (float)init.cap);
        if (!(jv >= 0.0f && jv <= 2.0f) || !(cv >= 0.0f && cv <= 2.0f)) atomicOr(b.irregular, 64u);
        st.join = (uint32_t)(int)jv; st.cap = (uint32_t)(int)cv;
        const int dc = latest(b, i, F_DASH);
        if (dc >= 0) {
            const uint32_t a = Sv(b, (uint32_t)dc, F_ARGS), na = b.cmds[dc] >> 8;
            st.dash_count = na - 1; st.dash_first = VKB_MAX_DASHES + Sv(b, (uint32_t)dc, F_DASHF); st.dash_offset = b.args[a];
            float tot = 0;
            for (uint32_t k = 0; k + 1 < na; k++) tot += b.args[a + 1 + k];
            if (st.dash_count && tot == 0) atomicOr(b.irregular, 128u);  // VKVG_STATUS_INVALID_DASH on the host
        } else { st.dash_count = init.dash_count; st.dash_first = 0; st.dash_offset = init.dash_offset; }
        b.strokes[sidx] = st;
        d.xform_stroke |= sidx << 16;
        atomicAdd(&b.census->n_sjobs, n_sp);
        if (n_sp) atomicAdd(&b.census->n_sdraws, 1u);
        if (st.dash_count && n_sp) b.census->any_dash = 1;
    } else {
        atomicAdd(&b.census->n_fjobs, n_sp);
        if (rule != 0 && n_sp) b.census->nz_any = 1;
    }
    b.draws[q] = d;
}
// counts the host needs, and what the context's state is after the last command
__global__ void vkd_census_k(DecodeBufs b, DecodeInit init) {
    if (threadIdx.x || blockIdx.x) return;
    const uint32_t *T = tot_row(b);
    vkb_decode_census *c = b.census;
    c->n_elems = T[F_EL]; c->n_data = T[F_DAT]; c->n_subpaths = T[F_SP]; c->n_draws = T[F_DRAW]; c->n_curves = T[F_CUBIC];
    c->n_grads = T[F_GRAD]; c->n_dash_floats = VKB_MAX_DASHES + T[F_DASHF]; c->n_xforms = T[F_XF] + 1; c->n_strokes = T[F_SDRAW];
    const int fields[9] = {F_SRC, F_RULE, F_LW, F_CAP, F_JOIN, F_MITER, F_DASH, F_OPAC, F_GRAD};
    for (int k = 0; k < 9; k++) {
        c->last_setter[k]     = T[fields[k]] ? (int32_t)b.lists[list_base(b, fields[k]) + T[fields[k]] - 1] : -1;
        c->last_setter_arg[k] = c->last_setter[k] >= 0 ? Sv(b, (uint32_t)c->last_setter[k], F_ARGS) : 0u;
    }
    const vkb_xform x = b.xforms[T[F_XF]];
    for (int q = 0; q < 6; q++) c->final_mat[q] = x.mat[q];
    c->final_band = x.band;
    // the path must be empty when the stream ends: no sub-path started after the last fill / stroke / new_path
    const uint32_t last_pb_sp = T[F_PB] ? Sv(b, b.lists[list_base(b, F_PB) + T[F_PB] - 1], F_SP) : 0u;
    if (T[F_SP] != last_pb_sp) atomicOr(b.irregular, 256u);
    if (T[F_XF] + 1 > 65000u || T[F_SDRAW] > 65000u) atomicOr(b.irregular, 512u);  // side tables are addressed with 16 bits
    for (uint32_t k = 0; k < VKB_MAX_DASHES; k++) b.dashes[k] = k < init.dash_count ? init.dashes[k] : 0.0f;
    b.grads[T[F_GRAD]] = init.grad;
    __threadfence();
    c->irregular = *b.irregular;
}

size_t   vkd_scan_words(uint32_t n_cmds) { return ((size_t)n_cmds + 1) * NF; }
static uint32_t vkd_scan_blocks(uint32_t n_cmds) { return vkb_div_up(n_cmds ? n_cmds : 1, VKD_CHUNK); }
size_t   vkd_blocksum_words(uint32_t n_cmds) { return (size_t)vkd_scan_blocks(n_cmds) * NF; }
uint32_t vkd_n_fields() { return NF; }
vkd_totals vkd_read_totals(const uint32_t *T) {
    vkd_totals t;
    t.n_elems = T[F_EL]; t.n_data = T[F_DAT]; t.n_subpaths = T[F_SP]; t.n_draws = T[F_DRAW]; t.n_grads = T[F_GRAD];
    t.n_dash_floats = VKB_MAX_DASHES + T[F_DASHF]; t.n_xforms = T[F_XF] + 1; t.n_strokes = T[F_SDRAW];
    const int lf[VKD_N_LISTS] = {F_SP, F_DRAW, F_NC, F_PB, F_SRC, F_RULE, F_LW, F_CAP, F_JOIN, F_MITER, F_DASH, F_OPAC, F_XF, F_GRAD};
    t.n_list_entries = 0;
    for (int k = 0; k < VKD_N_LISTS; k++) t.n_list_entries += T[lf[k]];
    return t;
}
void vkb_launch_decode_scan(const uint32_t *cmds, uint32_t n_cmds, uint32_t *S, uint32_t *blocksum, uint32_t *irregular, cudaStream_t st) {
    const uint32_t nb = vkd_scan_blocks(n_cmds);
    vkd_scan_reduce_k<<<nb, VKD_BLOCK, 0, st>>>(cmds, n_cmds, blocksum, irregular);
    VKB_LAUNCHED();
    vkd_scan_sums_k<<<1, 32, 0, st>>>(blocksum, nb, S + (size_t)n_cmds * NF);
    VKB_LAUNCHED();
    vkd_scan_apply_k<<<nb, VKD_BLOCK, 0, st>>>(cmds, n_cmds, blocksum, S);
    VKB_LAUNCHED();
}
void vkb_launch_decode_emit(const uint32_t *cmds, const float *args, uint32_t n_cmds, const vkd_totals &t, uint32_t *S, uint32_t *lists, uint32_t *irregular,
                            uint32_t *elem_hdr, float *elem_data, vkb_subpath *subpaths, vkb_draw *draws, vkb_xform *xforms, float *xf_scale, vkb_stroke *strokes,
                            vkb_gradient *grads, float *dashes, vkb_decode_census *census, uint32_t *sp_null, const vkb_decode_init &in, cudaStream_t st) {
    DecodeBufs b = {cmds, args, n_cmds, S, lists, irregular, elem_hdr, elem_data, subpaths, draws, xforms, xf_scale, strokes, grads, dashes, census, sp_null};
    VKB_CUDA_OK(cudaMemsetAsync(sp_null, 0, ((size_t)t.n_subpaths + 1) * 4, st));
    const DecodeInit &init = in;
    vkd_lists_k<<<vkb_div_up(n_cmds, 256), 256, 0, st>>>(b);
    VKB_LAUNCHED();
    vkd_xforms_k<<<1, 32, 0, st>>>(b, init);
    VKB_LAUNCHED();
    if (t.n_elems) { vkd_elems_k<<<vkb_div_up(t.n_elems, 256), 256, 0, st>>>(b); VKB_LAUNCHED(); }
    if (t.n_subpaths) { vkd_subpaths_k<<<vkb_div_up(t.n_subpaths, 256), 256, 0, st>>>(b); VKB_LAUNCHED(); }
    if (t.n_grads) { vkd_grads_k<<<vkb_div_up(t.n_grads, 128), 128, 0, st>>>(b); VKB_LAUNCHED(); }
    if (t.n_dash_floats > VKB_MAX_DASHES) { vkd_dashes_k<<<vkb_div_up(n_cmds, 128), 128, 0, st>>>(b); VKB_LAUNCHED(); }
    if (t.n_draws) { vkd_draws_k<<<vkb_div_up(t.n_draws, 128), 128, 0, st>>>(b, init); VKB_LAUNCHED(); }
    vkd_census_k<<<1, 32, 0, st>>>(b, init);
    VKB_LAUNCHED();
}
