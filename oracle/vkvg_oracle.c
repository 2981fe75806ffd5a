/* TEST INFRASTRUCTURE ONLY — see vkvg_oracle.h for scope, pinning and who may load this.
 *
 * Plain sequential C.  Every function that restates reference behaviour cites the reference
 * file:line it follows.  Float arithmetic is kept in the reference's evaluation order and types
 * (float vs double literals) and the file is compiled with -ffp-contract=off so results match a
 * baseline x86-64 gcc build of the reference (no FMA).
 */
#include "vkvg_oracle.h"
#include <float.h>
#include <math.h>
#include <stdbool.h>
#include <stdlib.h>
#include <string.h>

#define PIF 3.14159265358979323846f   /* M_PIF   src/vkvg_internal.h:43 */
#define PIF_2 1.57079632679489661923f /* M_PIF_2 src/vkvg_internal.h:44 */
#define TWO_OVER_PIF 0.63661977236758134308f /* M_2_PIF = 2/pi (NOT 2*pi) src/vkvg_internal.h:45 */
#define EQUF(a, b) (fabsf((a) - (b)) <= FLT_EPSILON) /* src/vkvg_internal.h:72 */

#define P_CLOSED 0x80000000u /* src/vkvg_internal.h:62-68 */
#define P_CURVES 0x40000000u
#define P_CONVEX 0x20000000u
#define P_MASK 0x1FFFFFFFu

typedef struct { float x, y; } v2;
static inline v2 v2add(v2 a, v2 b) { return (v2){a.x + b.x, a.y + b.y}; }
static inline v2 v2sub(v2 a, v2 b) { return (v2){a.x - b.x, a.y - b.y}; }
static inline v2 v2mul(v2 a, float m) { return (v2){a.x * m, a.y * m}; }
static inline v2 v2div(v2 a, float m) { return (v2){a.x / m, a.y / m}; }
static inline v2 v2perp(v2 a) { return (v2){a.y, -a.x}; }                        /* src/vectors.h:144 */
static inline float v2len(v2 a) { return sqrtf(a.x * a.x + a.y * a.y); }          /* src/vectors.h:93 */
static inline float v2dot(v2 a, v2 b) { return (a.x * b.x) + (a.y * b.y); }
static inline v2 v2norm(v2 a) { float m = sqrtf(a.x * a.x + a.y * a.y); return (v2){a.x / m, a.y / m}; } /* :139 */
static inline bool v2equ(v2 a, v2 b) { return EQUF(a.x, b.x) & EQUF(a.y, b.y); }  /* src/vectors.h:129 */

typedef struct { float xx, yx, xy, yy, x0, y0; } mat_t;
static const mat_t MAT_ID = {1, 0, 0, 1, 0, 0};

/* src/vkvg_matrix.c:193-206 */
static mat_t mat_mul(const mat_t *a, const mat_t *b) {
    mat_t r;
    r.xx = a->xx * b->xx + a->yx * b->xy;
    r.yx = a->xx * b->yx + a->yx * b->yy;
    r.xy = a->xy * b->xx + a->yy * b->xy;
    r.yy = a->xy * b->yx + a->yy * b->yy;
    r.x0 = a->x0 * b->xx + a->y0 * b->xy + b->x0;
    r.y0 = a->x0 * b->yx + a->y0 * b->yy + b->y0;
    return r;
}
/* src/vkvg_matrix.c:207-222 */
static void mat_distance(const mat_t *m, float *dx, float *dy) {
    float nx = (m->xx * *dx + m->xy * *dy);
    float ny = (m->yx * *dx + m->yy * *dy);
    *dx = nx;
    *dy = ny;
}
static void mat_point(const mat_t *m, float *x, float *y) {
    mat_distance(m, x, y);
    *x += m->x0;
    *y += m->y0;
}
/* src/vkvg_matrix.c:222-229 (double sqrt of a float argument, stored to float) */
static void mat_scale_of(const mat_t *m, float *sx, float *sy) {
    *sx = sqrt(m->xx * m->xx + m->xy * m->xy);
    *sy = sqrt(m->yx * m->yx + m->yy * m->yy);
}

enum { PAT_SOLID = 0, PAT_SURFACE = 1, PAT_LINEAR = 2, PAT_RADIAL = 3 };
typedef struct {
    float    colors[16][4];
    float    stops[16];
    float    cp[2][4];
    uint32_t count;
} grad_t; /* src/vkvg_pattern.h:38-47 (scalar block layout) */

typedef struct {
    const uint32_t *img; /* premultiplied RGBA8, R in the low byte */
    uint32_t        w, h, linear, extend; /* extend: vkvg_extend_t (NONE, REPEAT, REFLECT, PAD) */
    float           sx, sy;
    mat_t           minv;
} surf_src;
struct ovk_save { /* vkvg_context_save_t, src/vkvg_context_internal.h:101-125 */
    struct ovk_save *next;
    float    lineWidth, miterLimit, dashOffset, *dashes, opacity;
    uint32_t dashCount;
    int      lineCap, fillRule, clippingState;
    uint32_t curColor;
    int      patType;
    grad_t   grad;
    mat_t    mat, matInv;
    surf_src surf;
};

#define PIX(c, px, py) ((size_t)((uint32_t)(py) - (c)->wy0) * (c)->ww + ((uint32_t)(px) - (c)->wx0))
struct ovk_ctx {
    uint32_t W, H, S;       /* LOGICAL surface size: vertex stage, paint evaluation, scissor clamps */
    uint32_t wx0, wy0, ww, wh; /* stored window of it (ovk_create_window; the whole surface otherwise): only these pixels exist */
    int      status;
    /* surface */
    uint32_t *samples; /* H*W*S premultiplied RGBA8, R in low byte */
    uint8_t  *stencil; /* H*W*S */
    uint8_t  *resolved;
    bool      resolved_dirty;
    int32_t  *coverage; /* optional capture */
    int       capture;
    int       analytic; /* 1: analytic-coverage mode (one colour per pixel, exact area coverage), see analytic_draw */
    double   *area;     /* analytic mode: integral of the winding number over each pixel, last draw (W*H) */
    /* path storage, the reference's encoding (src/vkvg_context_internal.h:184-196) */
    v2       *points;
    uint32_t  pointCount, sizePoints;
    uint32_t *pathes;
    uint32_t  pathPtr, sizePathes, segmentPtr, subpathCount;
    bool      simpleConvex;
    /* state (defaults src/vkvg_context.c:24-61) */
    float    lineWidth, miterLimit, dashOffset, *dashes, opacity;
    uint32_t dashCount;
    int      lineCap, lineJoin, fillRule;
    uint32_t curColor;
    int      patType;
    grad_t   grad;
    mat_t    mat;
    /* output of the last tessellation */
    v2       *verts;
    uint32_t  vertCount, sizeVerts;
    uint32_t *inds;
    uint32_t  indCount, sizeInds;
    /* clip bookkeeping of the reference context (src/vkvg_context_internal.h:93-99, :123, :225-226) */
    mat_t            matInv;  /* pushConsts.matInv: recomputed on every CTM change (src/vkvg_context_internal.c:672-676) */
    surf_src         surf;    /* current surface source (patType 1) */
    int              op;      /* 0 OVER (and every operator without a pipeline of its own), 1 CLEAR, 2 DIFFERENCE */
    int              curClipState; /* 0 none, 1 clear, 2 clip (6 = clip_saved, only in saved entries) */
    uint32_t         curSavBit;
    struct ovk_save *saved;
    uint8_t        **spills; /* whole-stencil copies taken every 6 nested clip saves (vkvg_context.c:1268-1318) */
    uint32_t         nspills;
};

/* ------------------------------------------------------------------ */
/* sample positions (Vulkan standard sample locations, in 1/16 pixel)  */
/* ------------------------------------------------------------------ */
static const int8_t SP1[][2]  = {{8, 8}};
static const int8_t SP2[][2]  = {{12, 12}, {4, 4}};
static const int8_t SP4[][2]  = {{6, 2}, {14, 6}, {2, 10}, {10, 14}};
static const int8_t SP8[][2]  = {{9, 5}, {7, 11}, {13, 9}, {5, 3}, {3, 13}, {1, 7}, {11, 15}, {15, 1}};
static const int8_t SP16[][2] = {{9, 9}, {7, 5}, {5, 10}, {12, 7}, {3, 6}, {10, 13}, {13, 11}, {11, 3},
                                 {6, 14}, {8, 1}, {4, 2}, {2, 12}, {0, 8}, {15, 4}, {14, 15}, {1, 0}};
static const int8_t (*sample_table(uint32_t S))[2] {
    switch (S) {
    case 1: return SP1;
    case 2: return SP2;
    case 4: return SP4;
    case 8: return SP8;
    case 16: return SP16;
    }
    return NULL;
}
int ovk_sample_positions(uint32_t samples, int32_t *xy16) {
    const int8_t(*t)[2] = sample_table(samples);
    if (!t) return 0;
    for (uint32_t i = 0; i < samples; i++) { xy16[2 * i] = t[i][0]; xy16[2 * i + 1] = t[i][1]; }
    return 1;
}

/* ------------------------------------------------------------------ */
/* vertex stage + viewport + snap                                     */
/* ------------------------------------------------------------------ */
/* shaders/vkvg_main.vert:74-79: p = M*pos ; ndc = p*2/size - 1.  Viewport (0,0,W,H): win = ndc*W/2 + W/2.
 * Snap: 8 sub-pixel bits, round half up (the ICD's rule; defined here, see header). */
static inline int32_t snap_fixed(float w) { return (int32_t)floorf(w * 256.0f + 0.5f); }
static inline void vs_chain(const mat_t *m, float W, float H, float x, float y, int32_t *fx, int32_t *fy) {
    float px = m->xx * x + m->xy * y + m->x0;
    float py = m->yx * x + m->yy * y + m->y0;
    float nx = px * 2.0f / W - 1.0f;
    float ny = py * 2.0f / H - 1.0f;
    float wx = nx * (W * 0.5f) + (W * 0.5f);
    float wy = ny * (H * 0.5f) + (H * 0.5f);
    *fx = snap_fixed(wx);
    *fy = snap_fixed(wy);
}
void ovk_transform_snap(const float m[6], uint32_t width, uint32_t height, const float *xy, uint64_t n, int32_t *out) {
    mat_t M = {m[0], m[1], m[2], m[3], m[4], m[5]};
    for (uint64_t i = 0; i < n; i++)
        vs_chain(&M, (float)width, (float)height, xy[2 * i], xy[2 * i + 1], &out[2 * i], &out[2 * i + 1]);
}

/* ------------------------------------------------------------------ */
/* paint evaluation + blending                                        */
/* ------------------------------------------------------------------ */
static inline float clamp01(float x) { return x < 0.0f ? 0.0f : (x > 1.0f ? 1.0f : x); }
static inline float smoothstepf(float e0, float e1, float x) {
    float t = clamp01((x - e0) / (e1 - e0));
    return t * t * (3.0f - 2.0f * t);
}
static inline void mix4(float *c, const float *b, float t) {
    for (int k = 0; k < 4; k++) c[k] = c[k] * (1.0f - t) + b[k] * t;
}
/* A surface as paint: what the fragment shader receives (push constants source.xy / matInv, the bound image and its sampler:
 * src/vkvg_context_internal.c:705-773) and texture(source, uv) of shaders/vkvg_main.frag:72-82 under the Vulkan
 * texel-addressing rules (nearest / linear on unnormalised coordinates, REPEAT / MIRRORED_REPEAT / CLAMP_TO_EDGE /
 * CLAMP_TO_BORDER with a transparent-black border).  Like the rasterisation rules this lives in the ICD, not in the reference
 * tree: defined here from the specification. */
static const surf_src *g_surf_src; /* source of the draw being rasterised (set by cur_paint; the oracle is single threaded) */
static int tex_wrap(int i, int n, uint32_t mode, bool *border) {
    if (mode == 1) { i %= n; return i < 0 ? i + n : i; }
    if (mode == 2) {
        int m = i % (2 * n);
        if (m < 0) m += 2 * n;
        m -= n;
        return (n - 1) - (m >= 0 ? m : -(1 + m));
    }
    if (mode == 3) return i < 0 ? 0 : (i >= n ? n - 1 : i);
    *border = *border || i < 0 || i >= n;
    return i < 0 ? 0 : (i >= n ? n - 1 : i);
}
static void tex_fetch(const surf_src *s, int i, int j, float t[4]) {
    bool border = false;
    i = tex_wrap(i, (int)s->w, s->extend, &border);
    j = tex_wrap(j, (int)s->h, s->extend, &border);
    uint32_t p = border ? 0u : s->img[(size_t)j * s->w + i];
    for (int k = 0; k < 4; k++) t[k] = (float)((p >> (8 * k)) & 0xFF) / 255.0f;
}
static void sample_surface(const surf_src *s, float fx, float fy, float c[4]) {
    float px = fx - s->sx, py = fy - s->sy;
    float u = (s->minv.xx * px + s->minv.xy * py + s->minv.x0) / (float)s->w;
    float v = (s->minv.yx * px + s->minv.yy * py + s->minv.y0) / (float)s->h;
    float U = u * (float)s->w, V = v * (float)s->h;
    if (!(fabsf(U) < 1.0e9f) || !(fabsf(V) < 1.0e9f)) { c[0] = c[1] = c[2] = c[3] = 0.0f; return; }
    if (!s->linear) { tex_fetch(s, (int)floorf(U), (int)floorf(V), c); return; }
    U -= 0.5f; V -= 0.5f;
    float fi = floorf(U), fj = floorf(V);
    float al = U - fi, be = V - fj;
    float t00[4], t10[4], t01[4], t11[4];
    tex_fetch(s, (int)fi, (int)fj, t00);
    tex_fetch(s, (int)fi + 1, (int)fj, t10);
    tex_fetch(s, (int)fi, (int)fj + 1, t01);
    tex_fetch(s, (int)fi + 1, (int)fj + 1, t11);
    for (int k = 0; k < 4; k++)
        c[k] = ((1.0f - al) * (1.0f - be)) * t00[k] + (al * (1.0f - be)) * t10[k] + ((1.0f - al) * be) * t01[k] + (al * be) * t11[k];
}
/* shaders/vkvg_main.frag:68-157 for pattern types SOLID / LINEAR / RADIAL at fragment centre (fx,fy).
 * `src` is pc.source = (W, H, 0, 0) for gradients (src/vkvg_context_internal.c:783-787). */
static void eval_paint(int patType, const grad_t *g, float W, float H, uint32_t solid, float opacity, float fx, float fy,
                       float out[4]) {
    float c[4];
    if (patType == PAT_LINEAR) {
        float p0x = g->cp[0][0] / W, p0y = g->cp[0][1] / H;
        float p1x = g->cp[0][2] / W, p1y = g->cp[0][3] / H;
        float px = fx / W, py = fy / H;
        float dx = p1x - p0x, dy = p1y - p0y;
        float l  = sqrtf(dx * dx + dy * dy);
        float ux = dx / l, uy = dy / l;
        float dist;
        if (uy == 0.0f) {
            if (ux < 0.0f) dist = -(px - p0x) / l;
            else dist = (px - p0x) / l;
        } else {
            float m  = -ux / uy;
            float bb = p0y - m * p0x;
            dist     = ((py - m * px - bb) / sqrtf(1.0f + m * m)) / l;
            if (uy < 0.0f) dist = -dist;
        }
        for (int k = 0; k < 4; k++) c[k] = g->colors[0][k];
        mix4(c, g->colors[1], smoothstepf(g->stops[0], g->stops[1], dist));
        for (uint32_t i = 1; i + 1 < g->count; ++i) mix4(c, g->colors[i + 1], smoothstepf(g->stops[i], g->stops[i + 1], dist));
    } else if (patType == PAT_RADIAL) {
        float px = fx / W, py = fy / H;
        float c0x = g->cp[0][0] / W, c0y = g->cp[0][1] / H;
        float c1x = g->cp[1][0] / W, c1y = g->cp[1][1] / H;
        float r0 = g->cp[0][2] / W, r1 = g->cp[1][2] / W;
        float gradLength = 1.0f;
        float dfx = c0x - c1x, dfy = c0y - c1y;
        float rx = px - c0x, ry = py - c0y;
        float rl = sqrtf(rx * rx + ry * ry);
        float rdx = rx / rl, rdy = ry / rl;
        float a    = rdx * rdx + rdy * rdy;
        float b    = 2.0f * (rdx * dfx + rdy * dfy);
        float cc   = (dfx * dfx + dfy * dfy) - r1 * r1;
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
    } else if (patType == 1 /* SURFACE */) {
        if (g_surf_src) sample_surface(g_surf_src, fx, fy, c);
        else c[0] = c[1] = c[2] = c[3] = 0.0f;
    } else { /* SOLID: vertex colour, R8G8B8A8_UNORM attribute (src/vkvg_device_internal.c:282-284) */
        c[0] = (float)(solid & 0xFF) / 255.0f;
        c[1] = (float)((solid >> 8) & 0xFF) / 255.0f;
        c[2] = (float)((solid >> 16) & 0xFF) / 255.0f;
        c[3] = (float)((solid >> 24) & 0xFF) / 255.0f;
    }
    for (int k = 0; k < 4; k++) out[k] = c[k] * opacity; /* frag:152 (VKVG_PREMULT_ALPHA) */
}
static inline uint32_t unorm8(float v) {
    float q = v * 255.0f + 0.5f;
    if (!(q > 0.0f)) return 0;
    if (q >= 255.5f) return 255;
    return (uint32_t)q;
}
/* premultiplied OVER, src/vkvg_device_internal.c:203-209: dst = src*ONE + dst*(1-src.a), RGB and A alike */
static int g_blend_op; /* operator of the draw being rasterised (set with the paint; single threaded like g_surf_src):
                        * 0 pipe_OVER, 1 pipe_CLEAR (logic op CLEAR: zeros), 2 pipe_SUB (blend op SUBTRACT, same factors):
                        * src/vkvg_device_internal.c:358-373 */
static inline uint32_t blend_general(uint32_t dst, const float s[4], float ia) { /* r = src + dst * ia per channel, UNORM8 store */
    uint32_t out = 0;
    for (int k = 0; k < 4; k++) {
        float d = (float)((dst >> (8 * k)) & 0xFF) / 255.0f;
        float r = s[k] + d * ia;
        out |= unorm8(r) << (8 * k);
    }
    return out;
}
static inline uint32_t blend_over(uint32_t dst, const float s[4]) {
    uint32_t out = 0;
    float    ia  = 1.0f - s[3];
    if (g_blend_op == 1) return 0;
    if (g_blend_op == 2) ia = -ia; /* src - dst * (1 - src.a); the UNORM store clamps at 0 */
    for (int k = 0; k < 4; k++) {
        float d = (float)((dst >> (8 * k)) & 0xFF) / 255.0f;
        float r = s[k] + d * ia;
        out |= unorm8(r) << (8 * k);
    }
    return out;
}

/* ------------------------------------------------------------------ */
/* rasteriser core                                                    */
/* ------------------------------------------------------------------ */
typedef struct {
    int      patType;
    grad_t   grad;
    uint32_t solid;
    float    opacity;
    int      op;
} paint_t;

typedef struct { int32_t x0, y0, x1, y1; } rect_i; /* pixel scissor, half open */

static inline bool tri_edge_in(int64_t E, int64_t dx, int64_t dy) {
    /* top-left rule for triangles normalised to cross > 0 in y-down window space */
    return E > 0 || (E == 0 && (dy < 0 || (dy == 0 && dx > 0)));
}
typedef void (*sample_fn)(ovk_ctx *c, uint32_t px, uint32_t py, uint32_t s, void *user);

/* visits every sample of every pixel covered by the triangle, pixel-major; calls `pix` once per pixel with mask */
typedef void (*pixel_fn)(ovk_ctx *c, uint32_t px, uint32_t py, uint32_t mask, void *user);
static void raster_tri(ovk_ctx *c, const int32_t v[6], rect_i sc, pixel_fn fn, void *user) {
    int64_t x0 = v[0], y0 = v[1], x1 = v[2], y1 = v[3], x2 = v[4], y2 = v[5];
    int64_t area = (x1 - x0) * (y2 - y0) - (x2 - x0) * (y1 - y0);
    if (area == 0) return;
    if (area < 0) { int64_t t; t = x1; x1 = x2; x2 = t; t = y1; y1 = y2; y2 = t; }
    int64_t minx = x0 < x1 ? (x0 < x2 ? x0 : x2) : (x1 < x2 ? x1 : x2);
    int64_t maxx = x0 > x1 ? (x0 > x2 ? x0 : x2) : (x1 > x2 ? x1 : x2);
    int64_t miny = y0 < y1 ? (y0 < y2 ? y0 : y2) : (y1 < y2 ? y1 : y2);
    int64_t maxy = y0 > y1 ? (y0 > y2 ? y0 : y2) : (y1 > y2 ? y1 : y2);
    int64_t pxa = minx >> 8, pxb = (maxx >> 8) + 1, pya = miny >> 8, pyb = (maxy >> 8) + 1;
    if (pxa < sc.x0) pxa = sc.x0;
    if (pya < sc.y0) pya = sc.y0;
    if (pxb > sc.x1) pxb = sc.x1;
    if (pyb > sc.y1) pyb = sc.y1;
    const int8_t(*sp)[2] = sample_table(c->S);
    int64_t dx01 = x1 - x0, dy01 = y1 - y0, dx12 = x2 - x1, dy12 = y2 - y1, dx20 = x0 - x2, dy20 = y0 - y2;
    for (int64_t py = pya; py < pyb; py++)
        for (int64_t px = pxa; px < pxb; px++) {
            uint32_t mask = 0;
            for (uint32_t s = 0; s < c->S; s++) {
                int64_t sx = px * 256 + sp[s][0] * 16, sy = py * 256 + sp[s][1] * 16;
                int64_t E0 = dx01 * (sy - y0) - dy01 * (sx - x0);
                int64_t E1 = dx12 * (sy - y1) - dy12 * (sx - x1);
                int64_t E2 = dx20 * (sy - y2) - dy20 * (sx - x2);
                if (tri_edge_in(E0, dx01, dy01) && tri_edge_in(E1, dx12, dy12) && tri_edge_in(E2, dx20, dy20)) mask |= 1u << s;
            }
            if (mask) fn(c, (uint32_t)px, (uint32_t)py, mask, user);
        }
}

/* winding of one sample against one directed edge: the definition shared with the CUDA rasteriser.
 * The sample is treated as displaced by (+eps', +eps), eps << eps' (equivalent to the top-left rule). */
static inline int edge_winding(int64_t ax, int64_t ay, int64_t bx, int64_t by, int64_t sx, int64_t sy) {
    if ((ay <= sy) == (by <= sy)) return 0;
    int64_t dx = bx - ax, dy = by - ay;
    int64_t E  = dx * (sy - ay) - dy * (sx - ax);
    if (dy > 0) return E <= 0 ? 1 : 0;
    return E >= 0 ? -1 : 0;
}
void ovk_winding_brute(const int32_t *e, uint64_t n, uint32_t W, uint32_t H, uint32_t S, int32_t *out) {
    const int8_t(*sp)[2] = sample_table(S);
    memset(out, 0, (size_t)W * H * S * sizeof(int32_t));
    for (uint64_t i = 0; i < n; i++) {
        int64_t ax = e[4 * i], ay = e[4 * i + 1], bx = e[4 * i + 2], by = e[4 * i + 3];
        if (ay == by) continue;
        int64_t miny = ay < by ? ay : by, maxy = ay > by ? ay : by;
        int64_t minx = ax < bx ? ax : bx;
        int64_t pya = miny >> 8, pyb = (maxy >> 8) + 1, pxa = minx >> 8;
        if (pya < 0) pya = 0;
        if (pyb > (int64_t)H) pyb = H;
        if (pxa < 0) pxa = 0;
        for (int64_t py = pya; py < pyb; py++)
            for (int64_t px = pxa; px < (int64_t)W; px++)
                for (uint32_t s = 0; s < S; s++)
                    out[((size_t)py * W + px) * S + s] +=
                        edge_winding(ax, ay, bx, by, px * 256 + sp[s][0] * 16, py * 256 + sp[s][1] * 16);
    }
}

/* ------------------------------------------------------------------ */
/* analytic-coverage mode (north_star: "with an analytic-coverage mode alongside"; no counterpart in the      */
/* reference, whose coverage is the ICD's MSAA).  Definition, shared with the CUDA fine pass:                 */
/*   A(px,py) = integral over the pixel square of the winding number W(x,y) of the draw's directed edges      */
/*            = sum_e sign(dy_e) * integral_{y in span_e, py <= y <= py+1} clamp(px + 1 - x_e(y), 0, 1) dy     */
/*   coverage = min(|A|, 1) for NON_ZERO and for strokes (union of the triangles), 1 - |A mod 2 - 1| for       */
/*   EVEN_ODD; the pixel (one colour, no samples) is blended once with the paint scaled by the coverage.      */
/* Evaluated here edge by edge in double precision over the whole surface (no tiles, no backdrop), i.e. by a  */
/* different decomposition than the CUDA path's backdrop + V + H split.                                       */
/* ------------------------------------------------------------------ */
static inline double clampd(double x, double a, double b) { return x < a ? a : (x > b ? b : x); }
/* mean over t in [0,1] of clamp(lo + (hi - lo) t, 0, 1), lo <= hi */
static inline double mean_clamp01(double lo, double hi) {
    double d = hi - lo;
    if (d < 1e-300) return clampd(lo, 0.0, 1.0);
    double t0 = clampd(-lo / d, 0.0, 1.0), t1 = clampd((hi - 1.0) / d, 0.0, 1.0);
    double clo = lo > 0.0 ? lo : 0.0, chi = hi < 1.0 ? hi : 1.0;
    return t1 + (1.0 - t0 - t1) * 0.5 * (clo + chi);
}
void ovk_area_brute(const int32_t *e, uint64_t n, uint32_t W, uint32_t H, double *out) {
    size_t  stride = (size_t)W + 1;
    double *acc    = (double *)calloc(stride * H, sizeof(double)); /* per-row deltas: prefix sum along x gives A */
    for (uint64_t i = 0; i < n; i++) {
        double ax = e[4 * i] / 256.0, ay = e[4 * i + 1] / 256.0, bx = e[4 * i + 2] / 256.0, by = e[4 * i + 3] / 256.0;
        if (ay == by) continue;
        double sgn = by > ay ? 1.0 : -1.0;
        double xt = by > ay ? ax : bx, yt = by > ay ? ay : by, xb = by > ay ? bx : ax, yb = by > ay ? by : ay;
        double slope = (xb - xt) / (yb - yt);
        int64_t r0 = (int64_t)floor(yt), r1 = (int64_t)ceil(yb) - 1;
        if (r0 < 0) r0 = 0;
        if (r1 > (int64_t)H - 1) r1 = (int64_t)H - 1;
        for (int64_t r = r0; r <= r1; r++) {
            double ys = yt > (double)r ? yt : (double)r, ye = yb < (double)(r + 1) ? yb : (double)(r + 1);
            if (ye <= ys) continue;
            double xs = xt + (ys - yt) * slope, xe = xt + (ye - yt) * slope, h = ye - ys;
            double xmin = xs < xe ? xs : xe, xmax = xs < xe ? xe : xs;
            double *row = acc + (size_t)r * stride;
            if (xmin >= (double)W) continue;
            int64_t c0 = (int64_t)floor(xmin), c1 = (int64_t)floor(xmax);
            if (c0 < 0) c0 = 0;
            if (c1 > (int64_t)W - 1) c1 = (int64_t)W - 1;
            double prev = 0.0;
            for (int64_t c = c0; c <= c1; c++) {
                double a = h * mean_clamp01((double)(c + 1) - xmax, (double)(c + 1) - xmin);
                row[c] += sgn * (a - prev);
                prev = a;
            }
            /* every pixel right of the last one the edge touches is covered over the whole height h
             * (an edge wholly left of the surface, c1 < 0, covers from column 0) */
            int64_t cn = c1 < c0 ? c0 : c1 + 1;
            if (cn < (int64_t)W) row[cn] += sgn * (h - prev);
        }
    }
    for (uint32_t y = 0; y < H; y++) {
        double run = 0.0;
        for (uint32_t x = 0; x < W; x++) {
            run += acc[(size_t)y * stride + x];
            out[(size_t)y * W + x] = run;
        }
    }
    free(acc);
}
static inline float analytic_coverage(double A, int rule) {
    if (rule == OVK_RULE_EVEN_ODD) {
        double t = A - 2.0 * floor(A * 0.5);
        return (float)(1.0 - fabs(t - 1.0));
    }
    double a = fabs(A);
    return (float)(a < 1.0 ? a : 1.0);
}
/* one draw in analytic mode: edges -> A -> coverage -> paint * coverage OVER the single colour of the pixel */
static void analytic_draw(ovk_ctx *c, const int32_t *e, uint64_t n, int rule, int patType, const grad_t *grad, uint32_t solid, float opacity) {
    if (!c->area) c->area = (double *)calloc((size_t)c->ww * c->wh, sizeof(double));
    ovk_area_brute(e, n, c->W, c->H, c->area);
    for (uint32_t py = 0; py < c->H; py++)
        for (uint32_t px = 0; px < c->W; px++) {
            float cov = analytic_coverage(c->area[PIX(c, px, py)], rule);
            if (!(cov > 0.0f)) continue;
            if (c->stencil[PIX(c, px, py) * c->S] & 0x2) continue; /* clipped out (analytic mode: one flag per pixel) */
            float col[4], s[4];
            eval_paint(patType, grad, (float)c->W, (float)c->H, solid, opacity, (float)px + 0.5f, (float)py + 0.5f, col);
            for (int k = 0; k < 4; k++) s[k] = col[k] * cov;
            size_t base = PIX(c, px, py) * c->S;
            if (g_blend_op == 1) { /* analytic CLEAR: the covered part of the pixel is wiped */
                const float z[4] = {0, 0, 0, 0};
                for (uint32_t q = 0; q < c->S; q++) c->samples[base + q] = blend_general(c->samples[base + q], z, 1.0f - cov);
                continue;
            }
            for (uint32_t q = 0; q < c->S; q++) c->samples[base + q] = blend_over(c->samples[base + q], s);
        }
    c->resolved_dirty = true;
}
void ovk_set_coverage_mode(ovk_ctx *c, int analytic) { c->analytic = analytic; }
const double *ovk_last_area(ovk_ctx *c) { return c->area; }

static void cov_add(ovk_ctx *c, uint32_t px, uint32_t py, uint32_t s, int32_t v) {
    if (c->capture) c->coverage[PIX(c, px, py) * c->S + s] += v;
}
static void cov_reset(ovk_ctx *c) {
    if (c->capture) memset(c->coverage, 0, (size_t)c->ww * c->wh * c->S * sizeof(int32_t));
}

/* --- colour draw: triangle blended directly with pipe_OVER, stencil compare mask = CLIP
 *     (src/vkvg_device_internal.c:240-246, src/vkvg_context_internal.c:667) --- */
typedef struct { const paint_t *p; uint32_t cmpMask; } blend_user;
static void px_blend(ovk_ctx *c, uint32_t px, uint32_t py, uint32_t mask, void *user) {
    blend_user *u = (blend_user *)user;
    float col[4];
    eval_paint(u->p->patType, &u->p->grad, (float)c->W, (float)c->H, u->p->solid, u->p->opacity, (float)px + 0.5f,
               (float)py + 0.5f, col);
    size_t base = PIX(c, px, py) * c->S;
    for (uint32_t s = 0; s < c->S; s++)
        if (mask & (1u << s)) {
            uint8_t st = c->stencil[base + s];
            if ((st & u->cmpMask) != (0x1 & u->cmpMask)) continue; /* compare EQUAL, static reference 0x1 */
            c->stencil[base + s] = st & ~0x1;                        /* passOp ZERO under write mask FILL */
            c->samples[base + s] = blend_over(c->samples[base + s], col);
            cov_add(c, px, py, s, 1);
        }
    c->resolved_dirty = true;
}
/* --- stencil fan: pipelinePolyFill, src/vkvg_device_internal.c:226-232 --- */
static void px_invert(ovk_ctx *c, uint32_t px, uint32_t py, uint32_t mask, void *user) {
    size_t base = PIX(c, px, py) * c->S;
    for (uint32_t s = 0; s < c->S; s++)
        if (mask & (1u << s)) {
            uint8_t st = c->stencil[base + s];
            if ((st & 0x2) != 0) continue; /* compare mask CLIP, reference 0 */
            c->stencil[base + s] = st ^ 0x1;
        }
}
static rect_i full_rect(ovk_ctx *c) { return (rect_i){(int32_t)c->wx0, (int32_t)c->wy0, (int32_t)(c->wx0 + c->ww), (int32_t)(c->wy0 + c->wh)}; }
static rect_i clip_rect(ovk_ctx *c, int64_t x, int64_t y, int64_t w, int64_t h) {
    rect_i r;
    int64_t x1 = x + w, y1 = y + h;
    r.x0 = (int32_t)(x < 0 ? 0 : (x > c->W ? c->W : x));
    r.y0 = (int32_t)(y < 0 ? 0 : (y > c->H ? c->H : y));
    r.x1 = (int32_t)(x1 < 0 ? 0 : (x1 > c->W ? c->W : x1));
    r.y1 = (int32_t)(y1 < 0 ? 0 : (y1 > c->H ? c->H : y1));
    /* (window mode) only the stored part of the surface exists */
    const rect_i win = full_rect(c);
    if (r.x0 < win.x0) r.x0 = win.x0;
    if (r.y0 < win.y0) r.y0 = win.y0;
    if (r.x1 > win.x1) r.x1 = win.x1;
    if (r.y1 > win.y1) r.y1 = win.y1;
    return r;
}
/* full-screen cover: every sample inside the scissor takes the stencil test */
static void cover_rect(ovk_ctx *c, rect_i sc, const paint_t *p, uint32_t cmpMask) {
    blend_user u = {p, cmpMask};
    for (int32_t py = sc.y0; py < sc.y1; py++)
        for (int32_t px = sc.x0; px < sc.x1; px++) {
            /* only run the paint when some sample passes, to keep the oracle usable on big surfaces */
            size_t base = PIX(c, px, py) * c->S;
            bool any = false;
            for (uint32_t s = 0; s < c->S; s++)
                if ((c->stencil[base + s] & cmpMask) == (0x1 & cmpMask)) any = true;
            if (any) px_blend(c, (uint32_t)px, (uint32_t)py, (1u << c->S) - 1, &u);
        }
}

static void resolve(ovk_ctx *c) {
    if (!c->resolved_dirty) return;
    size_t n = (size_t)c->ww * c->wh;
    for (size_t i = 0; i < n; i++)
        for (int k = 0; k < 4; k++) {
            uint32_t sum = 0;
            for (uint32_t s = 0; s < c->S; s++) sum += (c->samples[i * c->S + s] >> (8 * k)) & 0xFF;
            c->resolved[4 * i + k] = (uint8_t)((sum + c->S / 2) / c->S);
        }
    c->resolved_dirty = false;
}

/* ------------------------------------------------------------------ */
/* context                                                            */
/* ------------------------------------------------------------------ */
/* A W x H surface of which only the window [x0, x0 + w) x [y0, y0 + h) is stored and rasterised: vertex stage, paint evaluation and
 * scissors use the logical size, so the window holds exactly the pixels the same region of the whole surface would (test infrastructure
 * for the BASELINE configs whose full surfaces - 8192^2, 16384^2 at 4 samples - the scalar rasteriser cannot hold or finish).
 * MSAA mode only (the analytic-coverage restatement integrates over the whole surface). */
ovk_ctx *ovk_create_window(uint32_t W, uint32_t H, uint32_t S, uint32_t x0, uint32_t y0, uint32_t w, uint32_t h) {
    if (!sample_table(S) || x0 + w > W || y0 + h > H || !w || !h) return NULL;
    ovk_ctx *c = (ovk_ctx *)calloc(1, sizeof(ovk_ctx));
    c->W = W; c->H = H; c->S = S;
    c->wx0 = x0; c->wy0 = y0; c->ww = w; c->wh = h;
    c->samples  = (uint32_t *)calloc((size_t)w * h * S, 4);
    c->stencil  = (uint8_t *)calloc((size_t)w * h * S, 1);
    c->resolved = (uint8_t *)calloc((size_t)w * h, 4);
    c->sizePoints = 1024; c->points = (v2 *)malloc(c->sizePoints * sizeof(v2));
    c->sizePathes = 64; c->pathes = (uint32_t *)calloc(c->sizePathes, 4);
    c->sizeVerts = 4096; c->verts = (v2 *)malloc(c->sizeVerts * sizeof(v2));
    c->sizeInds = 4096 * 6; c->inds = (uint32_t *)malloc(c->sizeInds * 4);
    /* src/vkvg_context.c:24-61 */
    c->lineWidth = 1.f; c->miterLimit = 10.f; c->fillRule = OVK_FILL_NON_ZERO;
    c->lineCap = OVK_CAP_BUTT; c->lineJoin = OVK_JOIN_MITER;
    c->curColor = 0xff000000; c->patType = PAT_SOLID; c->opacity = 1.0f; c->mat = MAT_ID; c->matInv = MAT_ID;
    return c;
}
ovk_ctx *ovk_create(uint32_t W, uint32_t H, uint32_t S) { return ovk_create_window(W, H, S, 0, 0, W, H); }
void ovk_destroy(ovk_ctx *c) {
    if (!c) return;
    free(c->samples); free(c->stencil); free(c->resolved); free(c->coverage); free(c->area);
    free(c->points); free(c->pathes); free(c->verts); free(c->inds); free(c->dashes);
    while (c->saved) { struct ovk_save *n = c->saved->next; free(c->saved->dashes); free(c->saved); c->saved = n; }
    for (uint32_t i = 0; i < c->nspills; i++) free(c->spills[i]);
    free(c->spills);
    free(c);
}
int  ovk_status(ovk_ctx *c) { return c->status; }
void ovk_clear(ovk_ctx *c) { /* src/vkvg_context.c:734-753: colour + stencil cleared */
    memset(c->samples, 0, (size_t)c->ww * c->wh * c->S * 4);
    memset(c->stencil, 0, (size_t)c->ww * c->wh * c->S);
    c->resolved_dirty = true;
}
void ovk_set_capture_coverage(ovk_ctx *c, int on) {
    c->capture = on;
    if (on && !c->coverage) c->coverage = (int32_t *)calloc((size_t)c->ww * c->wh * c->S, sizeof(int32_t));
}
const int32_t *ovk_last_coverage(ovk_ctx *c) { return c->coverage; }
const uint8_t *ovk_pixels(ovk_ctx *c) { resolve(c); return c->resolved; }
const uint8_t *ovk_sample_pixels(ovk_ctx *c) { return (const uint8_t *)c->samples; }
/* src/vkvg_surface.c:371-382: un-premultiply in double with truncating casts; alpha 0 -> 0 */
void ovk_write_to_memory(ovk_ctx *c, uint8_t *out) {
    resolve(c);
    size_t n = (size_t)c->ww * c->wh;
    for (size_t i = 0; i < n; i++) {
        const uint8_t *p = c->resolved + 4 * i;
        double alpha = (double)p[3] / 255.f;
        for (int k = 0; k < 3; k++) {
            double v = (double)p[k] / alpha;
            out[4 * i + k] = (v == v && v < 2147483648.0) ? (uint8_t)(int32_t)v : 0; /* NaN/inf -> 0 as cvttsd2si does */
        }
        out[4 * i + 3] = p[3];
    }
}

/* ------------------------------------------------------------------ */
/* path storage — src/vkvg_context_internal.c:42-238                  */
/* ------------------------------------------------------------------ */
static void check_pathes(ovk_ctx *c) {
    if (c->pathPtr + c->segmentPtr + 8 >= c->sizePathes) {
        uint32_t old = c->sizePathes;
        c->sizePathes *= 2;
        c->pathes = (uint32_t *)realloc(c->pathes, c->sizePathes * 4);
        memset(c->pathes + old, 0, (c->sizePathes - old) * 4);
    }
}
static bool path_empty(ovk_ctx *c) { return c->pathes[c->pathPtr] == 0; }              /* :132 */
static v2   cur_pos(ovk_ctx *c) { return c->points[c->pointCount - 1]; }               /* :134 */
static void add_point(ovk_ctx *c, float x, float y) {                                  /* :221-238 */
    if (c->pointCount + 1 >= c->sizePoints) {
        c->sizePoints *= 2;
        c->points = (v2 *)realloc(c->points, c->sizePoints * sizeof(v2));
    }
    if (isnan(x) || isnan(y)) return;
    c->points[c->pointCount] = (v2){x, y};
    c->pointCount++;
    c->pathes[c->pathPtr]++;
    if (c->segmentPtr > 0) c->pathes[c->pathPtr + c->segmentPtr]++;
}
static void set_curve_start(ovk_ctx *c) {                                              /* :136-151 */
    if (c->segmentPtr > 0) {
        if ((c->pathes[c->pathPtr + c->segmentPtr] & P_MASK) > 0) c->segmentPtr++;
    } else {
        if (c->pathes[c->pathPtr] > 0) {
            c->pathes[c->pathPtr + 1] = c->pathes[c->pathPtr];
            c->segmentPtr             = 2;
        } else
            c->segmentPtr = 1;
    }
    check_pathes(c);
    c->pathes[c->pathPtr + c->segmentPtr] = 0;
}
static void set_curve_end(ovk_ctx *c) {                                                /* :153-160 */
    c->pathes[c->pathPtr + c->segmentPtr] |= P_CURVES;
    c->segmentPtr++;
    check_pathes(c);
    c->pathes[c->pathPtr + c->segmentPtr] = 0;
}
static void finish_path(ovk_ctx *c) {                                                  /* :163-197 */
    if (c->pathes[c->pathPtr] == 0) return;
    if ((c->pathes[c->pathPtr] & P_MASK) < 2) {
        c->pointCount -= c->pathes[c->pathPtr];
        c->pathes[c->pathPtr] = 0;
        c->segmentPtr         = 0;
        return;
    }
    if (c->pathPtr == 0 && c->simpleConvex) c->pathes[0] |= P_CONVEX;
    if (c->segmentPtr > 0) {
        c->pathes[c->pathPtr] |= P_CURVES;
        if ((c->pathes[c->pathPtr + c->segmentPtr] & P_CURVES) == 0 && (c->pathes[c->pathPtr + c->segmentPtr] & P_MASK) > 0)
            c->segmentPtr++;
        c->pathPtr += c->segmentPtr;
    } else
        c->pathPtr++;
    check_pathes(c);
    c->pathes[c->pathPtr] = 0;
    c->segmentPtr         = 0;
    c->subpathCount++;
    c->simpleConvex = false;
}
static void clear_path(ovk_ctx *c) {                                                   /* :199-206 */
    c->pathPtr = 0; c->pathes[0] = 0; c->pointCount = 0; c->segmentPtr = 0; c->subpathCount = 0; c->simpleConvex = false;
}
static void remove_last_point(ovk_ctx *c) {                                            /* :207-219 */
    c->pathes[c->pathPtr]--;
    c->pointCount--;
    if (c->segmentPtr > 0) {
        if (!c->pathes[c->pathPtr + c->segmentPtr]) c->segmentPtr--;
        c->pathes[c->pathPtr + c->segmentPtr]--;
        if ((c->pathes[c->pathPtr + c->segmentPtr] & P_MASK) == 0) c->pathes[c->pathPtr + c->segmentPtr] = 0;
        else if (c->pathes[c->pathPtr + c->segmentPtr] & P_CURVES) c->segmentPtr++;
    }
}
static void line_to_(ovk_ctx *c, float x, float y) {                                   /* :1463-1472 */
    v2 p = {x, y};
    if (!path_empty(c) && v2equ(cur_pos(c), p)) return;
    add_point(c, x, y);
    c->simpleConvex = false;
}
static float arc_step(ovk_ctx *c, float radius) {                                      /* :245-252 */
    float sx, sy;
    mat_scale_of(&c->mat, &sx, &sy);
    float r = radius * fabsf(fmaxf(sx, sy));
    if (r < 30.0f) return fminf(PIF / 3.f, PIF / r);
    return fminf(PIF / 3.f, PIF / (r * 0.4f));
}

void ovk_new_path(ovk_ctx *c) { clear_path(c); }                                       /* vkvg_context.c:341-349 */
void ovk_new_sub_path(ovk_ctx *c) { finish_path(c); }                                  /* :332-340 */
void ovk_close_path(ovk_ctx *c) {                                                      /* :350-373 */
    if (c->pathes[c->pathPtr] & P_CLOSED) return;
    if (c->pathes[c->pathPtr] < 3) return;
    if (v2equ(c->points[c->pointCount - 1], c->points[c->pointCount - c->pathes[c->pathPtr]])) {
        if (c->pathes[c->pathPtr] < 4) return;
        remove_last_point(c);
    }
    c->pathes[c->pathPtr] |= P_CLOSED;
    finish_path(c);
}
void ovk_move_to(ovk_ctx *c, float x, float y) { finish_path(c); add_point(c, x, y); } /* :515-522 */
void ovk_line_to(ovk_ctx *c, float x, float y) { line_to_(c, x, y); }                  /* :386-393 */
void ovk_rel_line_to(ovk_ctx *c, float dx, float dy) {                                 /* :374-385 */
    if (path_empty(c)) add_point(c, 0, 0);
    v2 cp = cur_pos(c);
    line_to_(c, cp.x + dx, cp.y + dy);
}
void ovk_rel_move_to(ovk_ctx *c, float x, float y) {                                   /* :504-514 */
    if (path_empty(c)) add_point(c, 0, 0);
    v2 cp = cur_pos(c);
    finish_path(c);
    add_point(c, cp.x + x, cp.y + y);
}
void ovk_get_current_point(ovk_ctx *c, float *x, float *y) {                           /* :528-540 */
    if (path_empty(c)) { *x = *y = 0; return; }
    v2 cp = cur_pos(c);
    *x = cp.x; *y = cp.y;
}
int ovk_rectangle(ovk_ctx *c, float x, float y, float w, float h) {                    /* :620-639 */
    finish_path(c);
    if (w <= 0 || h <= 0) return 14; /* VKVG_STATUS_INVALID_RECT */
    add_point(c, x, y);
    add_point(c, x + w, y);
    add_point(c, x + w, y + h);
    add_point(c, x, y + h);
    c->pathes[c->pathPtr] |= (P_CLOSED | P_CONVEX);
    finish_path(c);
    return 0;
}

/* arcs — src/vkvg_context.c:394-503 */
static void arc_impl(ovk_ctx *c, float xc, float yc, float radius, float a1, float a2, bool negative) {
    if (!negative) {
        while (a2 < a1) a2 += 2.f * PIF;
        if (a2 - a1 > 2.f * PIF) a2 = a1 + 2.f * PIF;
    } else {
        while (a2 > a1) a2 -= 2.f * PIF;
        if (a1 - a2 > a1 + 2.f * PIF) a2 = a1 - 2.f * PIF; /* sic, :457 */
    }
    v2    v    = {cosf(a1) * radius + xc, sinf(a1) * radius + yc};
    float step = arc_step(c, radius);
    float a    = a1;
    if (path_empty(c)) {
        set_curve_start(c);
        add_point(c, v.x, v.y);
        c->simpleConvex = !c->pathPtr;
    } else {
        line_to_(c, v.x, v.y);
        set_curve_start(c);
        c->simpleConvex = false;
    }
    if (!negative) a += step; else a -= step;
    if (EQUF(a2, a1)) return;
    if (!negative)
        while (a < a2) { add_point(c, cosf(a) * radius + xc, sinf(a) * radius + yc); a += step; }
    else
        while (a > a2) { add_point(c, cosf(a) * radius + xc, sinf(a) * radius + yc); a -= step; }
    if (EQUF(negative ? a1 - a2 : a2 - a1, PIF * 2.f)) {
        set_curve_end(c);
        ovk_close_path(c);
        return;
    }
    a = a2;
    add_point(c, cosf(a) * radius + xc, sinf(a) * radius + yc);
    set_curve_end(c);
}
void ovk_arc(ovk_ctx *c, float xc, float yc, float r, float a1, float a2) { arc_impl(c, xc, yc, r, a1, a2, false); }
void ovk_arc_negative(ovk_ctx *c, float xc, float yc, float r, float a1, float a2) { arc_impl(c, xc, yc, r, a1, a2, true); }

/* ------------------------------------------------------------------ */
/* cubic flattening — src/vkvg_context_internal.c:1305-1461           */
/* ------------------------------------------------------------------ */
typedef struct { ovk_ctx *c; float *out; uint32_t n, cap; } flat_sink;
static void sink_pt(flat_sink *s, float x, float y) {
    if (s->c) { add_point(s->c, x, y); return; }
    if (isnan(x) || isnan(y)) return;
    if (s->out && s->n < s->cap) { s->out[2 * s->n] = x; s->out[2 * s->n + 1] = y; }
    s->n++;
}
static void rec_bezier(flat_sink *s, float tol, float x1, float y1, float x2, float y2, float x3, float y3, float x4, float y4,
                       unsigned level) {
    if (level > 100) return; /* CURVE_RECURSION_LIMIT */
    float x12 = (x1 + x2) / 2, y12 = (y1 + y2) / 2, x23 = (x2 + x3) / 2, y23 = (y2 + y3) / 2;
    float x34 = (x3 + x4) / 2, y34 = (y3 + y4) / 2;
    float x123 = (x12 + x23) / 2, y123 = (y12 + y23) / 2, x234 = (x23 + x34) / 2, y234 = (y23 + y34) / 2;
    float x1234 = (x123 + x234) / 2, y1234 = (y123 + y234) / 2;
    if (level > 0) { /* first level always subdivides, :1334 */
        float dx = x4 - x1, dy = y4 - y1;
        float d2 = fabsf(((x2 - x4) * dy - (y2 - y4) * dx));
        float d3 = fabsf(((x3 - x4) * dy - (y3 - y4) * dx));
        float da1, da2;
        /* the thresholds below are double literals in the reference: 1.7, 0.01 */
        if (d2 > 1.7 && d3 > 1.7) {
            if ((d2 + d3) * (d2 + d3) <= (dx * dx + dy * dy) * tol) {
                float a23 = atan2f(y3 - y2, x3 - x2);
                da1 = fabsf(a23 - atan2f(y2 - y1, x2 - x1));
                da2 = fabsf(atan2f(y4 - y3, x4 - x3) - a23);
                if (da1 >= PIF) da1 = TWO_OVER_PIF - da1; /* the 2/pi quirk, :1363-1366 */
                if (da2 >= PIF) da2 = TWO_OVER_PIF - da2;
                if (da1 + da2 < (float)0.01) { sink_pt(s, x1234, y1234); return; }
                if (da1 > 0.01) { sink_pt(s, x2, y2); return; }
                if (da2 > 0.01) { sink_pt(s, x3, y3); return; }
            }
        } else {
            if (d2 > 1.7) {
                if (d2 * d2 <= tol * (dx * dx + dy * dy)) {
                    da1 = fabsf(atan2f(y3 - y2, x3 - x2) - atan2f(y2 - y1, x2 - x1));
                    if (da1 >= PIF) da1 = TWO_OVER_PIF - da1;
                    if (da1 < 0.01) { sink_pt(s, x2, y2); sink_pt(s, x3, y3); return; }
                    if (da1 > 0.01) { sink_pt(s, x2, y2); return; }
                }
            } else if (d3 > 1.7) {
                if (d3 * d3 <= tol * (dx * dx + dy * dy)) {
                    da1 = fabsf(atan2f(y4 - y3, x4 - x3) - atan2f(y3 - y2, x3 - x2));
                    if (da1 >= PIF) da1 = TWO_OVER_PIF - da1;
                    if (da1 < 0.01) { sink_pt(s, x2, y2); sink_pt(s, x3, y3); return; }
                    if (da1 > 0.01) { sink_pt(s, x3, y3); return; }
                }
            } else {
                dx = x1234 - (x1 + x4) / 2;
                dy = y1234 - (y1 + y4) / 2;
                if (dx * dx + dy * dy <= tol) { sink_pt(s, x1234, y1234); return; }
            }
        }
    }
    rec_bezier(s, tol, x1, y1, x12, y12, x123, y123, x1234, y1234, level + 1);
    rec_bezier(s, tol, x1234, y1234, x234, y234, x34, y34, x4, y4, level + 1);
}
uint32_t ovk_flatten_cubic(float x0, float y0, float x1, float y1, float x2, float y2, float x3, float y3, float tol, float *out,
                           uint32_t cap) {
    flat_sink s = {NULL, out, 0, cap};
    rec_bezier(&s, tol, x0, y0, x1, y1, x2, y2, x3, y3, 0);
    sink_pt(&s, x3, y3);
    return s.n;
}
static void curve_to_(ovk_ctx *c, float x1, float y1, float x2, float y2, float x3, float y3) { /* vkvg_context.c:541-566 */
    if (EQUF(x1, x2) && EQUF(x2, x3) && EQUF(y1, y2) && EQUF(y2, y3)) {
        if (path_empty(c) || (EQUF(cur_pos(c).x, x1) && EQUF(cur_pos(c).y, y1))) return;
    }
    c->simpleConvex = false;
    set_curve_start(c);
    if (path_empty(c)) add_point(c, x1, y1);
    v2    cp = cur_pos(c);
    float sx = 1, sy = 1;
    mat_scale_of(&c->mat, &sx, &sy);
    float     tol = fabs(0.25f / fmaxf(sx, sy));
    flat_sink s   = {c, NULL, 0, 0};
    rec_bezier(&s, tol, cp.x, cp.y, x1, y1, x2, y2, x3, y3, 0);
    add_point(c, x3, y3);
    set_curve_end(c);
}
void ovk_curve_to(ovk_ctx *c, float x1, float y1, float x2, float y2, float x3, float y3) { curve_to_(c, x1, y1, x2, y2, x3, y3); }
void ovk_rel_curve_to(ovk_ctx *c, float x1, float y1, float x2, float y2, float x3, float y3) { /* :600-611 */
    if (path_empty(c)) { c->status = 4; return; } /* VKVG_STATUS_NO_CURRENT_POINT */
    v2 cp = cur_pos(c);
    curve_to_(c, cp.x + x1, cp.y + y1, cp.x + x2, cp.y + y2, cp.x + x3, cp.y + y3);
}
void ovk_quadratic_to(ovk_ctx *c, float x1, float y1, float x2, float y2) { /* :567-577, quadraticFact is a double */
    const double qf = 2.0 / 3.0;
    float x0, y0;
    if (path_empty(c)) { x0 = x1; y0 = y1; } else ovk_get_current_point(c, &x0, &y0);
    curve_to_(c, x0 + (x1 - x0) * qf, y0 + (y1 - y0) * qf, x2 + (x1 - x2) * qf, y2 + (y1 - y2) * qf, x2, y2);
}

/* ------------------------------------------------------------------ */
/* state                                                              */
/* ------------------------------------------------------------------ */
void ovk_set_line_width(ovk_ctx *c, float w) { c->lineWidth = w; }
void ovk_set_miter_limit(ovk_ctx *c, float l) { c->miterLimit = l; }
void ovk_set_line_cap(ovk_ctx *c, int cap) { c->lineCap = cap; }
void ovk_set_line_join(ovk_ctx *c, int j) { c->lineJoin = j; }
void ovk_set_fill_rule(ovk_ctx *c, int r) { c->fillRule = r; }
void ovk_set_opacity(ovk_ctx *c, float o) { c->opacity = o; }
void ovk_set_operator(ovk_ctx *c, int op) { c->op = op == 0 ? 1 : (op == 3 ? 2 : 0); } /* vkvg_operator_t: CLEAR 0, SOURCE 1, OVER 2, DIFFERENCE 3 */
void ovk_set_dash(ovk_ctx *c, const float *d, uint32_t n, float off) { /* vkvg_context.c:1103-1115 */
    free(c->dashes); c->dashes = NULL;
    c->dashCount = n; c->dashOffset = off;
    if (!n) return;
    c->dashes = (float *)malloc(n * sizeof(float));
    memcpy(c->dashes, d, n * sizeof(float));
}
/* CreateRgbaf, src/vkvg_context_internal.h:60-62: truncation, premultiplied */
static uint32_t rgbaf(float r, float g, float b, float a) {
    return (((uint32_t)(a * 255.0f) & 0xFF) << 24) | (((uint32_t)(b * a * 255.0f) & 0xFF) << 16) |
           (((uint32_t)(g * a * 255.0f) & 0xFF) << 8) | ((uint32_t)(r * a * 255.0f) & 0xFF);
}
void ovk_set_source_rgba(ovk_ctx *c, float r, float g, float b, float a) { c->curColor = rgbaf(r, g, b, a); c->patType = PAT_SOLID; }
void ovk_set_source_color(ovk_ctx *c, uint32_t col) { c->curColor = col; c->patType = PAT_SOLID; }
static void load_stops(grad_t *g, const float *stops, uint32_t n) { /* vkvg_pattern.c:149-167 */
    g->count = 0;
    for (uint32_t i = 0; i < n && i < 16; i++) {
        const float *s = stops + 5 * i;
        g->colors[i][0] = s[4] * s[1]; g->colors[i][1] = s[4] * s[2]; g->colors[i][2] = s[4] * s[3]; g->colors[i][3] = s[4];
        g->stops[i] = s[0];
        g->count++;
    }
}
/* control points are transformed by the CTM when the source is set, src/vkvg_context_internal.c:774-826 */
void ovk_set_source_linear(ovk_ctx *c, float x0, float y0, float x1, float y1, const float *stops, uint32_t n) {
    grad_t g; memset(&g, 0, sizeof g);
    load_stops(&g, stops, n);
    if (g.count < 2) { c->status = 10; return; } /* VKVG_STATUS_PATTERN_INVALID_GRADIENT */
    g.cp[0][0] = x0; g.cp[0][1] = y0; g.cp[0][2] = x1; g.cp[0][3] = y1;
    mat_point(&c->mat, &g.cp[0][0], &g.cp[0][1]);
    mat_point(&c->mat, &g.cp[0][2], &g.cp[0][3]);
    c->grad = g; c->patType = PAT_LINEAR;
}
void ovk_set_source_radial(ovk_ctx *c, float cx0, float cy0, float r0, float cx1, float cy1, float r1, const float *stops,
                           uint32_t n) {
    grad_t g; memset(&g, 0, sizeof g);
    load_stops(&g, stops, n);
    if (g.count < 2) { c->status = 10; return; }
    /* vkvg_pattern_edit_radial, src/vkvg_pattern.c:95-118 */
    v2 c0 = {cx0, cy0}, c1 = {cx1, cy1};
    if (r0 > r1 - 1.0f) r0 = r1 - 1.0f;
    v2    u = v2sub(c0, c1);
    float l = v2len(u);
    if (l + r0 + 1.0f >= r1) { v2 v = v2div(u, l); c0 = v2add(c1, v2mul(v, r1 - r0 - 1.0f)); }
    g.cp[0][0] = c0.x; g.cp[0][1] = c0.y; g.cp[0][2] = r0; g.cp[0][3] = 0;
    g.cp[1][0] = c1.x; g.cp[1][1] = c1.y; g.cp[1][2] = r1; g.cp[1][3] = 0;
    mat_point(&c->mat, &g.cp[0][0], &g.cp[0][1]);
    mat_point(&c->mat, &g.cp[1][0], &g.cp[1][1]);
    mat_distance(&c->mat, &g.cp[0][2], &g.cp[0][3]); /* :820-821 (the .w partner is cp[0].w in both calls) */
    mat_distance(&c->mat, &g.cp[1][2], &g.cp[0][3]);
    c->grad = g; c->patType = PAT_RADIAL;
}
static void mat_invert(mat_t *m) { /* vkvg_matrix_invert, src/vkvg_matrix.c:107-147 */
    if (m->xy == 0. && m->yx == 0.) {
        m->x0 = -m->x0; m->y0 = -m->y0;
        if (m->xx != 1.f) { if (m->xx == 0.) return; m->xx = 1.f / m->xx; m->x0 *= m->xx; }
        if (m->yy != 1.f) { if (m->yy == 0.) return; m->yy = 1.f / m->yy; m->y0 *= m->yy; }
        return;
    }
    float det = m->xx * m->yy - m->yx * m->xy;
    if (!((det) * (det) >= 0.) || det == 0) return;
    float a = m->xx, b = m->yx, c = m->xy, d = m->yy, tx = m->x0, ty = m->y0;
    *m = (mat_t){d, -b, -c, a, c * ty - d * tx, b * tx - a * ty};
    float s = 1 / det;
    m->xx *= s; m->yx *= s; m->xy *= s; m->yy *= s; m->x0 *= s; m->y0 *= s;
}
static void set_mat_inv(ovk_ctx *c) { c->matInv = c->mat; mat_invert(&c->matInv); } /* internal.c:672-676 */
void ovk_translate(ovk_ctx *c, float dx, float dy) { mat_t t = {1, 0, 0, 1, dx, dy}; c->mat = mat_mul(&t, &c->mat); set_mat_inv(c); }
void ovk_scale(ovk_ctx *c, float sx, float sy) { mat_t t = {sx, 0, 0, sy, 0, 0}; c->mat = mat_mul(&t, &c->mat); set_mat_inv(c); }
void ovk_rotate(ovk_ctx *c, float r) { float s = sinf(r), co = cosf(r); mat_t t = {co, s, -s, co, 0, 0}; c->mat = mat_mul(&t, &c->mat); set_mat_inv(c); }
void ovk_set_matrix(ovk_ctx *c, const float m[6]) { c->mat = (mat_t){m[0], m[1], m[2], m[3], m[4], m[5]}; set_mat_inv(c); }
/* vkvg_set_source_surface (src/vkvg_context.c:1025-1033) / vkvg_set_source with a surface pattern (:1034-1040): img is the
 * source surface's premultiplied RGBA8 pixels (kept by the caller), (x, y) the source offset, extend a vkvg_extend_t, linear 0/1,
 * pat_matrix the pattern matrix or NULL.  keep_offset != 0: only the pattern changes (vkvg_set_source keeps source.xy). */
void ovk_set_source_surface(ovk_ctx *c, const uint32_t *img, uint32_t w, uint32_t h, float x, float y, int extend, int linear, const float *pat_matrix,
                            int keep_offset) {
    if (!keep_offset) { c->surf.sx = x; c->surf.sy = y; }
    c->surf.img = img; c->surf.w = w; c->surf.h = h; c->surf.extend = (uint32_t)extend; c->surf.linear = (uint32_t)linear;
    if (pat_matrix) { /* internal.c:762-770: folded into matInv until the next CTM change recomputes it */
        mat_t m = {pat_matrix[0], pat_matrix[1], pat_matrix[2], pat_matrix[3], pat_matrix[4], pat_matrix[5]};
        c->matInv = mat_mul(&c->matInv, &m);
    }
    c->patType = 1;
}
void ovk_get_matrix(ovk_ctx *c, float m[6]) { memcpy(m, &c->mat, sizeof(mat_t)); }
void ovk_identity_matrix(ovk_ctx *c) { c->mat = MAT_ID; set_mat_inv(c); }
/* the surface-paint part of the push constants a draw issued now would carry: source.xywh, matInv (xx yx xy yy x0 y0) */
void ovk_get_source_push(ovk_ctx *c, float out[10]) {
    out[0] = c->surf.sx; out[1] = c->surf.sy; out[2] = (float)c->surf.w; out[3] = (float)c->surf.h;
    memcpy(out + 4, &c->matInv, sizeof(mat_t));
}

/* ------------------------------------------------------------------ */
/* tessellation output                                                */
/* ------------------------------------------------------------------ */
static void add_vertex(ovk_ctx *c, v2 p) {
    if (c->vertCount == c->sizeVerts) { c->sizeVerts *= 2; c->verts = (v2 *)realloc(c->verts, c->sizeVerts * sizeof(v2)); }
    c->verts[c->vertCount++] = p;
}
static void add_tri(ovk_ctx *c, uint32_t a, uint32_t b, uint32_t d) {
    if (c->indCount + 3 > c->sizeInds) { c->sizeInds *= 2; c->inds = (uint32_t *)realloc(c->inds, c->sizeInds * 4); }
    c->inds[c->indCount++] = a; c->inds[c->indCount++] = b; c->inds[c->indCount++] = d;
}
static void add_rect_inds(ovk_ctx *c, uint32_t i) { add_tri(c, i, i + 2, i + 1); add_tri(c, i + 1, i + 2, i + 3); } /* :341-354 */

typedef struct { uint32_t iL, iR, cp, firstIdx; float hw, lhMax, arcStep; } stroke_t; /* internal.h:246-255 */
typedef struct { bool dashOn; uint32_t curDash; float curDashOffset, totDashLength; v2 normal; } dash_t; /* :238-244 */

/* one join — src/vkvg_context_internal.c:924-1163 */
static bool build_vb_step(ovk_ctx *c, stroke_t *str, bool isCurve) {
    v2    p0 = c->points[str->cp];
    v2    v0 = v2sub(p0, c->points[str->iL]);
    v2    v1 = v2sub(c->points[str->iR], p0);
    float length_v0 = v2len(v0), length_v1 = v2len(v1);
    if (length_v0 < FLT_EPSILON || length_v1 < FLT_EPSILON) return false;
    v2    v0n = v2div(v0, length_v0), v1n = v2div(v1, length_v1);
    float dot = v2dot(v0n, v1n);
    float det = v0n.x * v1n.y - v0n.y * v1n.x;
    if (EQUF(dot, 1.0f)) return false;
    uint32_t idx = c->vertCount;
    if (EQUF(dot, -1.0f)) { /* cusp */
        v2 vPerp = v2mul(v2perp(v0n), str->hw);
        add_vertex(c, v2add(p0, vPerp));
        add_vertex(c, v2sub(p0, vPerp));
        add_tri(c, idx, idx + 1, idx + 2);
        add_tri(c, idx, idx + 2, idx + 3);
        return true;
    }
    v2    bisec_n = v2norm(v2add(v0n, v1n));
    float alpha   = acosf(dot);
    if (det < 0) alpha = -alpha;
    float halfAlpha    = alpha / 2.f;
    float cosHalfAlpha = cosf(halfAlpha);
    float lh           = str->hw / cosHalfAlpha;
    v2    bisec_n_perp = v2perp(bisec_n);
    float rlh          = lh;
    if (dot < 0.f) rlh = fminf(rlh, fminf(length_v0, length_v1));
    v2 bisec = v2mul(bisec_n_perp, rlh);
    v2 in_pos, out_pos;
    if (rlh < lh) {
        v2 vnPerp  = length_v0 < length_v1 ? v2perp(v1n) : v2perp(v0n);
        v2 vHwPerp = v2mul(vnPerp, str->hw);
        double lbc = cosHalfAlpha * rlh;
        if (det < 0.f) {
            in_pos  = v2add(v2add(v2mul(vnPerp, -lbc), v2add(p0, bisec)), vHwPerp);
            out_pos = v2sub(p0, v2mul(bisec_n_perp, lh));
        } else {
            in_pos  = v2sub(v2add(v2mul(vnPerp, lbc), v2sub(p0, bisec)), vHwPerp);
            out_pos = v2add(p0, v2mul(bisec_n_perp, lh));
        }
    } else {
        if (det < 0.0) { in_pos = v2add(p0, bisec); out_pos = v2sub(p0, bisec); }
        else { in_pos = v2sub(p0, bisec); out_pos = v2add(p0, bisec); }
    }
    int join = c->lineJoin;
    if (isCurve) join = dot < 0.8f ? OVK_JOIN_ROUND : OVK_JOIN_MITER;
    if (join == OVK_JOIN_MITER) {
        if (lh > str->lhMax) {
            double x = (lh - str->lhMax) * cosHalfAlpha;
            v2 bisecPerp = v2mul(bisec_n, x);
            bisec        = v2mul(bisec_n_perp, str->lhMax);
            if (det < 0) {
                add_vertex(c, in_pos);
                v2 p = v2sub(p0, bisec);
                add_vertex(c, v2sub(p, bisecPerp));
                add_vertex(c, v2add(p, bisecPerp));
                add_tri(c, idx, idx + 2, idx + 1);
                add_tri(c, idx + 2, idx + 4, idx);
                add_tri(c, idx, idx + 3, idx + 4);
                return true;
            } else {
                v2 p = v2add(p0, bisec);
                add_vertex(c, v2sub(p, bisecPerp));
                add_vertex(c, in_pos);
                add_vertex(c, v2add(p, bisecPerp));
                add_tri(c, idx, idx + 2, idx + 1);
                add_tri(c, idx + 2, idx + 3, idx + 1);
                add_tri(c, idx + 1, idx + 3, idx + 4);
                return false;
            }
        } else {
            if (det < 0) { add_vertex(c, in_pos); add_vertex(c, out_pos); }
            else { add_vertex(c, out_pos); add_vertex(c, in_pos); }
            add_rect_inds(c, idx);
            return false;
        }
    } else {
        v2 vp = v2perp(v0n);
        if (det < 0) {
            add_vertex(c, (dot < 0 && rlh < lh) ? in_pos : v2add(p0, bisec));
            add_vertex(c, v2sub(p0, v2mul(vp, str->hw)));
        } else {
            add_vertex(c, v2add(p0, v2mul(vp, str->hw)));
            add_vertex(c, (dot < 0 && rlh < lh) ? in_pos : v2sub(p0, bisec));
        }
        if (join == OVK_JOIN_BEVEL) {
            if (det < 0) { add_tri(c, idx, idx + 2, idx + 1); add_tri(c, idx + 2, idx + 4, idx + 0); add_tri(c, idx, idx + 3, idx + 4); }
            else { add_tri(c, idx, idx + 2, idx + 1); add_tri(c, idx + 2, idx + 3, idx + 1); add_tri(c, idx + 1, idx + 3, idx + 4); }
        } else if (join == OVK_JOIN_ROUND) {
            if (!str->arcStep) str->arcStep = arc_step(c, str->hw);
            float a = acosf(vp.x);
            if (vp.y < 0) a = -a;
            if (det < 0) {
                a += PIF;
                float a1 = a + alpha;
                a -= str->arcStep;
                while (a > a1) { add_vertex(c, (v2){cosf(a) * str->hw + p0.x, sinf(a) * str->hw + p0.y}); a -= str->arcStep; }
            } else {
                float a1 = a + alpha;
                a += str->arcStep;
                while (a < a1) { add_vertex(c, (v2){cosf(a) * str->hw + p0.x, sinf(a) * str->hw + p0.y}); a += str->arcStep; }
            }
            uint32_t p0Idx = c->vertCount;
            add_tri(c, idx, idx + 2, idx + 1);
            if (det < 0) {
                for (uint32_t p = idx + 2; p < p0Idx; p++) add_tri(c, p, p + 1, idx);
                add_tri(c, p0Idx, p0Idx + 2, idx);
                add_tri(c, idx, p0Idx + 1, p0Idx + 2);
            } else {
                for (uint32_t p = idx + 2; p < p0Idx; p++) add_tri(c, p, p + 1, idx + 1);
                add_tri(c, p0Idx, p0Idx + 1, idx + 1);
                add_tri(c, idx + 1, p0Idx + 1, p0Idx + 2);
            }
        }
        vp = v2mul(v2perp(v1n), str->hw);
        add_vertex(c, det < 0 ? v2sub(p0, vp) : v2add(p0, vp));
    }
    return (det < 0);
}
/* caps — src/vkvg_context_internal.c:1165-1239 */
static void draw_cap(ovk_ctx *c, stroke_t *str, v2 p0, v2 n, bool isStart) {
    uint32_t firstIdx = c->vertCount;
    if (isStart) {
        v2 vhw = v2mul(n, str->hw);
        if (c->lineCap == OVK_CAP_SQUARE) p0 = v2sub(p0, vhw);
        vhw = v2perp(vhw);
        if (c->lineCap == OVK_CAP_ROUND) {
            if (!str->arcStep) str->arcStep = arc_step(c, str->hw);
            float a = acosf(n.x) + PIF_2;
            if (n.y < 0) a = PIF - a;
            float a1 = a + PIF;
            a += str->arcStep;
            while (a < a1) { add_vertex(c, (v2){cosf(a) * str->hw + p0.x, sinf(a) * str->hw + p0.y}); a += str->arcStep; }
            uint32_t p0Idx = c->vertCount;
            for (uint32_t p = firstIdx; p < p0Idx; p++) add_tri(c, p0Idx + 1, p, p + 1);
            firstIdx = p0Idx;
        }
        add_vertex(c, v2add(p0, vhw));
        add_vertex(c, v2sub(p0, vhw));
        add_rect_inds(c, firstIdx);
    } else {
        v2 vhw = v2mul(n, str->hw);
        if (c->lineCap == OVK_CAP_SQUARE) p0 = v2add(p0, vhw);
        vhw = v2perp(vhw);
        add_vertex(c, v2add(p0, vhw));
        add_vertex(c, v2sub(p0, vhw));
        firstIdx = c->vertCount;
        if (c->lineCap == OVK_CAP_ROUND) {
            if (!str->arcStep) str->arcStep = arc_step(c, str->hw);
            float a = acosf(n.x) + PIF_2;
            if (n.y < 0) a = PIF - a;
            float a1 = a - PIF;
            a -= str->arcStep;
            while (a > a1) { add_vertex(c, (v2){cosf(a) * str->hw + p0.x, sinf(a) * str->hw + p0.y}); a -= str->arcStep; }
            uint32_t p0Idx = c->vertCount - 1;
            for (uint32_t p = firstIdx - 1; p < p0Idx; p++) add_tri(c, p + 1, p, firstIdx - 2);
        }
    }
}
/* src/vkvg_context_internal.c:1240-1264 */
static void draw_dashed_segment(ovk_ctx *c, stroke_t *str, dash_t *dc, bool isCurve) {
    v2 p = c->points[str->cp], pR = c->points[str->iR];
    if (!dc->dashOn) build_vb_step(c, str, isCurve);
    v2 d       = v2sub(pR, p);
    dc->normal = v2norm(d);
    float segmentLength = v2len(d);
    while (dc->curDashOffset < segmentLength) {
        v2 p0 = v2add(p, v2mul(dc->normal, dc->curDashOffset));
        draw_cap(c, str, p0, dc->normal, dc->dashOn);
        dc->dashOn ^= true;
        dc->curDashOffset += c->dashes[dc->curDash];
        if (++dc->curDash == c->dashCount) dc->curDash = 0;
    }
    dc->curDashOffset -= segmentLength;
    dc->curDashOffset = fmodf(dc->curDashOffset, dc->totDashLength);
}
static void draw_segment(ovk_ctx *c, stroke_t *str, dash_t *dc, bool isCurve) { /* :1265-1282 (no 2^32/3 batch split) */
    str->iR = str->cp + 1;
    if (c->dashCount > 0) draw_dashed_segment(c, str, dc, isCurve);
    else build_vb_step(c, str, isCurve);
    str->iL = str->cp++;
}
/* src/vkvg_context.c:822-948 */
static void tessellate_stroke(ovk_ctx *c) {
    c->vertCount = c->indCount = 0;
    stroke_t str; memset(&str, 0, sizeof str);
    str.hw    = c->lineWidth * 0.5f;
    str.lhMax = c->miterLimit * c->lineWidth;
    uint32_t ptrPath = 0;
    while (ptrPath < c->pathPtr) {
        uint32_t ptrSegment = 0, lastSegmentPointIdx = 0;
        uint32_t firstPathPointIdx = str.cp;
        uint32_t pathPointCount    = c->pathes[ptrPath] & P_MASK;
        uint32_t lastPathPointIdx  = str.cp + pathPointCount - 1;
        bool     has_curves = c->pathes[ptrPath] & P_CURVES, closed = c->pathes[ptrPath] & P_CLOSED;
        dash_t dc; memset(&dc, 0, sizeof dc);
        if (has_curves) {
            ptrSegment          = 1;
            lastSegmentPointIdx = str.cp + (c->pathes[ptrPath + ptrSegment] & P_MASK) - 1;
        }
        str.firstIdx = c->vertCount;
        if (c->dashCount > 0) {
            dc.dashOn = true;
            for (uint32_t i = 0; i < c->dashCount; i++) dc.totDashLength += c->dashes[i];
            if (dc.totDashLength == 0) { c->status = 13; return; } /* VKVG_STATUS_INVALID_DASH */
            dc.curDashOffset = fmodf(c->dashOffset, dc.totDashLength);
            str.iL           = lastPathPointIdx;
        } else if (closed) {
            str.iL = lastPathPointIdx;
        } else {
            draw_cap(c, &str, c->points[str.cp], v2norm(v2sub(c->points[str.cp + 1], c->points[str.cp])), true);
            str.iL = str.cp++;
        }
        if (has_curves) {
            while (str.cp < lastPathPointIdx) {
                bool curved = c->pathes[ptrPath + ptrSegment] & P_CURVES;
                if (lastSegmentPointIdx == lastPathPointIdx) lastSegmentPointIdx--;
                while (str.cp <= lastSegmentPointIdx) draw_segment(c, &str, &dc, curved);
                ptrSegment++;
                uint32_t cptSegPts  = c->pathes[ptrPath + ptrSegment] & P_MASK;
                lastSegmentPointIdx = str.cp + cptSegPts - 1;
                if (lastSegmentPointIdx == lastPathPointIdx && cptSegPts == 1) { ptrSegment++; break; }
            }
        } else
            while (str.cp < lastPathPointIdx) draw_segment(c, &str, &dc, false);
        if (c->dashCount > 0) {
            if (closed) {
                str.iR = firstPathPointIdx;
                draw_dashed_segment(c, &str, &dc, false);
                str.iL++; str.cp++;
            }
            if (!dc.dashOn) {
                /* :907-916.  The reference indexes dashes[curDash-1]; with an even dash count curDash is odd
                 * here so the index is valid (odd counts hit UB in the reference and are excluded) */
                int32_t prevDash = (int32_t)dc.curDash - 1;
                if (prevDash < 0) { dc.curDash = c->dashCount - 1; prevDash = 0; }
                float m = fminf(c->dashes[prevDash] - dc.curDashOffset, c->dashes[dc.curDash]);
                v2    p = v2sub(c->points[str.iR], v2mul(dc.normal, m));
                draw_cap(c, &str, p, dc.normal, false);
            }
        } else if (closed) {
            str.iR = firstPathPointIdx;
            bool inverse = build_vb_step(c, &str, false);
            uint32_t *inds = &c->inds[c->indCount - 6];
            uint32_t  ii   = str.firstIdx;
            if (inverse) { inds[1] = ii + 1; inds[4] = ii + 1; inds[5] = ii; }
            else { inds[1] = ii; inds[4] = ii; inds[5] = ii + 1; }
            str.cp++;
        } else
            draw_cap(c, &str, c->points[str.cp], v2norm(v2sub(c->points[str.cp], c->points[str.cp - 1])), false);
        str.cp = firstPathPointIdx + pathPointCount;
        ptrPath += ptrSegment > 0 ? ptrSegment : 1;
    }
}

/* ------------------------------------------------------------------ */
/* draws                                                              */
/* ------------------------------------------------------------------ */
static paint_t cur_paint(ovk_ctx *c) {
    paint_t p;
    g_surf_src = c->patType == 1 ? &c->surf : NULL;
    if (g_surf_src) { c->surf.minv = c->matInv; }
    p.patType = c->patType; p.grad = c->grad; p.solid = c->curColor; p.opacity = c->opacity;
    p.op = g_blend_op = c->op;
    return p;
}
static void snap_all(ovk_ctx *c, const v2 *pts, uint32_t n, int32_t *out) {
    for (uint32_t i = 0; i < n; i++) vs_chain(&c->mat, (float)c->W, (float)c->H, pts[i].x, pts[i].y, &out[2 * i], &out[2 * i + 1]);
}
/* indexed triangle list blended in order (stroke, convex non-zero) */
static void draw_indexed(ovk_ctx *c, const v2 *verts, uint32_t nv, const uint32_t *inds, uint32_t ni) {
    int32_t *fx = (int32_t *)malloc((size_t)nv * 8 + 8);
    snap_all(c, verts, nv, fx);
    paint_t    p = cur_paint(c);
    if (c->analytic) { /* every triangle oriented to wind +1: A = total triangle area inside the pixel, coverage = min(A, 1) */
        int32_t *e = (int32_t *)malloc((size_t)(ni / 3 + 1) * 3 * 16);
        uint64_t n = 0;
        for (uint32_t t = 0; t + 2 < ni; t += 3) {
            if (inds[t] >= nv || inds[t + 1] >= nv || inds[t + 2] >= nv) continue;
            const int32_t *a = fx + 2 * inds[t], *b = fx + 2 * inds[t + 1], *d = fx + 2 * inds[t + 2];
            int64_t area = (int64_t)(b[0] - a[0]) * (d[1] - a[1]) - (int64_t)(d[0] - a[0]) * (b[1] - a[1]);
            if (area == 0) continue;
            if (area > 0) { const int32_t *tmp = b; b = d; d = tmp; } /* cross > 0 winds -1 under W's convention: reverse */
            const int32_t *v[4] = {a, b, d, a};
            for (int k = 0; k < 3; k++, n++) { e[4 * n] = v[k][0]; e[4 * n + 1] = v[k][1]; e[4 * n + 2] = v[k + 1][0]; e[4 * n + 3] = v[k + 1][1]; }
        }
        analytic_draw(c, e, n, OVK_RULE_COUNT, p.patType, &p.grad, p.solid, p.opacity);
        free(e); free(fx);
        return;
    }
    blend_user u = {&p, 0x2};
    cov_reset(c);
    for (uint32_t t = 0; t + 2 < ni; t += 3) {
        if (inds[t] >= nv || inds[t + 1] >= nv || inds[t + 2] >= nv) continue; /* dangling forward refs of a skipped join */
        int32_t v[6] = {fx[2 * inds[t]], fx[2 * inds[t] + 1], fx[2 * inds[t + 1]], fx[2 * inds[t + 1] + 1],
                        fx[2 * inds[t + 2]], fx[2 * inds[t + 2] + 1]};
        raster_tri(c, v, full_rect(c), px_blend, &u);
    }
    free(fx);
}
/* walks the sub-path table the way _poly_fill / _fill_non_zero do (internal.c:1615-1653) */
typedef void (*subpath_fn)(ovk_ctx *c, uint32_t first, uint32_t count, void *user);
static void for_each_subpath(ovk_ctx *c, subpath_fn fn, void *user) {
    uint32_t ptrPath = 0, firstPtIdx = 0;
    while (ptrPath < c->pathPtr) {
        uint32_t n = c->pathes[ptrPath] & P_MASK;
        if (n > 2) fn(c, firstPtIdx, n, user);
        firstPtIdx += n;
        if (c->pathes[ptrPath] & P_CURVES) {
            ptrPath++;
            uint32_t tot = 0;
            while (tot < n) tot += (c->pathes[ptrPath++] & P_MASK);
        } else
            ptrPath++;
    }
}
typedef struct { float xMin, yMin, xMax, yMax; int32_t *fx; } eo_user;
static void eo_fan(ovk_ctx *c, uint32_t first, uint32_t n, void *user) { /* internal.c:1617-1642 */
    eo_user *u = (eo_user *)user;
    for (uint32_t i = 0; i < n; i++) {
        float x = c->points[first + i].x, y = c->points[first + i].y;
        mat_point(&c->mat, &x, &y);
        if (x < u->xMin) u->xMin = x;
        if (x > u->xMax) u->xMax = x;
        if (y < u->yMin) u->yMin = y;
        if (y > u->yMax) u->yMax = y;
    }
    snap_all(c, c->points + first, n, u->fx);
    for (uint32_t i = 1; i + 1 < n; i++) {
        int32_t v[6] = {u->fx[0], u->fx[1], u->fx[2 * i], u->fx[2 * i + 1], u->fx[2 * i + 2], u->fx[2 * i + 3]};
        raster_tri(c, v, full_rect(c), px_invert, NULL);
    }
}
typedef struct { int32_t *e; uint64_t n, cap; int64_t minx, miny, maxx, maxy; v2 *ua, *ub; uint32_t nu; } nz_user;
/* the sub-paths _fill_non_zero hands to libtess (> 2 points, internal.c:1757-1775), as user-space edges */
static void nz_collect(ovk_ctx *c, uint32_t first, uint32_t n, void *user) {
    nz_user *u = (nz_user *)user;
    for (uint32_t i = 0; i < n; i++) {
        u->ua[u->nu]   = c->points[first + i];
        u->ub[u->nu++] = c->points[first + (i + 1) % n];
    }
}
/* Proper crossing of edge A = a -> b with edge B = c -> d (A the one that comes first in the path): parameters along both and the
 * crossing point as libtess's combine callback stores it - computed in double (libtess works in GLdouble), kept as a float vec2
 * (combine2, src/vkvg_context_internal.c:1706-1712).  The same sequence of operations as nz_cross in vkvg_b200/csrc/raster.cu. */
static bool nz_cross(v2 a, v2 b, v2 c, v2 d, double *t, double *u, v2 *p) {
    if (fmaxf(a.x, b.x) < fminf(c.x, d.x) || fmaxf(c.x, d.x) < fminf(a.x, b.x) || fmaxf(a.y, b.y) < fminf(c.y, d.y) || fmaxf(c.y, d.y) < fminf(a.y, b.y)) return false;
    const double rx = (double)b.x - (double)a.x, ry = (double)b.y - (double)a.y, sx = (double)d.x - (double)c.x, sy = (double)d.y - (double)c.y;
    const double den = rx * sy - ry * sx;
    if (den == 0.0) return false;
    const double qx = (double)c.x - (double)a.x, qy = (double)c.y - (double)a.y;
    const double tn = qx * sy - qy * sx, un = qx * ry - qy * rx;   /* t = tn / den, u = un / den: proper when both lie strictly inside (0, 1) */
    if (den > 0.0 ? !(tn > 0.0 && tn < den && un > 0.0 && un < den) : !(tn < 0.0 && tn > den && un < 0.0 && un > den)) return false;
    *t = tn / den;
    *u = un / den;
    p->x = (float)((double)a.x + *t * rx);
    p->y = (float)((double)a.y + *t * ry);
    return true;
}
#define OVK_NZ_SPLIT_MAX 1024 /* paths of more edges are not split (the search is quadratic); same limit as VKB_NZ_SPLIT_MAX */
typedef struct { double key; uint32_t other; v2 p; } nz_hit;
static int nz_hit_cmp(const void *A, const void *B) {
    const nz_hit *a = (const nz_hit *)A, *b = (const nz_hit *)B;
    if (a->key != b->key) return a->key < b->key ? -1 : 1;
    return a->other < b->other ? -1 : (a->other > b->other ? 1 : 0);
}
static void nz_push_edge(ovk_ctx *c, nz_user *u, v2 a, v2 b) {
    if (u->n == u->cap) { u->cap = u->cap ? u->cap * 2 : 256; u->e = (int32_t *)realloc(u->e, (size_t)u->cap * 16); }
    v2      pts[2] = {a, b};
    int32_t fx[4];
    snap_all(c, pts, 2, fx);
    int32_t *e = u->e + 4 * u->n++;
    e[0] = fx[0]; e[1] = fx[1]; e[2] = fx[2]; e[3] = fx[3];
    for (int k = 0; k < 2; k++) {
        if (fx[2 * k] < u->minx) u->minx = fx[2 * k];
        if (fx[2 * k] > u->maxx) u->maxx = fx[2 * k];
        if (fx[2 * k + 1] < u->miny) u->miny = fx[2 * k + 1];
        if (fx[2 * k + 1] > u->maxy) u->maxy = fx[2 * k + 1];
    }
}
/* Device-space edges of a NON_ZERO fill / clip.  The reference blends the triangles libtess makes of the path
 * (src/vkvg_context_internal.c:1748-1792, GLU_TESS_WINDING_NONZERO).  libtess adds a vertex wherever two edges cross, and that vertex
 * reaches the vertex buffer as a float (combine2) and the rasteriser on the 1/256 grid like any other, so the region its triangles tile
 * is { winding != 0 } of the path whose edges are BENT at those vertices - not of the straight edges: next to every self-intersection
 * single samples differ (measured on C2: 0.43 % of the frame's pixels, p99.9 = 11/255, without this).  Restated here: every edge is
 * split at its proper crossings with the other edges of the path and the winding rule runs on the pieces. */
static void nz_edges(ovk_ctx *c, nz_user *u, bool split) {
    memset(u, 0, sizeof *u);
    u->minx = u->miny = INT64_MAX; u->maxx = u->maxy = INT64_MIN;
    u->ua = (v2 *)malloc((size_t)(c->pointCount + 1) * sizeof(v2));
    u->ub = (v2 *)malloc((size_t)(c->pointCount + 1) * sizeof(v2));
    for_each_subpath(c, nz_collect, u);
    const uint32_t n = u->nu;
    nz_hit *hits = (nz_hit *)malloc((size_t)(n + 1) * sizeof(nz_hit));
    for (uint32_t i = 0; i < n; i++) {
        uint32_t m = 0;
        if (split && n <= OVK_NZ_SPLIT_MAX)
            for (uint32_t j = 0; j < n; j++) {
                if (j == i) continue;
                double t, w;
                v2     p;
                if (i < j ? nz_cross(u->ua[i], u->ub[i], u->ua[j], u->ub[j], &t, &w, &p) : nz_cross(u->ua[j], u->ub[j], u->ua[i], u->ub[i], &w, &t, &p))
                    hits[m++] = (nz_hit){t, j, p};   /* t: the parameter along edge i */
            }
        qsort(hits, m, sizeof(nz_hit), nz_hit_cmp);
        v2 prev = u->ua[i];
        for (uint32_t k = 0; k < m; k++) { nz_push_edge(c, u, prev, hits[k].p); prev = hits[k].p; }
        nz_push_edge(c, u, prev, u->ub[i]);
    }
    free(hits); free(u->ua); free(u->ub);
    u->ua = u->ub = NULL;
}
/* libtess's fast path (external/glutess/src/tess.c:376-443 CacheVertex, render.c:362-516 __gl_renderCache): a polygon of ONE contour
 * (the reference opens a contour per sub-path of > 2 points, internal.c:1757-1775) of at most TESS_MAX_CACHE = 100 vertices never
 * reaches the sweep: when the triangles (v0, vk, vk+1) of the fan about its first vertex all turn the same way it is emitted as that
 * fan - original vertices only, no vertex at self-intersections, overlapping triangles where the polygon winds twice (the reference
 * then blends those samples twice).  Returns the number of points of that contour and its first point when the fast path applies with
 * a consistent orientation, 0 when the fan is all-degenerate or the polygon goes to the sweep (*sweep = true). */
typedef struct { uint32_t n_contours, first, n; } fan_user;
static void fan_collect(ovk_ctx *c, uint32_t first, uint32_t n, void *user) {
    fan_user *u = (fan_user *)user;
    (void)c;
    if (u->n_contours++ == 0) { u->first = first; u->n = n; }
}
static uint32_t nz_fan(ovk_ctx *c, uint32_t *first, bool *sweep) {
    fan_user u = {0, 0, 0};
    for_each_subpath(c, fan_collect, &u);
    *sweep = true;
    if (u.n_contours != 1 || u.n > 100) return 0;
    const v2 *p = c->points + u.first;
    /* ComputeNormal(check = FALSE): the normal is +-z; its sign is that of the first fan triangle that is not degenerate (later
     * back-facing contributions are reversed, so the sum never changes sign); ComputeNormal(check = TRUE): every non-degenerate
     * triangle must agree with it */
    double norm = 0.0;
    int    sign = 0;
    for (int pass = 0; pass < 2; pass++) {
        double xc = (double)p[1].x - (double)p[0].x, yc = (double)p[1].y - (double)p[0].y;
        for (uint32_t k = 2; k < u.n; k++) {
            const double xp = xc, yp = yc;
            xc = (double)p[k].x - (double)p[0].x; yc = (double)p[k].y - (double)p[0].y;
            const double nz = xp * yc - yp * xc, dot = nz * norm;
            if (pass == 0) { if (dot >= 0) norm += nz; else norm -= nz; }
            else if (dot != 0) {
                if (dot > 0) { if (sign < 0) return 0; sign = 1; }
                else { if (sign > 0) return 0; sign = -1; }
            }
        }
    }
    *sweep = false;          /* rendered (or dropped) from the cache */
    *first = u.first;
    return sign ? u.n : 0;   /* sign == 0: all triangles degenerate, nothing is drawn */
}
/* whether the edges of the current path are split at their crossings: NON_ZERO rule and a polygon that reaches libtess's sweep */
static bool nz_wants_split(ovk_ctx *c) {
    if (c->fillRule == OVK_FILL_EVEN_ODD) return false;
    if (c->pathPtr == 1 && (c->pathes[0] & P_CONVEX)) return false;  /* the reference's own fan, internal.c:1726-1746 */
    uint32_t first;
    bool     sweep;
    nz_fan(c, &first, &sweep);
    return sweep;
}
static void fill_preserve_(ovk_ctx *c) { /* vkvg_context.c:796-821 */
    finish_path(c);
    if (!c->pathPtr) return;
    paint_t p = cur_paint(c);
    cov_reset(c);
    if (c->analytic) { /* polygon edges of every sub-path with > 2 points, either rule */
        nz_user u;
        nz_edges(c, &u, nz_wants_split(c));
        if (u.n) analytic_draw(c, u.e, u.n, c->fillRule == OVK_FILL_EVEN_ODD ? OVK_RULE_EVEN_ODD : OVK_RULE_NON_ZERO, p.patType, &p.grad, p.solid, p.opacity);
        free(u.e);
        return;
    }
    if (c->fillRule == OVK_FILL_EVEN_ODD) {
        eo_user u = {FLT_MAX, FLT_MAX, FLT_MIN, FLT_MIN, (int32_t *)malloc((size_t)c->pointCount * 8 + 8)};
        for_each_subpath(c, eo_fan, &u);
        free(u.fx);
        /* cover pass scissor, internal.c:1924-1926 */
        int32_t sx = (int32_t)fmaxf(u.xMin, 0), sy = (int32_t)fmaxf(u.yMin, 0);
        float   fw = u.xMax - (int32_t)u.xMin + 1, fh = u.yMax - (int32_t)u.yMin + 1;
        int32_t sw = (int32_t)(fw > 1 ? fw : 1), sh = (int32_t)(fh > 1 ? fh : 1);
        cover_rect(c, clip_rect(c, sx, sy, sw, sh), &p, 0x1);
        return;
    }
    if (c->pathPtr == 1 && (c->pathes[0] & P_CONVEX)) { /* internal.c:1726-1746: plain fan, blended directly */
        uint32_t n = c->pathes[0] & P_MASK;
        c->vertCount = c->indCount = 0;
        for (uint32_t i = 0; i < n; i++) add_vertex(c, c->points[i]);
        for (uint32_t i = 2; i < n; i++) add_tri(c, 0, i - 1, i);
        draw_indexed(c, c->verts, c->vertCount, c->inds, c->indCount);
        return;
    }
    /* general non-zero: the reference tessellates with libtess (internal.c:1748-1792).  Its one-contour fast path emits the fan about the
     * first vertex as it is (nz_fan); everything else goes through the sweep, whose non-overlapping triangles are blended once each:
     * restated as "blend once where the winding number != 0" of the edges split at their crossings (nz_edges) */
    {
        uint32_t first = 0;
        bool     sweep = true;
        const uint32_t n = nz_fan(c, &first, &sweep);
        if (!sweep) {
            if (n) {
                c->vertCount = c->indCount = 0;
                for (uint32_t i = 0; i < n; i++) add_vertex(c, c->points[first + i]);
                for (uint32_t i = 2; i < n; i++) add_tri(c, 0, i - 1, i);
                draw_indexed(c, c->verts, c->vertCount, c->inds, c->indCount);
            }
            return;
        }
    }
    nz_user u;
    nz_edges(c, &u, nz_wants_split(c));
    if (u.n) {
        const int8_t(*sp)[2] = sample_table(c->S);
        rect_i     r = clip_rect(c, u.minx >> 8, u.miny >> 8, ((u.maxx >> 8) + 1) - (u.minx >> 8), ((u.maxy >> 8) + 1) - (u.miny >> 8));
        blend_user bu = {&p, 0x2};
        for (int32_t py = r.y0; py < r.y1; py++)
            for (int32_t px = r.x0; px < r.x1; px++) {
                uint32_t mask = 0;
                for (uint32_t s = 0; s < c->S; s++) {
                    int32_t w = 0;
                    int64_t sx = (int64_t)px * 256 + sp[s][0] * 16, sy = (int64_t)py * 256 + sp[s][1] * 16;
                    for (uint64_t i = 0; i < u.n; i++) w += edge_winding(u.e[4 * i], u.e[4 * i + 1], u.e[4 * i + 2], u.e[4 * i + 3], sx, sy);
                    if (w) mask |= 1u << s;
                    if (c->capture && w) c->coverage[PIX(c, px, py) * c->S + s] = w - 1; /* px_blend adds 1 */
                }
                if (mask) px_blend(c, (uint32_t)px, (uint32_t)py, mask, &bu);
            }
    }
    free(u.e);
}
void ovk_fill_preserve(ovk_ctx *c) { if (!c->status) fill_preserve_(c); }
void ovk_fill(ovk_ctx *c) { if (c->status) return; fill_preserve_(c); clear_path(c); }
static void stroke_preserve_(ovk_ctx *c) {
    finish_path(c);
    if (!c->pathPtr) return;
    tessellate_stroke(c);
    if (c->status) return;
    draw_indexed(c, c->verts, c->vertCount, c->inds, c->indCount);
}
void ovk_stroke_preserve(ovk_ctx *c) { if (!c->status) stroke_preserve_(c); }
void ovk_stroke(ovk_ctx *c) { if (c->status) return; stroke_preserve_(c); clear_path(c); }
void ovk_paint(ovk_ctx *c) { /* vkvg_context.c:990-1003 */
    if (c->status) return;
    finish_path(c);
    if (c->pathPtr) { ovk_fill(c); return; }
    paint_t p = cur_paint(c);
    cov_reset(c);
    if (c->analytic) {
        int32_t x1 = (int32_t)c->W * 256 + 4096, y1 = (int32_t)c->H * 256 + 4096;
        int32_t e[16] = {-4096, -4096, x1, -4096, x1, -4096, x1, y1, x1, y1, -4096, y1, -4096, y1, -4096, -4096};
        analytic_draw(c, e, 4, OVK_RULE_NON_ZERO, p.patType, &p.grad, p.solid, p.opacity);
        return;
    }
    cover_rect(c, full_rect(c), &p, 0x2);
}


/* ------------------------------------------------------------------ */
/* clipping and save / restore                                         */
/* ------------------------------------------------------------------ */
/* pipelineClipping: stencil test EQUAL on (ref & cmp) vs (stencil & cmp); pass -> REPLACE, fail -> ZERO, both under
 * the write mask; colour writes off (clipingOpState, src/vkvg_device_internal.c:233-239, dynamic ref / masks) */
static inline uint8_t clip_stencil_op(uint8_t st, uint32_t ref, uint32_t cmp, uint32_t write) {
    bool pass = (ref & cmp) == (st & cmp);
    return (uint8_t)(pass ? ((st & ~write) | (ref & write)) : (st & ~write));
}
typedef struct { uint32_t ref, cmp, write; } clip_user;
static void px_clip(ovk_ctx *c, uint32_t px, uint32_t py, uint32_t mask, void *user) {
    clip_user *u   = (clip_user *)user;
    size_t     base = PIX(c, px, py) * c->S;
    for (uint32_t s = 0; s < c->S; s++)
        if (mask & (1u << s)) c->stencil[base + s] = clip_stencil_op(c->stencil[base + s], u->ref, u->cmp, u->write);
}
static void stencil_quad(ovk_ctx *c, rect_i sc, uint32_t ref, uint32_t cmp, uint32_t write) { /* _draw_full_screen_quad under pipelineClipping */
    clip_user u = {ref, cmp, write};
    for (int32_t py = sc.y0; py < sc.y1; py++)
        for (int32_t px = sc.x0; px < sc.x1; px++) px_clip(c, (uint32_t)px, (uint32_t)py, (1u << c->S) - 1, &u);
}
static void clip_preserve_(ovk_ctx *c) { /* _clip_preserve, src/vkvg_context.c:754-795 */
    finish_path(c);
    if (!c->pathPtr) return;
    if (c->analytic) { /* one flag per pixel: inside where the coverage of the clip path exceeds one half */
        nz_user u;
        nz_edges(c, &u, nz_wants_split(c));
        if (!c->area) c->area = (double *)calloc((size_t)c->ww * c->wh, sizeof(double));
        ovk_area_brute(u.e, u.n, c->W, c->H, c->area);
        for (size_t i = 0; i < (size_t)c->ww * c->wh; i++) {
            float cov = analytic_coverage(c->area[i], c->fillRule == OVK_FILL_EVEN_ODD ? OVK_RULE_EVEN_ODD : OVK_RULE_NON_ZERO);
            if (!(cov > 0.5f))
                for (uint32_t s = 0; s < c->S; s++) c->stencil[i * c->S + s] |= 0x2;
        }
        free(u.e);
        c->curClipState = 2;
        return;
    }
    if (c->fillRule == OVK_FILL_EVEN_ODD) { /* fan inverts FILL where not yet clipped */
        eo_user u = {FLT_MAX, FLT_MAX, FLT_MIN, FLT_MIN, (int32_t *)malloc((size_t)c->pointCount * 8 + 8)};
        for_each_subpath(c, eo_fan, &u);
        free(u.fx);
    } else { /* non-zero: libtess triangles through pipelineClipping with ref FILL, compare CLIP, write FILL
              * (set FILL inside where not clipped); restated like fill_preserve_ as winding != 0 on the polygon edges */
        nz_user u;
        nz_edges(c, &u, nz_wants_split(c));
        if (u.n) {
            const int8_t(*sp)[2] = sample_table(c->S);
            rect_i    r = clip_rect(c, u.minx >> 8, u.miny >> 8, ((u.maxx >> 8) + 1) - (u.minx >> 8), ((u.maxy >> 8) + 1) - (u.miny >> 8));
            clip_user cu = {0x1, 0x2, 0x1};
            for (int32_t py = r.y0; py < r.y1; py++)
                for (int32_t px = r.x0; px < r.x1; px++) {
                    uint32_t mask = 0;
                    for (uint32_t s = 0; s < c->S; s++) {
                        int32_t w = 0;
                        int64_t sx = (int64_t)px * 256 + sp[s][0] * 16, sy = (int64_t)py * 256 + sp[s][1] * 16;
                        for (uint64_t i = 0; i < u.n; i++) w += edge_winding(u.e[4 * i], u.e[4 * i + 1], u.e[4 * i + 2], u.e[4 * i + 3], sx, sy);
                        if (w) mask |= 1u << s;
                    }
                    if (mask) px_clip(c, (uint32_t)px, (uint32_t)py, mask, &cu);
                }
        }
        free(u.e);
    }
    /* cover: ref CLIP, compare FILL, write ALL: FILL set -> 0 (kept), FILL clear -> CLIP */
    stencil_quad(c, full_rect(c), 0x2, 0x1, 0x3);
    c->curClipState = 2;
}
void ovk_clip_preserve(ovk_ctx *c) { if (!c->status) clip_preserve_(c); }
void ovk_clip(ovk_ctx *c) { if (c->status) return; clip_preserve_(c); clear_path(c); }

static int previous_clip_state(ovk_ctx *c) { return c->saved ? c->saved->clippingState : 1; } /* :698-702 */
void ovk_reset_clip(ovk_ctx *c) { /* :719-733; the clear wipes the whole stencil, save bits included (:706-717) */
    if (c->status) return;
    if (c->curClipState == 1) return;
    c->curClipState = previous_clip_state(c) == 1 ? 0 : 1;
    memset(c->stencil, 0, (size_t)c->ww * c->wh * c->S);
}
void ovk_clear_ctx(ovk_ctx *c) { /* vkvg_clear :734-753: clip state bookkeeping + colour and stencil wiped */
    if (c->status) return;
    c->curClipState = previous_clip_state(c) == 1 ? 0 : 1;
    ovk_clear(c);
}
void ovk_save(ovk_ctx *c) { /* :1251-1375 */
    if (c->status) return;
    struct ovk_save *sav = (struct ovk_save *)calloc(1, sizeof *sav);
    if (c->curClipState == 2) {
        sav->clippingState = 6;
        if (c->curSavBit > 0 && c->curSavBit % 6 == 0) { /* all six save bits in use: park the whole stencil */
            c->spills = (uint8_t **)realloc(c->spills, (c->nspills + 1) * sizeof(uint8_t *));
            c->spills[c->nspills] = (uint8_t *)malloc((size_t)c->ww * c->wh * c->S);
            memcpy(c->spills[c->nspills++], c->stencil, (size_t)c->ww * c->wh * c->S);
        }
        uint32_t bit = 1u << (c->curSavBit % 6 + 2);
        stencil_quad(c, full_rect(c), 0x2 | bit, 0x2, bit); /* save bit := CLIP */
        c->curSavBit++;
    } else if (c->curClipState == 0)
        sav->clippingState = previous_clip_state(c) & 0x03;
    else
        sav->clippingState = 1;
    sav->lineWidth = c->lineWidth; sav->miterLimit = c->miterLimit; sav->dashOffset = c->dashOffset; sav->dashCount = c->dashCount;
    if (c->dashCount) { sav->dashes = (float *)malloc(sizeof(float) * c->dashCount); memcpy(sav->dashes, c->dashes, sizeof(float) * c->dashCount); }
    sav->lineCap = c->lineCap; sav->fillRule = c->fillRule; sav->opacity = c->opacity; sav->mat = c->mat;
    sav->curColor = c->curColor; sav->patType = c->patType; sav->grad = c->grad; sav->matInv = c->matInv; sav->surf = c->surf;
    sav->next = c->saved;
    c->saved  = sav;
}
void ovk_restore(ovk_ctx *c) { /* :1376-1512 */
    if (c->status) return;
    if (!c->saved) { c->status = 3; /* VKVG_STATUS_INVALID_RESTORE */ return; }
    struct ovk_save *sav = c->saved;
    c->saved             = sav->next;
    if (c->curClipState) {
        if (c->curClipState == 2 && sav->clippingState == 1) memset(c->stencil, 0, (size_t)c->ww * c->wh * c->S); /* _reset_clip */
        else {
            uint32_t bit = 1u << ((c->curSavBit - 1) % 6 + 2);
            stencil_quad(c, full_rect(c), 0x2 | bit, bit, 0x2); /* CLIP := save bit */
        }
    }
    if (sav->clippingState == 6) {
        c->curSavBit--;
        if (c->curSavBit > 0 && c->curSavBit % 6 == 0) { /* the parked stencil comes back whole */
            memcpy(c->stencil, c->spills[--c->nspills], (size_t)c->ww * c->wh * c->S);
            free(c->spills[c->nspills]);
        }
    }
    c->curClipState = 0;
    c->dashOffset = sav->dashOffset;
    free(c->dashes);
    c->dashes = sav->dashes; c->dashCount = sav->dashCount;
    c->lineWidth = sav->lineWidth; c->miterLimit = sav->miterLimit; c->lineCap = sav->lineCap;
    c->lineJoin = OVK_JOIN_MITER; /* sav->lineJoint is never written by vkvg_save: restore reads the calloc'ed 0 (:1492) */
    c->fillRule = sav->fillRule; c->opacity = sav->opacity; c->mat = sav->mat;
    c->curColor = sav->curColor; c->patType = sav->patType; c->grad = sav->grad; c->matInv = sav->matInv; c->surf = sav->surf;
    free(sav);
}

uint32_t ovk_path_points(ovk_ctx *c, const float **pts) { finish_path(c); *pts = (const float *)c->points; return c->pointCount; }
uint32_t ovk_path_table(ovk_ctx *c, const uint32_t **t) { finish_path(c); *t = c->pathes; return c->pathPtr; }
uint32_t ovk_last_vertices(ovk_ctx *c, const float **xy) { *xy = (const float *)c->verts; return c->vertCount; }
uint32_t ovk_last_indices(ovk_ctx *c, const uint32_t **idx) { *idx = c->inds; return c->indCount; }

/* ------------------------------------------------------------------ */
/* rasterise a draw list recorded from the reference (oracle/ref_shim) */
/* ------------------------------------------------------------------ */
#define ORACLE_REF_SHIM_NO_PROTOS
#include "ref_shim.h"
typedef struct { float x, y; uint32_t color; float uv[3]; } ref_vertex; /* src/vkvg_context_internal.h:68-72 */

void ovk_raster_ref_drawlist(ovk_ctx *c, const void *draws_v, uint32_t n, const uint8_t *blob) {
    const ref_draw_t *draws = (const ref_draw_t *)draws_v;
    for (uint32_t i = 0; i < n; i++) {
        const ref_draw_t *d = &draws[i];
        if (d->kind == REF_DRAW_BEGIN_PASS) {
            if (d->pipeline == REF_RP_CLEAR_ALL) ovk_clear(c);
            else if (d->pipeline == REF_RP_CLEAR_STENCIL) memset(c->stencil, 0, (size_t)c->ww * c->wh * c->S);
            continue;
        }
        if (d->kind == REF_DRAW_CLEAR) {
            if (d->first & 1) { memset(c->samples, 0, (size_t)c->ww * c->wh * c->S * 4); c->resolved_dirty = true; }
            if (d->first & 4) memset(c->stencil, 0, (size_t)c->ww * c->wh * c->S);
            continue;
        }
        if (d->kind != REF_DRAW_ARRAYS && d->kind != REF_DRAW_INDEXED) continue;
        const ref_vertex *vb = (const ref_vertex *)(blob + (size_t)d->vbo * 16);
        const uint32_t   *ib = (const uint32_t *)(blob + (size_t)d->ibo * 16);
        const float      *pcf = (const float *)d->push;
        uint32_t          fsq_pat;
        memcpy(&fsq_pat, d->push + 24, 4);
        mat_t   M;
        memcpy(&M, d->push + 32, sizeof M);
        paint_t p;
        p.patType = fsq_pat & 0xFF;
        p.opacity = pcf[7];
        p.op = g_blend_op = d->pipeline == REF_PIPE_CLEAR ? 1 : (d->pipeline == REF_PIPE_SUB ? 2 : 0);
        g_surf_src = NULL; /* textures do not travel through the recorded draw list */
        memcpy(&p.grad, blob + (size_t)d->ubo * 16, sizeof(grad_t));
        rect_i sc = clip_rect(c, d->sc_x, d->sc_y, d->sc_w, d->sc_h);
        mat_t  saved = c->mat;
        c->mat       = M;
        if (d->kind == REF_DRAW_ARRAYS && d->pipeline == REF_PIPE_POLYFILL) {
            int32_t *fx = (int32_t *)malloc((size_t)d->count * 8 + 8);
            for (uint32_t k = 0; k < d->count; k++)
                vs_chain(&M, (float)c->W, (float)c->H, vb[d->first + k].x, vb[d->first + k].y, &fx[2 * k], &fx[2 * k + 1]);
            for (uint32_t k = 1; k + 1 < d->count; k++) {
                int32_t v[6] = {fx[0], fx[1], fx[2 * k], fx[2 * k + 1], fx[2 * k + 2], fx[2 * k + 3]};
                raster_tri(c, v, sc, px_invert, NULL);
            }
            free(fx);
        } else if (d->kind == REF_DRAW_ARRAYS && d->pipeline == REF_PIPE_CLIPPING) { /* clip cover / save / restore quad */
            stencil_quad(c, sc, d->ref, d->cmpMask, d->writeMask);
        } else if (d->kind == REF_DRAW_ARRAYS) { /* full-screen triangle, FULLSCREEN_BIT set (internal.c:1939-1942) */
            p.solid = vb[d->first].color;
            cover_rect(c, sc, &p, d->cmpMask);
        } else if (d->pipeline == REF_PIPE_CLIPPING) { /* non-zero clip: libtess triangles set FILL where not clipped */
            clip_user cu = {d->ref, d->cmpMask, d->writeMask};
            for (uint32_t t = 0; t + 2 < d->count; t += 3) {
                int32_t v[6];
                for (int k = 0; k < 3; k++) {
                    const ref_vertex *a = &vb[d->vertexOffset + (int64_t)ib[d->first + t + k]];
                    vs_chain(&M, (float)c->W, (float)c->H, a->x, a->y, &v[2 * k], &v[2 * k + 1]);
                }
                raster_tri(c, v, sc, px_clip, &cu);
            }
        } else {
            blend_user u = {&p, d->cmpMask};
            for (uint32_t t = 0; t + 2 < d->count; t += 3) {
                const ref_vertex *a = &vb[d->vertexOffset + (int64_t)ib[d->first + t]];
                const ref_vertex *b = &vb[d->vertexOffset + (int64_t)ib[d->first + t + 1]];
                const ref_vertex *e = &vb[d->vertexOffset + (int64_t)ib[d->first + t + 2]];
                int32_t v[6];
                vs_chain(&M, (float)c->W, (float)c->H, a->x, a->y, &v[0], &v[1]);
                vs_chain(&M, (float)c->W, (float)c->H, b->x, b->y, &v[2], &v[3]);
                vs_chain(&M, (float)c->W, (float)c->H, e->x, e->y, &v[4], &v[5]);
                p.solid = a->color;
                raster_tri(c, v, sc, px_blend, &u);
            }
        }
        c->mat = saved;
    }
}
