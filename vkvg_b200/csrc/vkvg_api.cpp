// Host side of the drop-in: the vkvg.h entry points.  Path commands are recorded (not flattened) into a
// vkb_batch; fill / stroke / paint append draws; vkvg_flush hands the batch to the CUDA pipeline.
// Every entry point cites the reference function whose behaviour (argument meaning, status handling,
// bookkeeping quirks) it mirrors.  There is deliberately no CPU rendering path: without a CUDA device
// vkvg_device_create returns an object whose status is VKVG_STATUS_DEVICE_ERROR.
#include "../../include/vkvg.h"
#include "../../include/vkvg_b200.h"
#include "renderer.h"
#include "decode_types.h"
#include <float.h>
#include <math.h>
#include <atomic>
#include <mutex>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <strings.h>
#include <vector>

#define M_PIF 3.14159265358979323846f
#define M_PIF_2 1.57079632679489661923f
#define EQUF(a, b) (fabsf((a) - (b)) <= FLT_EPSILON)

// objects returned instead of NULL on failed construction: the first field of every handle is its status
// (reference src/vkvg_internal.h:104-109)
static vkvg_status_t s_no_memory       = VKVG_STATUS_NO_MEMORY;
static vkvg_status_t s_null_pointer    = VKVG_STATUS_NULL_POINTER; // (kept for parity with the reference table)
static vkvg_status_t s_invalid_dev_ci  = VKVG_STATUS_INVALID_DEVICE_CREATE_INFO;
static vkvg_status_t s_device_error    = VKVG_STATUS_DEVICE_ERROR;
static vkvg_status_t s_invalid_surface = VKVG_STATUS_INVALID_SURFACE;
static vkvg_status_t s_invalid_image   = VKVG_STATUS_INVALID_IMAGE;
static vkvg_status_t s_file_not_found  = VKVG_STATUS_FILE_NOT_FOUND;

struct _vkvg_device_t {
    vkvg_status_t    status;
    std::atomic<uint32_t> references;  // handles may be shared between threads (the reference guards them with the device mutex when threadAware)
    uint32_t         samples;
    bool             analytic;  // vkvg_b200_device_set_coverage_mode: exact-area coverage, one colour per pixel
    uint32_t         raster_samples() const { return analytic ? 0u : samples; }  // what the pipeline is asked for
    int              hdpi, vdpi;
    bool             threadAware;
    vkb_device_impl *impl;
    std::mutex       mtx;  // one stream + shared scratch buffers per device: flushes are serialised
    bool             profiling;
    vkb_stats        last;
    vkvg_debug_stats_t dbg;  // high-water marks (vkvg_device_get_stats)
};
struct _vkvg_surface_t {
    vkvg_status_t     status;
    std::atomic<uint32_t> references;
    VkvgDevice        dev;
    uint32_t          width, height;
    vkb_surface_impl *impl;
    uint32_t          canvas_height, n_canvases;  // batch surfaces (vkvg_b200_surface_create_batch); 0 otherwise
};
struct _vkvg_pattern_t {
    vkvg_status_t       status;
    std::atomic<uint32_t> references;
    vkvg_pattern_type_t type;
    vkvg_extend_t       extend;
    vkvg_filter_t       filter;
    vkvg_matrix_t       matrix;
    bool                hasMatrix;
    vkb_gradient        grad;
    VkvgSurface         surf;  // VKVG_PATTERN_TYPE_SURFACE: the source (referenced)
};

struct saved_state {  // vkvg_context_save_t, src/vkvg_context_internal.h:101-125
    float              lineWidth, miterLimit, dashOffset, opacity;
    std::vector<float> dashes;
    vkvg_operator_t    op;
    vkvg_line_cap_t    cap;
    vkvg_fill_rule_t   fillRule;
    vkvg_matrix_t      mat;
    uint32_t           curColor;
    VkvgPattern        pattern;
    uint32_t           patType;
    vkb_gradient       grad;
    int                clippingState;  // vkvg_clip_state_t of the entry (src/vkvg_context_internal.h:93-99, :123)
    vkvg_matrix_t      matInv;         // the rest of the reference's push constants (src/vkvg_context_internal.h:74-81)
    float              src[4];
};
enum { CLIP_STATE_NONE = 0, CLIP_STATE_CLEAR = 1, CLIP_STATE_CLIP = 2, CLIP_STATE_CLIP_SAVED = 6 };

// ---- recording (reference src/recording/: the optional VKVG_RECORDING build) ----
// While a context records, drawing calls are stored instead of executed (RECORD macro, src/recording/vkvg_record_internal.h:117-123);
// vkvg_replay issues them on any context.  Command codes are the reference's (vkvg_record_internal.h:28-95) so that
// vkvg_recording_get_command reports the same values; arguments are stored as consecutive 4-byte floats / uint32s (dashes: count,
// offset, values; matrices: six floats; set_source / set_source_surface: an index into the recording's object table).
enum : uint16_t {
    RC_SAVE = 0x0001, RC_RESTORE = 0x0002,
    RC_NEW_PATH = 0x0101, RC_NEW_SUB_PATH = 0x0102, RC_CLOSE_PATH = 0x0103, RC_MOVE_TO = 0x0104, RC_LINE_TO = 0x0105, RC_RECTANGLE = 0x0106,
    RC_ARC = 0x0107, RC_ARC_NEG = 0x0108, RC_CURVE_TO = 0x010A, RC_QUADRATIC_TO = 0x010B, RC_ELLIPTICAL_ARC_TO = 0x010C,
    RC_REL_MOVE_TO = 0x0504, RC_REL_LINE_TO = 0x0505, RC_REL_CURVE_TO = 0x050A, RC_REL_QUADRATIC_TO = 0x050B, RC_REL_ELLIPTICAL_ARC_TO = 0x050C,
    RC_SET_LINE_WIDTH = 0x1101, RC_SET_MITER_LIMIT = 0x1102, RC_SET_LINE_JOIN = 0x1103, RC_SET_LINE_CAP = 0x1104, RC_SET_OPERATOR = 0x1105,
    RC_SET_FILL_RULE = 0x1106, RC_SET_DASH = 0x1107,
    RC_TRANSLATE = 0x2001, RC_ROTATE = 0x2002, RC_SCALE = 0x2003, RC_TRANSFORM = 0x2004, RC_IDENTITY_MATRIX = 0x2005, RC_SET_MATRIX = 0x2006,
    RC_PAINT = 0x0201, RC_FILL = 0x0202, RC_STROKE = 0x0203, RC_CLIP = 0x0204, RC_RESET_CLIP = 0x0205, RC_CLEAR = 0x0206,
    RC_FILL_PRESERVE = 0x0602, RC_STROKE_PRESERVE = 0x0603, RC_CLIP_PRESERVE = 0x0604,
    RC_SET_SOURCE_RGB = 0x0801, RC_SET_SOURCE_RGBA = 0x0802, RC_SET_SOURCE_COLOR = 0x0803, RC_SET_SOURCE = 0x0804, RC_SET_SOURCE_SURFACE = 0x0805,
};
struct rec_entry { uint16_t cmd; size_t off; };
struct _vkvg_recording_t {
    std::vector<rec_entry>   cmds;
    std::vector<char>        buf;
    std::vector<VkvgPattern> pats;   // referenced until the recording is destroyed
    std::vector<VkvgSurface> surfs;
};

struct _vkvg_context_t {
    vkvg_status_t status;
    std::atomic<uint32_t> references;
    VkvgDevice    dev;
    VkvgSurface   pSurf;

    vkb_batch batch;
    uint32_t  path_first_sp;  // first sub-path of the current path inside batch.subpaths
    // the open sub-path (the reference's pathes[pathPtr] + segmentPtr bookkeeping, internal.c:136-238)
    uint32_t sp_first_elem;
    uint32_t sp_points;   // point count; exact for lines/arcs, a lower bound (>= true count impossible to miss 3/4 tests) once a curve is in
    float    first_x, first_y, cur_x, cur_y;
    bool     simpleConvex;

    float              lineWidth, miterLimit, dashOffset;
    std::vector<float> dashes;
    vkvg_operator_t    op;
    vkvg_line_cap_t    cap;
    vkvg_line_join_t   join;
    vkvg_fill_rule_t   fillRule;
    float              opacity;
    vkvg_matrix_t      mat;
    uint32_t           curColor;
    VkvgPattern        pattern;
    uint32_t           patType;
    vkb_gradient       grad;      // gradient as uploaded by _update_cur_pattern (control points already through the CTM)
    int32_t            grad_slot; // index of `grad` in batch.grads, or -1
    bool               clear_pending;
    std::vector<saved_state> saved;
    // clip bookkeeping exactly as the reference keeps it (src/vkvg_context_internal.h:225-226): what the context believes
    // about the stencil relative to the last saved state, and how many clip saves are stacked
    int                curClipState;
    uint32_t           curSavBit;
    uint32_t           canvas;  // batch surfaces: the canvas later draws go to (vkvg_b200_set_canvas)
    // surface paints: pushConsts.matInv / .source of the reference (recomputed on every CTM change, src/vkvg_context_internal.c:672-676)
    vkvg_matrix_t      matInv;
    float              src[4];  // x, y (vkvg_set_source_surface offset), width, height of the source
    std::vector<VkvgSurface> held, held_prev;  // sources referenced by recorded / in-flight draws
    VkvgRecording      recording;  // non-null while vkvg_start_recording is in effect
};
// stores one command when the context records; the caller then returns without executing
static bool rec(VkvgContext ctx, uint16_t cmd, std::initializer_list<float> f = {}, std::initializer_list<uint32_t> u = {}) {
    VkvgRecording r = ctx->recording;
    if (!r) return false;
    r->cmds.push_back(rec_entry{cmd, r->buf.size()});
    for (float v : f) { const char *p = (const char *)&v; r->buf.insert(r->buf.end(), p, p + 4); }
    for (uint32_t v : u) { const char *p = (const char *)&v; r->buf.insert(r->buf.end(), p, p + 4); }
    return true;
}

// ====================================================================================================
// matrices — cairo-derived 2x3 affine, reference src/vkvg_matrix.c
// ====================================================================================================
void vkvg_matrix_init(vkvg_matrix_t *m, float xx, float yx, float xy, float yy, float x0, float y0) {
    m->xx = xx; m->yx = yx; m->xy = xy; m->yy = yy; m->x0 = x0; m->y0 = y0;
}
void vkvg_matrix_init_identity(vkvg_matrix_t *m) { vkvg_matrix_init(m, 1, 0, 0, 1, 0, 0); }
void vkvg_matrix_init_translate(vkvg_matrix_t *m, float tx, float ty) { vkvg_matrix_init(m, 1, 0, 0, 1, tx, ty); }
void vkvg_matrix_init_scale(vkvg_matrix_t *m, float sx, float sy) { vkvg_matrix_init(m, sx, 0, 0, sy, 0, 0); }
void vkvg_matrix_init_rotate(vkvg_matrix_t *m, float radians) {
    float s = sinf(radians), c = cosf(radians);
    vkvg_matrix_init(m, c, s, -s, c, 0, 0);
}
void vkvg_matrix_multiply(vkvg_matrix_t *result, const vkvg_matrix_t *a, const vkvg_matrix_t *b) {  // :193-206
    vkvg_matrix_t r;
    r.xx = a->xx * b->xx + a->yx * b->xy;
    r.yx = a->xx * b->yx + a->yx * b->yy;
    r.xy = a->xy * b->xx + a->yy * b->xy;
    r.yy = a->xy * b->yx + a->yy * b->yy;
    r.x0 = a->x0 * b->xx + a->y0 * b->xy + b->x0;
    r.y0 = a->x0 * b->yx + a->y0 * b->yy + b->y0;
    *result = r;
}
void vkvg_matrix_translate(vkvg_matrix_t *m, float tx, float ty) {
    vkvg_matrix_t t;
    vkvg_matrix_init_translate(&t, tx, ty);
    vkvg_matrix_multiply(m, &t, m);
}
void vkvg_matrix_scale(vkvg_matrix_t *m, float sx, float sy) {
    vkvg_matrix_t t;
    vkvg_matrix_init_scale(&t, sx, sy);
    vkvg_matrix_multiply(m, &t, m);
}
void vkvg_matrix_rotate(vkvg_matrix_t *m, float radians) {
    vkvg_matrix_t t;
    vkvg_matrix_init_rotate(&t, radians);
    vkvg_matrix_multiply(m, &t, m);
}
void vkvg_matrix_transform_distance(const vkvg_matrix_t *m, float *dx, float *dy) {
    float nx = (m->xx * *dx + m->xy * *dy), ny = (m->yx * *dx + m->yy * *dy);
    *dx = nx; *dy = ny;
}
void vkvg_matrix_transform_point(const vkvg_matrix_t *m, float *x, float *y) {
    vkvg_matrix_transform_distance(m, x, y);
    *x += m->x0; *y += m->y0;
}
vkvg_status_t vkvg_matrix_invert(vkvg_matrix_t *m) {  // :107-147
    if (m->xy == 0. && m->yx == 0.) {
        m->x0 = -m->x0; m->y0 = -m->y0;
        if (m->xx != 1.f) {
            if (m->xx == 0.) return VKVG_STATUS_INVALID_MATRIX;
            m->xx = 1.f / m->xx;
            m->x0 *= m->xx;
        }
        if (m->yy != 1.f) {
            if (m->yy == 0.) return VKVG_STATUS_INVALID_MATRIX;
            m->yy = 1.f / m->yy;
            m->y0 *= m->yy;
        }
        return VKVG_STATUS_SUCCESS;
    }
    float det = m->xx * m->yy - m->yx * m->xy;
    if (!((det) * (det) >= 0.) || det == 0) return VKVG_STATUS_INVALID_MATRIX;
    float a = m->xx, b = m->yx, c = m->xy, d = m->yy, tx = m->x0, ty = m->y0;
    vkvg_matrix_init(m, d, -b, -c, a, c * ty - d * tx, b * tx - a * ty);
    float s = 1 / det;
    m->xx *= s; m->yx *= s; m->xy *= s; m->yy *= s; m->x0 *= s; m->y0 *= s;
    return VKVG_STATUS_SUCCESS;
}
void vkvg_matrix_get_scale(const vkvg_matrix_t *m, float *sx, float *sy) {  // :222-229 (double sqrt, float store)
    *sx = sqrt(m->xx * m->xx + m->xy * m->xy);
    *sy = sqrt(m->yx * m->yx + m->yy * m->yy);
}

// ====================================================================================================
// device — reference src/vkvg_device.c
// ====================================================================================================
VkvgDevice vkvg_device_create(vkvg_device_create_info_t *info) {
    uint32_t samples = info ? info->samples : 1;
    if (samples == 0) samples = 1;
    if (samples != 1 && samples != 2 && samples != 4 && samples != 8 && samples != 16) return (VkvgDevice)&s_invalid_dev_ci;
    int         ordinal = 0;
    const char *e       = getenv("VKVG_B200_DEVICE");
    if (!e) e = getenv("LOCAL_RANK");
    if (e) ordinal = atoi(e);
    vkb_device_impl *impl = vkb_device_open(ordinal);
    if (!impl) {
        fprintf(stderr, "vkvg_b200: no usable CUDA device — this library has no CPU path (vkvg_device_create fails)\n");
        return (VkvgDevice)&s_device_error;
    }
    VkvgDevice dev   = new _vkvg_device_t();
    dev->status      = VKVG_STATUS_SUCCESS;
    dev->references  = 1;
    dev->samples     = samples;
    {
        const char *cm = getenv("VKVG_B200_COVERAGE");
        dev->analytic  = cm && !strcmp(cm, "analytic");
    }
    dev->hdpi = dev->vdpi = 96;
    dev->threadAware = info ? info->threadAware : false;
    dev->impl        = impl;
    dev->profiling   = getenv("VKVG_B200_PROFILE") != nullptr;
    memset(&dev->last, 0, sizeof dev->last);
    memset(&dev->dbg, 0, sizeof dev->dbg);
    return dev;
}
vkvg_status_t vkvg_device_status(VkvgDevice dev) { return !dev ? VKVG_STATUS_NULL_POINTER : dev->status; }
void vkvg_device_destroy(VkvgDevice dev) {
    if (vkvg_device_status(dev)) return;
    if (--dev->references > 0) return;
    vkb_device_close(dev->impl);
    delete dev;
}
VkvgDevice vkvg_device_reference(VkvgDevice dev) {
    if (!vkvg_device_status(dev)) dev->references++;
    return dev;
}
uint32_t vkvg_device_get_reference_count(VkvgDevice dev) { return vkvg_device_status(dev) ? 0u : dev->references.load(); }
void vkvg_device_set_dpy(VkvgDevice dev, int hdpy, int vdpy) {
    if (vkvg_device_status(dev)) return;
    dev->hdpi = hdpy; dev->vdpi = vdpy;
}
void vkvg_device_get_dpy(VkvgDevice dev, int *hdpy, int *vdpy) {
    if (vkvg_device_status(dev)) return;
    *hdpy = dev->hdpi; *vdpy = dev->vdpi;
}
// VKVG_DBG_STATS build of the reference (include/vkvg.h:331-349, src/vkvg_device.c:512-519): high-water marks of the path / vertex arrays.
// Here: of the device-side arrays the flushes of this device produced (points, sub-paths, stroke vertices / indices; the VBO / IBO
// entries report the same, the "buffers" being those arrays).
vkvg_debug_stats_t vkvg_device_get_stats(VkvgDevice dev) {
    vkvg_debug_stats_t z = {0, 0, 0, 0, 0, 0};
    return vkvg_device_status(dev) ? z : dev->dbg;
}
void vkvg_device_reset_stats(VkvgDevice dev) {
    if (vkvg_device_status(dev)) return;
    memset(&dev->dbg, 0, sizeof dev->dbg);
}
void vkvg_device_set_context_cache_size(VkvgDevice, uint32_t) {}  // contexts hold no device objects here: nothing to cache

// ====================================================================================================
// surface — reference src/vkvg_surface.c
// ====================================================================================================
static VkvgSurface create_surface(VkvgDevice dev, uint32_t width, uint32_t height, uint32_t full_height, uint32_t origin_y) {
    if (vkvg_device_status(dev)) return (VkvgSurface)&s_device_error;  // _create_surface, surface_internal.c:213-224
    VkvgSurface surf = new _vkvg_surface_t();
    surf->status     = VKVG_STATUS_SUCCESS;
    surf->references = 1;
    surf->dev        = dev;
    surf->width      = width > 1 ? width : 1;
    surf->height     = height > 1 ? height : 1;
    surf->canvas_height = surf->n_canvases = 0;
    {
        std::lock_guard<std::mutex> lk(dev->mtx);
        surf->impl = vkb_surface_new(dev->impl, surf->width, surf->height, full_height ? full_height : surf->height, origin_y);
    }
    if (!surf->impl || vkb_device_failed(dev->impl)) surf->status = VKVG_STATUS_DEVICE_ERROR;
    vkvg_device_reference(dev);
    return surf;
}
VkvgSurface vkvg_surface_create(VkvgDevice dev, uint32_t width, uint32_t height) { return create_surface(dev, width, height, 0, 0); }
// `count` independent canvases of width x height stacked in one surface of height count * height: draws recorded after
// vkvg_b200_set_canvas(ctx, i) land in canvas i exactly as they would on a surface of their own (same snapping, same
// gradients, clipped to the canvas), and one flush renders them all
VkvgSurface vkvg_b200_surface_create_batch(VkvgDevice dev, uint32_t width, uint32_t height, uint32_t count) {
    if (height == 0 || height % VKB_TILE != 0 || count == 0 || (uint64_t)height * count > 0x7fffffu) return (VkvgSurface)&s_invalid_surface;
    VkvgSurface surf = create_surface(dev, width, height * count, height, 0);
    if (vkvg_surface_status(surf)) return surf;
    surf->canvas_height = height; surf->n_canvases = count;
    std::lock_guard<std::mutex> lk(dev->mtx);
    vkb_surface_set_band_height(surf->impl, height);
    return surf;
}
vkvg_status_t vkvg_b200_set_canvas(VkvgContext ctx, uint32_t index) {
    if (vkvg_status(ctx)) return vkvg_status(ctx);
    if (!ctx->pSurf->n_canvases ? index != 0 : index >= ctx->pSurf->n_canvases) return VKVG_STATUS_INVALID_INDEX;
    ctx->canvas = index;
    return VKVG_STATUS_SUCCESS;
}
VkvgSurface vkvg_b200_surface_create_stripe(VkvgDevice dev, uint32_t width, uint32_t full_height, uint32_t origin_y, uint32_t height) {
    if (origin_y % VKB_TILE != 0 || origin_y + height > full_height) return (VkvgSurface)&s_invalid_surface;
    return create_surface(dev, width, height, full_height, origin_y);
}
vkvg_status_t vkvg_b200_surface_copy_to_device(VkvgSurface surf, void *device_dst) {
    if (vkvg_surface_status(surf) || !device_dst) return VKVG_STATUS_INVALID_SURFACE;
    std::lock_guard<std::mutex> lk(surf->dev->mtx);
    return vkb_surface_copy_to_device(surf->impl, device_dst) ? VKVG_STATUS_DEVICE_ERROR : VKVG_STATUS_SUCCESS;
}
vkvg_status_t vkvg_surface_status(VkvgSurface surf) { return !surf ? VKVG_STATUS_NULL_POINTER : surf->status; }
VkvgSurface   vkvg_surface_reference(VkvgSurface surf) {
    if (!vkvg_surface_status(surf)) surf->references++;
    return surf;
}
uint32_t vkvg_surface_get_reference_count(VkvgSurface surf) { return vkvg_surface_status(surf) ? 0u : surf->references.load(); }
void     vkvg_surface_destroy(VkvgSurface surf) {
    if (vkvg_surface_status(surf)) return;
    if (--surf->references > 0) return;
    {
        std::lock_guard<std::mutex> lk(surf->dev->mtx);
        vkb_surface_free(surf->impl);
    }
    vkvg_device_destroy(surf->dev);
    delete surf;
}
void vkvg_surface_clear(VkvgSurface surf) {
    if (vkvg_surface_status(surf)) return;
    std::lock_guard<std::mutex> lk(surf->dev->mtx);
    vkb_surface_clear(surf->impl);
}
VkImage  vkvg_surface_get_vk_image(VkvgSurface) { return NULL; }
VkFormat vkvg_surface_get_vk_format(VkvgSurface surf) { return vkvg_surface_status(surf) ? 0 : VK_FORMAT_B8G8R8A8_UNORM; }
uint32_t vkvg_surface_get_width(VkvgSurface surf) { return vkvg_surface_status(surf) ? 0 : surf->width; }
uint32_t vkvg_surface_get_height(VkvgSurface surf) { return vkvg_surface_status(surf) ? 0 : surf->height; }
void     vkvg_surface_resolve(VkvgSurface) {}  // every flush resolves
vkvg_status_t vkvg_surface_write_to_memory(VkvgSurface surf, unsigned char *const bitmap) {  // :393-467
    if (vkvg_surface_status(surf)) return VKVG_STATUS_INVALID_STATUS;
    if (!bitmap) return VKVG_STATUS_WRITE_ERROR;
    std::lock_guard<std::mutex> lk(surf->dev->mtx);
    return vkb_surface_download(surf->impl, bitmap, true) ? VKVG_STATUS_DEVICE_ERROR : VKVG_STATUS_SUCCESS;
}
vkvg_status_t vkvg_b200_surface_set_readback(VkvgSurface surf, unsigned char *host_rgba) {
    if (vkvg_surface_status(surf)) return VKVG_STATUS_INVALID_SURFACE;
    std::lock_guard<std::mutex> lk(surf->dev->mtx);
    vkb_surface_set_readback(surf->impl, host_rgba);
    return VKVG_STATUS_SUCCESS;
}
vkvg_status_t vkvg_b200_surface_ipc_export(VkvgSurface surf, unsigned char *handle64) {
    if (vkvg_surface_status(surf)) return VKVG_STATUS_INVALID_SURFACE;
    if (!handle64) return VKVG_STATUS_NULL_POINTER;
    std::lock_guard<std::mutex> lk(surf->dev->mtx);
    return vkb_surface_ipc_export(surf->impl, handle64) ? VKVG_STATUS_DEVICE_ERROR : VKVG_STATUS_SUCCESS;
}
void *vkvg_b200_ipc_open(VkvgDevice dev, const unsigned char *handle64) {
    if (vkvg_device_status(dev) || !handle64) return NULL;
    std::lock_guard<std::mutex> lk(dev->mtx);
    return vkb_ipc_open(dev->impl, handle64);
}
vkvg_status_t vkvg_b200_ipc_close(VkvgDevice dev, void *ptr) {
    if (vkvg_device_status(dev)) return VKVG_STATUS_DEVICE_ERROR;
    if (!ptr) return VKVG_STATUS_NULL_POINTER;
    std::lock_guard<std::mutex> lk(dev->mtx);
    return vkb_ipc_close(dev->impl, ptr) ? VKVG_STATUS_DEVICE_ERROR : VKVG_STATUS_SUCCESS;
}
vkvg_status_t vkvg_b200_surface_read_premultiplied(VkvgSurface surf, unsigned char *rgba) {
    if (vkvg_surface_status(surf)) return VKVG_STATUS_INVALID_STATUS;
    if (!rgba) return VKVG_STATUS_WRITE_ERROR;
    std::lock_guard<std::mutex> lk(surf->dev->mtx);
    return vkb_surface_download(surf->impl, rgba, false) ? VKVG_STATUS_DEVICE_ERROR : VKVG_STATUS_SUCCESS;
}
int vkb_write_png(const char *path, const unsigned char *rgba, uint32_t w, uint32_t h);  // png.cpp
vkvg_status_t vkvg_surface_write_to_png(VkvgSurface surf, const char *path) {  // :283-391
    if (vkvg_surface_status(surf) || vkvg_device_status(surf->dev)) return VKVG_STATUS_INVALID_STATUS;
    if (!path) return VKVG_STATUS_WRITE_ERROR;
    std::vector<unsigned char> img((size_t)surf->width * surf->height * 4);
    vkvg_status_t st = vkvg_surface_write_to_memory(surf, img.data());
    if (st) return st;
    return vkb_write_png(path, img.data(), surf->width, surf->height) ? VKVG_STATUS_WRITE_ERROR : VKVG_STATUS_SUCCESS;
}
int vkb_read_png(const char *path, std::vector<unsigned char> &rgba, uint32_t &w, uint32_t &h);  // png.cpp
VkvgSurface vkvg_surface_create_from_bitmap(VkvgDevice dev, unsigned char *img, uint32_t width, uint32_t height) {  // :81-170
    if (vkvg_device_status(dev)) return (VkvgSurface)&s_device_error;
    if (!img || width == 0 || height == 0) return (VkvgSurface)&s_invalid_image;
    VkvgSurface surf = create_surface(dev, width, height, 0, 0);
    if (vkvg_surface_status(surf)) return surf;
    // the reference paints the bitmap onto the cleared surface through the OVER pipeline, which leaves exactly these bytes
    std::lock_guard<std::mutex> lk(dev->mtx);
    if (vkb_surface_upload(surf->impl, img)) surf->status = VKVG_STATUS_DEVICE_ERROR;
    return surf;
}
VkvgSurface vkvg_surface_create_from_image(VkvgDevice dev, const char *filePath) {  // :171-186 (stb_image there; PNG only here)
    if (vkvg_device_status(dev)) return (VkvgSurface)&s_device_error;
    std::vector<unsigned char> rgba;
    uint32_t w = 0, h = 0;
    int      r = vkb_read_png(filePath, rgba, w, h);
    if (r == 1) return (VkvgSurface)&s_file_not_found;
    if (r) return (VkvgSurface)&s_invalid_image;
    return vkvg_surface_create_from_bitmap(dev, rgba.data(), w, h);
}
const void *vkvg_b200_surface_device_pointer(VkvgSurface surf) { return vkvg_surface_status(surf) ? nullptr : vkb_surface_device_pixels(surf->impl); }

// ====================================================================================================
// patterns — reference src/vkvg_pattern.c (solid / linear / radial; surface patterns are out of scope)
// ====================================================================================================
vkvg_status_t vkvg_pattern_status(VkvgPattern pat) { return !pat ? VKVG_STATUS_NULL_POINTER : pat->status; }
static VkvgPattern new_pattern(vkvg_pattern_type_t type) {
    VkvgPattern pat = new _vkvg_pattern_t();
    pat->status = VKVG_STATUS_SUCCESS; pat->filter = VKVG_FILTER_FAST; pat->hasMatrix = false; pat->surf = NULL;
    memset(&pat->grad, 0, sizeof pat->grad);
    vkvg_matrix_init_identity(&pat->matrix);
    pat->type       = type;
    pat->extend     = VKVG_EXTEND_NONE;
    pat->references = 1;
    return pat;
}
vkvg_status_t vkvg_pattern_edit_linear(VkvgPattern pat, float x0, float y0, float x1, float y1) {
    if (vkvg_pattern_status(pat)) return vkvg_pattern_status(pat);
    if (pat->type != VKVG_PATTERN_TYPE_LINEAR) return VKVG_STATUS_PATTERN_TYPE_MISMATCH;
    pat->grad.cp[0][0] = x0; pat->grad.cp[0][1] = y0; pat->grad.cp[0][2] = x1; pat->grad.cp[0][3] = y1;
    return VKVG_STATUS_SUCCESS;
}
VkvgPattern vkvg_pattern_create_linear(float x0, float y0, float x1, float y1) {
    VkvgPattern pat = new_pattern(VKVG_PATTERN_TYPE_LINEAR);
    vkvg_pattern_edit_linear(pat, x0, y0, x1, y1);
    return pat;
}
vkvg_status_t vkvg_pattern_get_linear_points(VkvgPattern pat, float *x0, float *y0, float *x1, float *y1) {
    if (vkvg_pattern_status(pat)) return vkvg_pattern_status(pat);
    if (pat->type != VKVG_PATTERN_TYPE_LINEAR) return VKVG_STATUS_PATTERN_TYPE_MISMATCH;
    *x0 = pat->grad.cp[0][0]; *y0 = pat->grad.cp[0][1]; *x1 = pat->grad.cp[0][2]; *y1 = pat->grad.cp[0][3];
    return VKVG_STATUS_SUCCESS;
}
vkvg_status_t vkvg_pattern_edit_radial(VkvgPattern pat, float cx0, float cy0, float radius0, float cx1, float cy1, float radius1) {  // :95-118
    if (vkvg_pattern_status(pat)) return vkvg_pattern_status(pat);
    if (pat->type != VKVG_PATTERN_TYPE_RADIAL) return VKVG_STATUS_PATTERN_TYPE_MISMATCH;
    float c0x = cx0, c0y = cy0;
    if (radius0 > radius1 - 1.0f) radius0 = radius1 - 1.0f;
    float ux = c0x - cx1, uy = c0y - cy1;
    float l  = sqrtf(ux * ux + uy * uy);
    if (l + radius0 + 1.0f >= radius1) {
        float vx = ux / l, vy = uy / l, m = radius1 - radius0 - 1.0f;
        c0x = cx1 + vx * m; c0y = cy1 + vy * m;
    }
    pat->grad.cp[0][0] = c0x; pat->grad.cp[0][1] = c0y; pat->grad.cp[0][2] = radius0; pat->grad.cp[0][3] = 0;
    pat->grad.cp[1][0] = cx1; pat->grad.cp[1][1] = cy1; pat->grad.cp[1][2] = radius1; pat->grad.cp[1][3] = 0;
    return VKVG_STATUS_SUCCESS;
}
VkvgPattern vkvg_pattern_create_radial(float cx0, float cy0, float radius0, float cx1, float cy1, float radius1) {
    VkvgPattern pat = new_pattern(VKVG_PATTERN_TYPE_RADIAL);
    vkvg_pattern_edit_radial(pat, cx0, cy0, radius0, cx1, cy1, radius1);
    return pat;
}
VkvgPattern vkvg_pattern_reference(VkvgPattern pat) {
    if (!vkvg_pattern_status(pat)) pat->references++;
    return pat;
}
uint32_t vkvg_pattern_get_reference_count(VkvgPattern pat) { return vkvg_pattern_status(pat) ? 0u : pat->references.load(); }
void     vkvg_pattern_destroy(VkvgPattern pat) {
    if (vkvg_pattern_status(pat)) return;
    if (--pat->references > 0) return;
    if (pat->type == VKVG_PATTERN_TYPE_SURFACE && pat->surf) vkvg_surface_destroy(pat->surf);
    delete pat;
}
VkvgPattern vkvg_pattern_create_for_surface(VkvgSurface surf) {  // src/vkvg_pattern.c:28-50
    if (!surf) return (VkvgPattern)&s_null_pointer;
    VkvgPattern pat = new_pattern(VKVG_PATTERN_TYPE_SURFACE);
    pat->surf = surf;
    vkvg_surface_reference(surf);
    if (vkvg_surface_status(surf)) { pat->status = VKVG_STATUS_INVALID_SURFACE; pat->surf = NULL; }
    return pat;
}
vkvg_status_t vkvg_pattern_add_color_stop(VkvgPattern pat, float offset, float r, float g, float b, float a) {  // :149-167
    if (vkvg_pattern_status(pat)) return vkvg_pattern_status(pat);
    if (pat->type == VKVG_PATTERN_TYPE_SURFACE || pat->type == VKVG_PATTERN_TYPE_SOLID) return VKVG_STATUS_PATTERN_TYPE_MISMATCH;
    if (pat->grad.count >= 16) return VKVG_STATUS_INVALID_INDEX;  // the reference writes past its 16-entry arrays here
    uint32_t i = pat->grad.count++;
    pat->grad.colors[i][0] = a * r; pat->grad.colors[i][1] = a * g; pat->grad.colors[i][2] = a * b; pat->grad.colors[i][3] = a;
    pat->grad.stops[i] = offset;
    return VKVG_STATUS_SUCCESS;
}
void vkvg_pattern_set_extend(VkvgPattern pat, vkvg_extend_t extend) { if (!vkvg_pattern_status(pat)) pat->extend = extend; }
void vkvg_pattern_set_filter(VkvgPattern pat, vkvg_filter_t filter) { if (!vkvg_pattern_status(pat)) pat->filter = filter; }
vkvg_extend_t vkvg_pattern_get_extend(VkvgPattern pat) { return vkvg_pattern_status(pat) ? (vkvg_extend_t)0 : pat->extend; }
vkvg_filter_t vkvg_pattern_get_filter(VkvgPattern pat) { return vkvg_pattern_status(pat) ? (vkvg_filter_t)0 : pat->filter; }
vkvg_pattern_type_t vkvg_pattern_get_type(VkvgPattern pat) { return vkvg_pattern_status(pat) ? (vkvg_pattern_type_t)0 : pat->type; }
vkvg_status_t vkvg_pattern_get_color_stop_count(VkvgPattern pat, uint32_t *count) {
    if (vkvg_pattern_status(pat)) return vkvg_pattern_status(pat);
    if (pat->type == VKVG_PATTERN_TYPE_SURFACE || pat->type == VKVG_PATTERN_TYPE_SOLID) return VKVG_STATUS_PATTERN_TYPE_MISMATCH;
    *count = pat->grad.count;
    return VKVG_STATUS_SUCCESS;
}
vkvg_status_t vkvg_pattern_get_color_stop_rgba(VkvgPattern pat, uint32_t index, float *offset, float *r, float *g, float *b, float *a) {
    if (vkvg_pattern_status(pat)) return vkvg_pattern_status(pat);
    if (pat->type == VKVG_PATTERN_TYPE_SURFACE || pat->type == VKVG_PATTERN_TYPE_SOLID) return VKVG_STATUS_PATTERN_TYPE_MISMATCH;
    if (index >= pat->grad.count) return VKVG_STATUS_INVALID_INDEX;
    *offset = pat->grad.stops[index];
    *r = pat->grad.colors[index][0]; *g = pat->grad.colors[index][1]; *b = pat->grad.colors[index][2]; *a = pat->grad.colors[index][3];
    return VKVG_STATUS_SUCCESS;
}
void vkvg_pattern_set_matrix(VkvgPattern pat, const vkvg_matrix_t *matrix) {
    if (vkvg_pattern_status(pat)) return;
    pat->matrix = *matrix; pat->hasMatrix = true;
}
void vkvg_pattern_get_matrix(VkvgPattern pat, vkvg_matrix_t *matrix) {
    if (vkvg_pattern_status(pat)) return;
    if (pat->hasMatrix) *matrix = pat->matrix;
    else vkvg_matrix_init_identity(matrix);
}

// ====================================================================================================
// context — reference src/vkvg_context.c
// ====================================================================================================
static void init_ctx(VkvgContext ctx) {  // _init_ctx :24-61
    ctx->lineWidth = 1.f; ctx->miterLimit = 10.f;
    ctx->op = VKVG_OPERATOR_OVER; ctx->fillRule = VKVG_FILL_RULE_NON_ZERO;
    ctx->cap = VKVG_LINE_CAP_BUTT; ctx->join = VKVG_LINE_JOIN_MITER;
    ctx->opacity = 1.0f;
    vkvg_matrix_init_identity(&ctx->mat);
    ctx->pattern = NULL; ctx->patType = VKB_PAT_SOLID; ctx->grad_slot = -1;
    ctx->curColor = 0xff000000;
    ctx->dashOffset = 0;
    ctx->clear_pending = false;
    ctx->curClipState = CLIP_STATE_NONE;
    ctx->curSavBit = 0;
    ctx->canvas = 0;
    ctx->recording = NULL;
    vkvg_matrix_init_identity(&ctx->matInv);
    ctx->src[0] = ctx->src[1] = ctx->src[2] = ctx->src[3] = 0;
}
static void clear_path(VkvgContext ctx) {  // _clear_path, internal.c:199-206
    ctx->path_first_sp = (uint32_t)ctx->batch.subpaths.size();
    ctx->sp_first_elem = (uint32_t)ctx->batch.elem_hdr.size();
    ctx->sp_points = 0;
    ctx->simpleConvex = false;
}
VkvgContext vkvg_create(VkvgSurface surf) {
    if (vkvg_surface_status(surf)) return (VkvgContext)&s_invalid_surface;
    if (vkvg_device_status(surf->dev)) return (VkvgContext)&s_device_error;
    VkvgContext ctx = new _vkvg_context_t();
    ctx->status     = VKVG_STATUS_SUCCESS;
    ctx->references = 1;
    ctx->dev = surf->dev; ctx->pSurf = surf;
    init_ctx(ctx);
    vkvg_surface_reference(surf);
    clear_path(ctx);
    {   // the first render pass of a new context clears the stencil attachment (src/vkvg_context.c:44-49)
        std::lock_guard<std::mutex> lk(surf->dev->mtx);
        vkb_surface_stencil_reset(surf->impl);
    }
    return ctx;
}
vkvg_status_t vkvg_status(VkvgContext ctx) { return !ctx ? VKVG_STATUS_NULL_POINTER : ctx->status; }
VkvgContext   vkvg_reference(VkvgContext ctx) {
    if (!vkvg_status(ctx)) ctx->references++;
    return ctx;
}
uint32_t vkvg_get_reference_count(VkvgContext ctx) { return vkvg_status(ctx) ? 0u : ctx->references.load(); }

// ---- path bookkeeping ----
static inline bool path_empty(VkvgContext ctx) { return ctx->sp_points == 0; }  // _current_path_is_empty, internal.c:132
static void push_elem(VkvgContext ctx, uint32_t type_flags, const float *payload, int n) {
    vkb_batch &b = ctx->batch;
    b.elem_hdr.push_back(type_flags | ((uint32_t)b.elem_data.size() << VKB_EL_PAYLOAD_SHIFT));
    b.elem_data.append(payload, payload + n);
    if ((type_flags & VKB_EL_TYPE_MASK) != VKB_EL_POINT) b.n_curves++;
}
static inline void add_point(VkvgContext ctx, float x, float y, bool curved) {  // _add_point, internal.c:221-238
    if (isnan(x) || isnan(y)) return;
    vkb_batch &b = ctx->batch;
    b.elem_hdr.push_back(VKB_EL_POINT | (curved ? VKB_EL_CURVED : 0) | ((uint32_t)b.elem_data.size() << VKB_EL_PAYLOAD_SHIFT));
    b.elem_data.push_back(x);
    b.elem_data.push_back(y);
    if (ctx->sp_points == 0) { ctx->first_x = x; ctx->first_y = y; }
    ctx->cur_x = x; ctx->cur_y = y;
    ctx->sp_points++;
}
static void finish_path(VkvgContext ctx, uint32_t flags = 0);
// move_to(p[0]) followed by line_to(p[1..n-1]) with the same NaN / duplicate-point rules, written in bulk
static void add_polyline(VkvgContext ctx, const float *p, uint64_t n) {
    if (!n) return;
    finish_path(ctx, 0);
    vkb_batch &b = ctx->batch;
    const size_t h0 = b.elem_hdr.size(), d0 = b.elem_data.size();
    b.elem_hdr.resize(h0 + n);
    b.elem_data.resize(d0 + 2 * n);
    uint32_t *hdr = b.elem_hdr.data() + h0;
    float    *dat = b.elem_data.data() + d0;
    uint64_t  k = 0;
    float     cx = 0, cy = 0;
    for (uint64_t j = 0; j < n; j++) {
        const float x = p[2 * j], y = p[2 * j + 1];
        if (isnan(x) || isnan(y)) continue;                       // _add_point
        if (k && j && EQUF(cx, x) && EQUF(cy, y)) continue;         // _line_to (the first call is move_to: no test)
        hdr[k] = VKB_EL_POINT | ((uint32_t)(d0 + 2 * k) << VKB_EL_PAYLOAD_SHIFT);
        dat[2 * k] = x; dat[2 * k + 1] = y;
        cx = x; cy = y;
        k++;
    }
    b.elem_hdr.resize(h0 + k);
    b.elem_data.resize(d0 + 2 * k);
    if (k) {
        ctx->first_x = dat[0]; ctx->first_y = dat[1];
        ctx->cur_x = cx; ctx->cur_y = cy;
        ctx->sp_points = (uint32_t)k;
        ctx->simpleConvex = false;
    }
}
static void end_subpath(VkvgContext ctx, uint32_t flags) {
    vkb_subpath sp = {ctx->sp_first_elem, (uint32_t)ctx->batch.elem_hdr.size() - ctx->sp_first_elem, flags, 0};
    ctx->batch.subpaths.push_back(sp);
    ctx->sp_first_elem = (uint32_t)ctx->batch.elem_hdr.size();
    ctx->sp_points = 0;
    ctx->simpleConvex = false;
}
static void finish_path(VkvgContext ctx, uint32_t flags) {  // _finish_path, internal.c:163-197
    if (ctx->sp_points == 0) return;
    if (ctx->sp_points < 2) {  // only the current position is in the path: drop it
        ctx->batch.elem_data.resize(ctx->batch.elem_hdr[ctx->sp_first_elem] >> VKB_EL_PAYLOAD_SHIFT);
        ctx->batch.elem_hdr.resize(ctx->sp_first_elem);
        ctx->sp_points = 0;
        return;
    }
    if (ctx->path_first_sp == ctx->batch.subpaths.size() && ctx->simpleConvex) flags |= VKB_SP_CONVEX;
    end_subpath(ctx, flags);
}
static void line_to_(VkvgContext ctx, float x, float y) {  // _line_to, internal.c:1463-1472
    if (!path_empty(ctx) && EQUF(ctx->cur_x, x) && EQUF(ctx->cur_y, y)) return;
    add_point(ctx, x, y, false);
    ctx->simpleConvex = false;
}
void vkvg_new_sub_path(VkvgContext ctx) { if (!vkvg_status(ctx) && rec(ctx, RC_NEW_SUB_PATH)) return; if (!vkvg_status(ctx)) finish_path(ctx); }
void vkvg_new_path(VkvgContext ctx) {
    if (vkvg_status(ctx)) return;
    if (rec(ctx, RC_NEW_PATH)) return;
    // drop the elements of an unfinished sub-path; finished sub-paths may be referenced by recorded draws
    if (ctx->sp_points) {
        ctx->batch.elem_data.resize(ctx->batch.elem_hdr[ctx->sp_first_elem] >> VKB_EL_PAYLOAD_SHIFT);
        ctx->batch.elem_hdr.resize(ctx->sp_first_elem);
    }
    clear_path(ctx);
}
void vkvg_close_path(VkvgContext ctx) {  // :350-373
    if (vkvg_status(ctx)) return;
    if (rec(ctx, RC_CLOSE_PATH)) return;
    if (ctx->sp_points < 3) return;
    uint32_t flags = VKB_SP_CLOSED;
    if (EQUF(ctx->cur_x, ctx->first_x) && EQUF(ctx->cur_y, ctx->first_y)) {
        if (ctx->sp_points < 4) return;
        flags |= VKB_SP_DROP_LAST;  // _remove_last_point
    }
    finish_path(ctx, flags);
}
void vkvg_move_to(VkvgContext ctx, float x, float y) {  // :515-522
    if (vkvg_status(ctx)) return;
    if (rec(ctx, RC_MOVE_TO, {x, y})) return;
    finish_path(ctx);
    add_point(ctx, x, y, false);
}
void vkvg_rel_move_to(VkvgContext ctx, float x, float y) {  // :504-514
    if (vkvg_status(ctx)) return;
    if (rec(ctx, RC_REL_MOVE_TO, {x, y})) return;
    if (path_empty(ctx)) add_point(ctx, 0, 0, false);
    float cx = ctx->cur_x, cy = ctx->cur_y;
    finish_path(ctx);
    add_point(ctx, cx + x, cy + y, false);
}
void vkvg_line_to(VkvgContext ctx, float x, float y) { if (!vkvg_status(ctx) && !rec(ctx, RC_LINE_TO, {x, y})) line_to_(ctx, x, y); }
void vkvg_rel_line_to(VkvgContext ctx, float dx, float dy) {  // :374-385
    if (vkvg_status(ctx)) return;
    if (rec(ctx, RC_REL_LINE_TO, {dx, dy})) return;
    if (path_empty(ctx)) add_point(ctx, 0, 0, false);
    line_to_(ctx, ctx->cur_x + dx, ctx->cur_y + dy);
}
bool vkvg_has_current_point(VkvgContext ctx) { return vkvg_status(ctx) ? false : !path_empty(ctx); }
void vkvg_get_current_point(VkvgContext ctx, float *x, float *y) {  // :528-540
    if (vkvg_status(ctx)) return;
    if (path_empty(ctx)) { *x = *y = 0; return; }
    *x = ctx->cur_x; *y = ctx->cur_y;
}
static float get_arc_step(VkvgContext ctx, float radius) {  // _get_arc_step, internal.c:245-252
    float sx, sy;
    vkvg_matrix_get_scale(&ctx->mat, &sx, &sy);
    float r = radius * fabsf(fmaxf(sx, sy));
    if (r < 30.0f) return fminf(M_PIF / 3.f, M_PIF / r);
    return fminf(M_PIF / 3.f, M_PIF / (r * 0.4f));
}
static void arc_impl(VkvgContext ctx, float xc, float yc, float radius, float a1, float a2, bool negative) {  // :394-503
    if (!negative) {
        while (a2 < a1) a2 += 2.f * M_PIF;
        if (a2 - a1 > 2.f * M_PIF) a2 = a1 + 2.f * M_PIF;
    } else {
        while (a2 > a1) a2 -= 2.f * M_PIF;
        if (a1 - a2 > a1 + 2.f * M_PIF) a2 = a1 - 2.f * M_PIF;
    }
    float vx = cosf(a1) * radius + xc, vy = sinf(a1) * radius + yc;
    float step = get_arc_step(ctx, radius);
    float a    = a1;
    if (path_empty(ctx)) {
        add_point(ctx, vx, vy, true);  // inside _set_curve_start: the start point belongs to the curved segment
        ctx->simpleConvex = ctx->path_first_sp == ctx->batch.subpaths.size();
    } else {
        line_to_(ctx, vx, vy);
        ctx->simpleConvex = false;
    }
    if (!negative) a += step; else a -= step;
    if (EQUF(a2, a1)) return;
    // interior points are generated on the device; the host only needs their number for the 3/4-point tests
    uint32_t n = 0;
    float    t = a;
    if (!negative) while (t < a2) { n++; t += step; }
    else while (t > a2) { n++; t -= step; }
    if (n) {
        float p[6] = {xc, yc, radius, a, a2, negative ? -step : step};
        push_elem(ctx, VKB_EL_ARC | VKB_EL_CURVED, p, 6);
        ctx->sp_points += n;
        // current point after the interior points (only observable through the full-circle close test below)
        float la = negative ? t + step : t - step;
        ctx->cur_x = cosf(la) * radius + xc; ctx->cur_y = sinf(la) * radius + yc;
    }
    if (EQUF(negative ? a1 - a2 : a2 - a1, M_PIF * 2.f)) {  // complete circle: last point == first one
        vkvg_close_path(ctx);
        return;
    }
    add_point(ctx, cosf(a2) * radius + xc, sinf(a2) * radius + yc, true);
}
void vkvg_arc(VkvgContext ctx, float xc, float yc, float radius, float a1, float a2) { if (!vkvg_status(ctx) && !rec(ctx, RC_ARC, {xc, yc, radius, a1, a2})) arc_impl(ctx, xc, yc, radius, a1, a2, false); }
void vkvg_arc_negative(VkvgContext ctx, float xc, float yc, float radius, float a1, float a2) { if (!vkvg_status(ctx) && !rec(ctx, RC_ARC_NEG, {xc, yc, radius, a1, a2})) arc_impl(ctx, xc, yc, radius, a1, a2, true); }

static void curve_to_(VkvgContext ctx, float x1, float y1, float x2, float y2, float x3, float y3) {  // _curve_to :541-566
    if (EQUF(x1, x2) && EQUF(x2, x3) && EQUF(y1, y2) && EQUF(y2, y3)) {
        if (path_empty(ctx) || (EQUF(ctx->cur_x, x1) && EQUF(ctx->cur_y, y1))) return;
    }
    if (!(isfinite(x1) && isfinite(y1) && isfinite(x2) && isfinite(y2) && isfinite(x3) && isfinite(y3))) return;  // the reference recursion does not terminate on these
    ctx->simpleConvex = false;
    if (path_empty(ctx)) add_point(ctx, x1, y1, true);
    float sx = 1, sy = 1;
    vkvg_matrix_get_scale(&ctx->mat, &sx, &sy);
    float tol  = fabs(0.25f / fmaxf(sx, sy));
    float p[9] = {ctx->cur_x, ctx->cur_y, x1, y1, x2, y2, x3, y3, tol};
    push_elem(ctx, VKB_EL_CUBIC | VKB_EL_CURVED, p, 9);
    ctx->sp_points += 3;  // >= 2 recursion points (level 0 always splits) + the end point
    ctx->cur_x = x3; ctx->cur_y = y3;
}
void vkvg_curve_to(VkvgContext ctx, float x1, float y1, float x2, float y2, float x3, float y3) { if (!vkvg_status(ctx) && !rec(ctx, RC_CURVE_TO, {x1, y1, x2, y2, x3, y3})) curve_to_(ctx, x1, y1, x2, y2, x3, y3); }
void vkvg_rel_curve_to(VkvgContext ctx, float x1, float y1, float x2, float y2, float x3, float y3) {  // :600-611
    if (vkvg_status(ctx)) return;
    if (rec(ctx, RC_REL_CURVE_TO, {x1, y1, x2, y2, x3, y3})) return;
    if (path_empty(ctx)) { ctx->status = VKVG_STATUS_NO_CURRENT_POINT; return; }
    float cx = ctx->cur_x, cy = ctx->cur_y;
    curve_to_(ctx, cx + x1, cy + y1, cx + x2, cy + y2, cx + x3, cy + y3);
}
static void quadratic_to_(VkvgContext ctx, float x1, float y1, float x2, float y2) {  // _quadratic_to :567-577
    const double qf = 2.0 / 3.0;
    float x0, y0;
    if (path_empty(ctx)) { x0 = x1; y0 = y1; }
    else { x0 = ctx->cur_x; y0 = ctx->cur_y; }
    curve_to_(ctx, x0 + (x1 - x0) * qf, y0 + (y1 - y0) * qf, x2 + (x1 - x2) * qf, y2 + (y1 - y2) * qf, x2, y2);
}
void vkvg_quadratic_to(VkvgContext ctx, float x1, float y1, float x2, float y2) { if (!vkvg_status(ctx) && !rec(ctx, RC_QUADRATIC_TO, {x1, y1, x2, y2})) quadratic_to_(ctx, x1, y1, x2, y2); }
void vkvg_rel_quadratic_to(VkvgContext ctx, float x1, float y1, float x2, float y2) {
    if (vkvg_status(ctx)) return;
    if (rec(ctx, RC_REL_QUADRATIC_TO, {x1, y1, x2, y2})) return;
    float cx = ctx->cur_x, cy = ctx->cur_y;
    quadratic_to_(ctx, cx + x1, cy + y1, cx + x2, cy + y2);
}
vkvg_status_t vkvg_rectangle(VkvgContext ctx, float x, float y, float w, float h) {  // :620-639
    if (vkvg_status(ctx)) return ctx->status;
    if (rec(ctx, RC_RECTANGLE, {x, y, w, h})) return VKVG_STATUS_SUCCESS;
    finish_path(ctx);
    if (w <= 0 || h <= 0) return VKVG_STATUS_INVALID_RECT;
    add_point(ctx, x, y, false);
    add_point(ctx, x + w, y, false);
    add_point(ctx, x + w, y + h, false);
    add_point(ctx, x, y + h, false);
    finish_path(ctx, VKB_SP_CLOSED | VKB_SP_CONVEX);
    return VKVG_STATUS_SUCCESS;
}
vkvg_status_t vkvg_rounded_rectangle(VkvgContext ctx, float x, float y, float w, float h, float radius) {  // :640-664
    if (vkvg_status(ctx)) return ctx->status;
    finish_path(ctx);
    if (w <= 0 || h <= 0) return VKVG_STATUS_INVALID_RECT;
    if ((radius > w / 2.0f) || (radius > h / 2.0f)) radius = fmin(w / 2.0f, h / 2.0f);
    vkvg_move_to(ctx, x, y + radius);
    vkvg_arc(ctx, x + radius, y + radius, radius, M_PIF, -M_PIF_2);
    vkvg_line_to(ctx, x + w - radius, y);
    vkvg_arc(ctx, x + w - radius, y + radius, radius, -M_PIF_2, 0);
    vkvg_line_to(ctx, x + w, y + h - radius);
    vkvg_arc(ctx, x + w - radius, y + h - radius, radius, 0, M_PIF_2);
    vkvg_line_to(ctx, x + radius, y + h);
    vkvg_arc(ctx, x + radius, y + h - radius, radius, M_PIF_2, M_PIF);
    vkvg_line_to(ctx, x, y + radius);
    vkvg_close_path(ctx);
    return VKVG_STATUS_SUCCESS;
}
void vkvg_ellipse(VkvgContext ctx, float radiusX, float radiusY, float x, float y, float rotationAngle) {  // :1605-1639
    if (vkvg_status(ctx)) return;
    if (ctx->recording) {  // stored as the path calls it stands for
        float w23 = radiusX * 4 / 3, dx1 = sinf(rotationAngle) * radiusY, dy1 = cosf(rotationAngle) * radiusY;
        float dx2 = cosf(rotationAngle) * w23, dy2 = sinf(rotationAngle) * w23, tcx = x - dx1, tcy = y + dy1, bcx = x + dx1, bcy = y - dy1;
        vkvg_move_to(ctx, bcx, bcy);
        vkvg_curve_to(ctx, bcx + dx2, bcy + dy2, tcx + dx2, tcy + dy2, tcx, tcy);
        vkvg_curve_to(ctx, tcx - dx2, tcy - dy2, bcx - dx2, bcy - dy2, bcx, bcy);
        vkvg_close_path(ctx);
        return;
    }
    float w23 = radiusX * 4 / 3;
    float dx1 = sinf(rotationAngle) * radiusY, dy1 = cosf(rotationAngle) * radiusY;
    float dx2 = cosf(rotationAngle) * w23, dy2 = sinf(rotationAngle) * w23;
    float tcx = x - dx1, tcy = y + dy1, bcx = x + dx1, bcy = y - dy1;
    finish_path(ctx);
    add_point(ctx, bcx, bcy, false);
    curve_to_(ctx, bcx + dx2, bcy + dy2, tcx + dx2, tcy + dy2, tcx, tcy);
    curve_to_(ctx, tcx - dx2, tcy - dy2, bcx - dx2, bcy - dy2, bcx, bcy);
    finish_path(ctx, VKB_SP_CLOSED);
}
// SVG-style elliptical arc, flattened on the host at a tenth of the circular arc step like the reference
// (_elliptic_arc, src/vkvg_context_internal.c:1473-1580; float / double mix kept as it is there)
static void elliptic_arc(VkvgContext ctx, float x1, float y1, float x2, float y2, bool largeArc, bool counterClockWise, float _rx, float _ry, float phi) {
    if (_rx == 0 || _ry == 0) {
        if (path_empty(ctx)) vkvg_move_to(ctx, x1, y1);
        vkvg_line_to(ctx, x2, y2);
        return;
    }
    float rx = fabsf(_rx), ry = fabsf(_ry);
    const float cphi = cosf(phi), sphi = sinf(phi);
    // midpoint vector rotated into the ellipse frame
    float hx = (x1 - x2) / 2, hy = (y1 - y2) / 2;
    float p1x = (cphi * hx) + (sphi * hy), p1y = (-sphi * hx) + (cphi * hy);
    double lambda = powf(p1x, 2) / powf(rx, 2) + powf(p1y, 2) / powf(ry, 2);
    if (lambda > 1) {  // radii too small for the chord: scale them up
        lambda = sqrtf(lambda);
        rx *= lambda;
        ry *= lambda;
    }
    float qx = rx * p1y / ry, qy = -ry * p1x / rx;
    float k  = sqrtf(fabsf((powf(rx, 2) * powf(ry, 2) - powf(rx, 2) * powf(p1y, 2) - powf(ry, 2) * powf(p1x, 2)) /
                           (powf(rx, 2) * powf(p1y, 2) + powf(ry, 2) * powf(p1x, 2))));
    float cpx = qx * k, cpy = qy * k;
    if (largeArc == counterClockWise) { cpx = -cpx; cpy = -cpy; }
    // centre back in user space
    float mx = (x1 + x2) / 2, my = (y1 + y2) / 2;
    float cx = ((cphi * cpx) + (-sphi * cpy)) + mx, cy = ((sphi * cpx) + (cphi * cpy)) + my;
    auto angle = [](float ux, float uy, float vx, float vy) {
        double a = acosf(((ux * vx) + (uy * vy)) / (fabsf(sqrtf(vx * vx + vy * vy)) * fabsf(sqrtf(ux * ux + uy * uy))));
        if (isnan(a)) a = M_PIF;
        if (ux * vy - uy * vx < 0) a = -a;
        return a;
    };
    float  ux = 1.f, uy = 0, vx = (p1x - cpx) / rx, vy = (p1y - cpy) / ry;
    double sa = angle(ux, uy, vx, vy);
    ux = vx; uy = vy;
    vx = (-p1x - cpx) / rx; vy = (-p1y - cpy) / ry;
    double delta_theta = angle(ux, uy, vx, vy);
    if (counterClockWise) {
        if (delta_theta < 0) delta_theta += M_PIF * 2.0;
    } else if (delta_theta > 0)
        delta_theta -= M_PIF * 2.0;
    double theta = sa, ea = sa + delta_theta;
    float  step  = fmaxf(0.001f, fminf(M_PIF, get_arc_step(ctx, fminf(rx, ry)) * 0.1f));
    auto point_at = [&](double t, float &x, float &y) {
        float ex = rx * cosf(t), ey = ry * sinf(t);
        x = ((cphi * ex) + (-sphi * ey)) + cx;
        y = ((sphi * ex) + (cphi * ey)) + cy;
    };
    float x, y;
    point_at(theta, x, y);
    if (path_empty(ctx)) {
        add_point(ctx, x, y, true);
        ctx->simpleConvex = ctx->path_first_sp == ctx->batch.subpaths.size();
    } else {
        line_to_(ctx, x, y);
        ctx->simpleConvex = false;
    }
    if (sa < ea) {
        theta += step;
        while (theta < ea) { point_at(theta, x, y); add_point(ctx, x, y, true); theta += step; }
    } else {
        theta -= step;
        while (theta > ea) { point_at(theta, x, y); add_point(ctx, x, y, true); theta -= step; }
    }
    point_at(ea, x, y);
    add_point(ctx, x, y, true);
}
void vkvg_elliptic_arc_to(VkvgContext ctx, float x2, float y2, bool largeArc, bool sweepFlag, float rx, float ry, float phi) {  // :1579-1590
    if (vkvg_status(ctx)) return;
    if (rec(ctx, RC_ELLIPTICAL_ARC_TO, {x2, y2, rx, ry, phi}, {largeArc ? 1u : 0u, sweepFlag ? 1u : 0u})) return;
    float x1 = 0, y1 = 0;
    vkvg_get_current_point(ctx, &x1, &y1);
    elliptic_arc(ctx, x1, y1, x2, y2, largeArc, sweepFlag, rx, ry, phi);
}
void vkvg_rel_elliptic_arc_to(VkvgContext ctx, float x2, float y2, bool largeArc, bool sweepFlag, float rx, float ry, float phi) {  // :1591-1603
    if (vkvg_status(ctx)) return;
    if (rec(ctx, RC_REL_ELLIPTICAL_ARC_TO, {x2, y2, rx, ry, phi}, {largeArc ? 1u : 0u, sweepFlag ? 1u : 0u})) return;
    float x1 = 0, y1 = 0;
    vkvg_get_current_point(ctx, &x1, &y1);
    elliptic_arc(ctx, x1, y1, x2 + x1, y2 + y1, largeArc, sweepFlag, rx, ry, phi);
}
void vkvg_rounded_rectangle2(VkvgContext ctx, float x, float y, float w, float h, float rx, float ry) {  // :665-683
    if (vkvg_status(ctx)) return;
    vkvg_move_to(ctx, x + rx, y);
    vkvg_line_to(ctx, x + w - rx, y);
    vkvg_elliptic_arc_to(ctx, x + w, y + ry, false, true, rx, ry, 0);
    vkvg_line_to(ctx, x + w, y + h - ry);
    vkvg_elliptic_arc_to(ctx, x + w - rx, y + h, false, true, rx, ry, 0);
    vkvg_line_to(ctx, x + rx, y + h);
    vkvg_elliptic_arc_to(ctx, x, y + h - ry, false, true, rx, ry, 0);
    vkvg_line_to(ctx, x, y + ry);
    vkvg_elliptic_arc_to(ctx, x + rx, y, false, true, rx, ry, 0);
    vkvg_close_path(ctx);
}

// ---- state ----
void  vkvg_set_opacity(VkvgContext ctx, float opacity) { if (!vkvg_status(ctx)) ctx->opacity = opacity; }
float vkvg_get_opacity(VkvgContext ctx) { return vkvg_status(ctx) ? 0 : ctx->opacity; }
static uint32_t rgbaf(float r, float g, float b, float a) {  // CreateRgbaf, internal.h:60-62
    return (((uint32_t)(a * 255.0f) & 0xFF) << 24) | (((uint32_t)(b * a * 255.0f) & 0xFF) << 16) | (((uint32_t)(g * a * 255.0f) & 0xFF) << 8) |
           ((uint32_t)(r * a * 255.0f) & 0xFF);
}
static void set_solid(VkvgContext ctx, uint32_t c) {  // _update_cur_pattern(ctx, NULL), internal.c:682-832
    ctx->curColor = c;
    if (ctx->pattern) vkvg_pattern_destroy(ctx->pattern);
    ctx->pattern = NULL; ctx->patType = VKB_PAT_SOLID; ctx->grad_slot = -1;
}
void vkvg_set_source_color(VkvgContext ctx, uint32_t c) { if (!vkvg_status(ctx) && !rec(ctx, RC_SET_SOURCE_COLOR, {}, {c})) set_solid(ctx, c); }
// include/vkvg.h:1958 of the reference declares this without defining it anywhere in src/: "#rgb", "#rrggbb", "#rrggbbaa" and the
// sixteen CSS basic colour names are understood here; anything else leaves the source unchanged
void vkvg_set_source_color_name(VkvgContext ctx, const char *color) {
    if (vkvg_status(ctx) || !color) return;
    static const struct { const char *name; uint32_t rgb; } names[] = {
        {"black", 0x000000}, {"silver", 0xc0c0c0}, {"gray", 0x808080}, {"grey", 0x808080}, {"white", 0xffffff}, {"maroon", 0x800000}, {"red", 0xff0000},
        {"purple", 0x800080}, {"fuchsia", 0xff00ff}, {"magenta", 0xff00ff}, {"green", 0x008000}, {"lime", 0x00ff00}, {"olive", 0x808000}, {"yellow", 0xffff00},
        {"navy", 0x000080}, {"blue", 0x0000ff}, {"teal", 0x008080}, {"aqua", 0x00ffff}, {"cyan", 0x00ffff}, {"orange", 0xffa500}};
    uint32_t rgb = 0, alpha = 255;
    if (color[0] == '#') {
        const size_t n = strlen(color + 1);
        char        *end = nullptr;
        const unsigned long v = strtoul(color + 1, &end, 16);
        if (!end || *end) return;
        if (n == 3) rgb = (uint32_t)(((v >> 8) & 0xF) * 0x110000 + ((v >> 4) & 0xF) * 0x1100 + (v & 0xF) * 0x11);
        else if (n == 6) rgb = (uint32_t)v;
        else if (n == 8) { rgb = (uint32_t)(v >> 8); alpha = (uint32_t)(v & 0xFF); }
        else return;
    } else {
        bool found = false;
        for (const auto &e : names)
            if (!strcasecmp(e.name, color)) { rgb = e.rgb; found = true; break; }
        if (!found) return;
    }
    vkvg_set_source_rgba(ctx, ((rgb >> 16) & 0xFF) / 255.0f, ((rgb >> 8) & 0xFF) / 255.0f, (rgb & 0xFF) / 255.0f, alpha / 255.0f);
}
void vkvg_set_source_rgb(VkvgContext ctx, float r, float g, float b) { if (!vkvg_status(ctx) && !rec(ctx, RC_SET_SOURCE_RGB, {r, g, b})) set_solid(ctx, rgbaf(r, g, b, 1)); }
void vkvg_set_source_rgba(VkvgContext ctx, float r, float g, float b, float a) { if (!vkvg_status(ctx) && !rec(ctx, RC_SET_SOURCE_RGBA, {r, g, b, a})) set_solid(ctx, rgbaf(r, g, b, a)); }
static void update_cur_pattern(VkvgContext ctx, VkvgPattern pat) {  // surface branch internal.c:705-773, gradient branch :774-826
    VkvgPattern last = ctx->pattern;
    ctx->pattern     = pat;
    if (pat->type == VKVG_PATTERN_TYPE_SURFACE) {
        ctx->src[2] = (float)pat->surf->width;
        ctx->src[3] = (float)pat->surf->height;
        if (pat->hasMatrix) {  // :762-770: folded into matInv until the next CTM change recomputes it
            vkvg_matrix_t m = pat->matrix;
            vkvg_matrix_multiply(&ctx->matInv, &ctx->matInv, &m);
        }
        ctx->patType = VKB_PAT_SURFACE; ctx->grad_slot = -1;
        if (last) vkvg_pattern_destroy(last);
        return;
    }
    vkb_gradient g   = pat->grad;
    if (g.count < 2) {
        ctx->status = VKVG_STATUS_PATTERN_INVALID_GRADIENT;
        return;
    }
    vkvg_matrix_t mat;
    if (pat->hasMatrix) {
        mat = pat->matrix;
        if (vkvg_matrix_invert(&mat) != VKVG_STATUS_SUCCESS) vkvg_matrix_init_identity(&mat);
        vkvg_matrix_transform_point(&mat, &g.cp[0][0], &g.cp[0][1]);
    }
    vkvg_matrix_transform_point(&ctx->mat, &g.cp[0][0], &g.cp[0][1]);
    if (pat->type == VKVG_PATTERN_TYPE_LINEAR) {
        if (pat->hasMatrix) vkvg_matrix_transform_point(&mat, &g.cp[0][2], &g.cp[0][3]);
        vkvg_matrix_transform_point(&ctx->mat, &g.cp[0][2], &g.cp[0][3]);
    } else {
        if (pat->hasMatrix) vkvg_matrix_transform_point(&mat, &g.cp[1][0], &g.cp[1][1]);
        vkvg_matrix_transform_point(&ctx->mat, &g.cp[1][0], &g.cp[1][1]);
        if (pat->hasMatrix) {
            vkvg_matrix_transform_distance(&mat, &g.cp[0][2], &g.cp[0][3]);
            vkvg_matrix_transform_distance(&mat, &g.cp[1][2], &g.cp[0][3]);
        }
        vkvg_matrix_transform_distance(&ctx->mat, &g.cp[0][2], &g.cp[0][3]);
        vkvg_matrix_transform_distance(&ctx->mat, &g.cp[1][2], &g.cp[0][3]);
    }
    ctx->grad = g; ctx->grad_slot = -1;
    ctx->patType = pat->type == VKVG_PATTERN_TYPE_LINEAR ? VKB_PAT_LINEAR : VKB_PAT_RADIAL;
    if (last) vkvg_pattern_destroy(last);
}
void vkvg_set_source(VkvgContext ctx, VkvgPattern pat) {  // :1034-1040
    if (vkvg_status(ctx) || vkvg_pattern_status(pat)) return;
    if (pat->type != VKVG_PATTERN_TYPE_LINEAR && pat->type != VKVG_PATTERN_TYPE_RADIAL && pat->type != VKVG_PATTERN_TYPE_SURFACE) return;
    if (ctx->recording) {  // the recording keeps the pattern alive (record :79-85)
        ctx->recording->pats.push_back(vkvg_pattern_reference(pat));
        rec(ctx, RC_SET_SOURCE, {}, {(uint32_t)ctx->recording->pats.size() - 1});
        return;
    }
    update_cur_pattern(ctx, pat);
    vkvg_pattern_reference(pat);
}
void vkvg_set_source_surface(VkvgContext ctx, VkvgSurface surf, float x, float y) {  // :1025-1033
    if (vkvg_status(ctx) || vkvg_surface_status(surf)) return;
    if (ctx->recording) {
        ctx->recording->surfs.push_back(vkvg_surface_reference(surf));
        rec(ctx, RC_SET_SOURCE_SURFACE, {x, y}, {(uint32_t)ctx->recording->surfs.size() - 1});
        return;
    }
    ctx->src[0] = x; ctx->src[1] = y;
    update_cur_pattern(ctx, vkvg_pattern_create_for_surface(surf));  // the context owns the pattern's only reference
}
VkvgPattern vkvg_get_source(VkvgContext ctx) {
    if (vkvg_status(ctx)) return NULL;
    vkvg_pattern_reference(ctx->pattern);
    return ctx->pattern;
}
VkvgSurface vkvg_get_target(VkvgContext ctx) { return vkvg_status(ctx) ? NULL : ctx->pSurf; }
void  vkvg_set_line_width(VkvgContext ctx, float width) { if (!vkvg_status(ctx) && !rec(ctx, RC_SET_LINE_WIDTH, {width})) ctx->lineWidth = width; }
void  vkvg_set_miter_limit(VkvgContext ctx, float limit) { if (!vkvg_status(ctx) && !rec(ctx, RC_SET_MITER_LIMIT, {limit})) ctx->miterLimit = limit; }  // (the reference records this one under the line-width code: :1050)
void  vkvg_set_line_cap(VkvgContext ctx, vkvg_line_cap_t cap) { if (!vkvg_status(ctx) && !rec(ctx, RC_SET_LINE_CAP, {}, {(uint32_t)cap})) ctx->cap = cap; }
void  vkvg_set_line_join(VkvgContext ctx, vkvg_line_join_t join) { if (!vkvg_status(ctx) && !rec(ctx, RC_SET_LINE_JOIN, {}, {(uint32_t)join})) ctx->join = join; }
void  vkvg_set_operator(VkvgContext ctx, vkvg_operator_t op) { if (!vkvg_status(ctx) && !rec(ctx, RC_SET_OPERATOR, {}, {(uint32_t)op})) ctx->op = op; }  // :1065-1079; takes effect for later draws
void  vkvg_set_fill_rule(VkvgContext ctx, vkvg_fill_rule_t fr) { if (!vkvg_status(ctx) && !rec(ctx, RC_SET_FILL_RULE, {}, {(uint32_t)fr})) ctx->fillRule = fr; }
float vkvg_get_line_width(VkvgContext ctx) { return vkvg_status(ctx) ? 0 : ctx->lineWidth; }
float vkvg_get_miter_limit(VkvgContext ctx) { return vkvg_status(ctx) ? 0 : ctx->miterLimit; }
vkvg_line_cap_t  vkvg_get_line_cap(VkvgContext ctx) { return vkvg_status(ctx) ? (vkvg_line_cap_t)0 : ctx->cap; }
vkvg_line_join_t vkvg_get_line_join(VkvgContext ctx) { return vkvg_status(ctx) ? (vkvg_line_join_t)0 : ctx->join; }
vkvg_operator_t  vkvg_get_operator(VkvgContext ctx) { return vkvg_status(ctx) ? VKVG_OPERATOR_OVER : ctx->op; }
vkvg_fill_rule_t vkvg_get_fill_rule(VkvgContext ctx) { return vkvg_status(ctx) ? VKVG_FILL_RULE_NON_ZERO : ctx->fillRule; }
void vkvg_set_dash(VkvgContext ctx, const float *dashes, uint32_t num_dashes, float offset) {  // :1103-1115
    if (vkvg_status(ctx)) return;
    if (ctx->recording) {
        rec(ctx, RC_SET_DASH, {}, {dashes ? num_dashes : 0u});
        VkvgRecording r = ctx->recording;
        const char *po = (const char *)&offset;
        r->buf.insert(r->buf.end(), po, po + 4);
        if (dashes) r->buf.insert(r->buf.end(), (const char *)dashes, (const char *)(dashes + num_dashes));
        return;
    }
    ctx->dashOffset = offset;
    ctx->dashes.assign(dashes, dashes + (dashes ? num_dashes : 0));
}
void vkvg_get_dash(VkvgContext ctx, const float *dashes, uint32_t *num_dashes, float *offset) {
    if (vkvg_status(ctx)) return;
    *num_dashes = (uint32_t)ctx->dashes.size();
    *offset     = ctx->dashOffset;
    if (ctx->dashes.empty() || dashes == NULL) return;
    memcpy((float *)dashes, ctx->dashes.data(), sizeof(float) * ctx->dashes.size());
}
// ---- clipping: src/vkvg_context.c:698-795.  The stencil work becomes entries of the draw list (VKB_DRAW_CLIP /
//      VKB_DRAW_STENCIL) that the fine pass applies per sample, in submission order with the colour draws. ----
static vkb_draw base_draw(VkvgContext ctx, uint32_t kind, uint32_t rule);
static void     flush_impl(VkvgContext ctx, vkb_capture *cap, bool keep_resident);
static void     reserve_draw_tables(VkvgContext ctx);
static void     finish_path(VkvgContext ctx, uint32_t flags);
static void stencil_op(VkvgContext ctx, uint32_t rule, uint32_t bit) {
    reserve_draw_tables(ctx);
    vkb_draw d = base_draw(ctx, VKB_DRAW_STENCIL, rule);  // (registers the transform entry that carries the canvas of a batch surface)
    d.n_subpaths = 0; d.gradient = 0;
    d.rule_pattern = rule;
    uint32_t sh = 0;
    while (bit >> (sh + 1)) sh++;
    d.color = bit | (sh << 8);
    d.opacity = 1.0f;
    ctx->batch.draws.push_back(d);
}
static int previous_clip_state(VkvgContext ctx) { return ctx->saved.empty() ? CLIP_STATE_CLEAR : ctx->saved.back().clippingState; }  // :698-702
static void clip_preserve_(VkvgContext ctx) {  // _clip_preserve :754-795
    finish_path(ctx, 0);
    if (ctx->batch.subpaths.size() == ctx->path_first_sp) return;  // nothing to clip
    reserve_draw_tables(ctx);
    vkb_draw d = base_draw(ctx, VKB_DRAW_CLIP, ctx->fillRule == VKVG_FILL_RULE_EVEN_ODD ? VKB_RULE_CLIP_EO : VKB_RULE_CLIP_NZ);
    d.rule_pattern &= 0xFF;  // no paint
    d.gradient = 0;
    ctx->batch.draws.push_back(d);
    ctx->curClipState = CLIP_STATE_CLIP;
}
void vkvg_clip_preserve(VkvgContext ctx) { if (!vkvg_status(ctx) && !rec(ctx, RC_CLIP_PRESERVE)) clip_preserve_(ctx); }
void vkvg_clip(VkvgContext ctx) {
    if (vkvg_status(ctx)) return;
    if (rec(ctx, RC_CLIP)) return;
    clip_preserve_(ctx);
    clear_path(ctx);
}
void vkvg_reset_clip(VkvgContext ctx) {  // :719-733; _reset_clip clears the whole stencil attachment, save bits included (:706-717)
    if (vkvg_status(ctx)) return;
    if (rec(ctx, RC_RESET_CLIP)) return;
    if (ctx->curClipState == CLIP_STATE_CLEAR) return;
    ctx->curClipState = previous_clip_state(ctx) == CLIP_STATE_CLEAR ? CLIP_STATE_NONE : CLIP_STATE_CLEAR;
    stencil_op(ctx, VKB_RULE_ST_CLEAR, 0);
}
void vkvg_save(VkvgContext ctx) {  // :1251-1375
    if (vkvg_status(ctx)) return;
    if (rec(ctx, RC_SAVE)) return;
    saved_state s;
    if (ctx->curClipState == CLIP_STATE_CLIP) {
        s.clippingState = CLIP_STATE_CLIP_SAVED;
        if (ctx->curSavBit > 0 && ctx->curSavBit % 6 == 0) {  // the six save bits are taken: park the whole stencil plane (:1268-1318)
            flush_impl(ctx, nullptr, false);
            std::lock_guard<std::mutex> lk(ctx->dev->mtx);
            if (vkb_surface_stencil_push(ctx->pSurf->impl, ctx->dev->raster_samples())) ctx->status = VKVG_STATUS_DEVICE_ERROR;
        }
        stencil_op(ctx, VKB_RULE_ST_SAVE, 1u << (ctx->curSavBit % 6 + 2));
        ctx->curSavBit++;
    } else if (ctx->curClipState == CLIP_STATE_NONE)
        s.clippingState = previous_clip_state(ctx) & 0x03;
    else
        s.clippingState = CLIP_STATE_CLEAR;
    s.lineWidth = ctx->lineWidth; s.miterLimit = ctx->miterLimit; s.dashOffset = ctx->dashOffset; s.dashes = ctx->dashes;
    s.op = ctx->op; s.cap = ctx->cap; s.fillRule = ctx->fillRule; s.opacity = ctx->opacity; s.mat = ctx->mat;
    s.curColor = ctx->curColor; s.pattern = ctx->pattern; s.patType = ctx->patType; s.grad = ctx->grad;
    s.matInv = ctx->matInv; memcpy(s.src, ctx->src, sizeof s.src);
    if (ctx->pattern) vkvg_pattern_reference(ctx->pattern);
    ctx->saved.push_back(s);
}
void vkvg_restore(VkvgContext ctx) {  // :1376-1512
    if (vkvg_status(ctx)) return;
    if (rec(ctx, RC_RESTORE)) return;
    if (ctx->saved.empty()) { ctx->status = VKVG_STATUS_INVALID_RESTORE; return; }
    saved_state s = ctx->saved.back();
    ctx->saved.pop_back();
    if (ctx->curClipState != CLIP_STATE_NONE) {  // :1398-1424
        if (ctx->curClipState == CLIP_STATE_CLIP && s.clippingState == CLIP_STATE_CLEAR) stencil_op(ctx, VKB_RULE_ST_CLEAR, 0);
        else stencil_op(ctx, VKB_RULE_ST_RESTORE, 1u << ((ctx->curSavBit - 1) % 6 + 2));
    }
    if (s.clippingState == CLIP_STATE_CLIP_SAVED) {  // :1425-1470
        ctx->curSavBit--;
        if (ctx->curSavBit > 0 && ctx->curSavBit % 6 == 0) {
            flush_impl(ctx, nullptr, false);
            std::lock_guard<std::mutex> lk(ctx->dev->mtx);
            if (vkb_surface_stencil_pop(ctx->pSurf->impl, ctx->dev->raster_samples())) ctx->status = VKVG_STATUS_DEVICE_ERROR;
        }
    }
    ctx->curClipState = CLIP_STATE_NONE;
    ctx->mat = s.mat; ctx->opacity = s.opacity;
    ctx->matInv = s.matInv; memcpy(ctx->src, s.src, sizeof s.src);  // pushConsts restored wholesale (:1395)
    ctx->dashOffset = s.dashOffset; ctx->dashes = s.dashes;
    ctx->lineWidth = s.lineWidth; ctx->miterLimit = s.miterLimit; ctx->op = s.op; ctx->cap = s.cap;
    ctx->join     = VKVG_LINE_JOIN_MITER;  // the reference never saves lineJoin: restore reads a zeroed field (:1492)
    ctx->fillRule = s.fillRule;
    if (s.pattern) {  // :1501-1509: a different saved pattern is re-uploaded through the restored CTM
        if (s.pattern != ctx->pattern) update_cur_pattern(ctx, s.pattern);
        else vkvg_pattern_destroy(s.pattern);
    } else
        set_solid(ctx, s.curColor);
}
static void set_mat_inv(VkvgContext ctx) {  // _set_mat_inv_and_vkCmdPush, internal.c:672-676 (a pattern matrix folded in earlier is dropped, as there)
    ctx->matInv = ctx->mat;
    vkvg_matrix_invert(&ctx->matInv);
}
void vkvg_translate(VkvgContext ctx, float dx, float dy) { if (!vkvg_status(ctx) && !rec(ctx, RC_TRANSLATE, {dx, dy})) { vkvg_matrix_translate(&ctx->mat, dx, dy); set_mat_inv(ctx); } }
void vkvg_scale(VkvgContext ctx, float sx, float sy) { if (!vkvg_status(ctx) && !rec(ctx, RC_SCALE, {sx, sy})) { vkvg_matrix_scale(&ctx->mat, sx, sy); set_mat_inv(ctx); } }
void vkvg_rotate(VkvgContext ctx, float radians) { if (!vkvg_status(ctx) && !rec(ctx, RC_ROTATE, {radians})) { vkvg_matrix_rotate(&ctx->mat, radians); set_mat_inv(ctx); } }
void vkvg_transform(VkvgContext ctx, const vkvg_matrix_t *matrix) {
    if (vkvg_status(ctx)) return;
    if (rec(ctx, RC_TRANSFORM, {matrix->xx, matrix->yx, matrix->xy, matrix->yy, matrix->x0, matrix->y0})) return;
    vkvg_matrix_t res;
    vkvg_matrix_multiply(&res, &ctx->mat, matrix);
    ctx->mat = res;
    set_mat_inv(ctx);
}
void vkvg_identity_matrix(VkvgContext ctx) { if (!vkvg_status(ctx) && !rec(ctx, RC_IDENTITY_MATRIX)) { vkvg_matrix_init_identity(&ctx->mat); set_mat_inv(ctx); } }
void vkvg_set_matrix(VkvgContext ctx, const vkvg_matrix_t *matrix) { if (!vkvg_status(ctx) && !rec(ctx, RC_SET_MATRIX, {matrix->xx, matrix->yx, matrix->xy, matrix->yy, matrix->x0, matrix->y0})) { ctx->mat = *matrix; set_mat_inv(ctx); } }
void vkvg_get_matrix(VkvgContext ctx, vkvg_matrix_t *const matrix) { if (!vkvg_status(ctx) && matrix) *matrix = ctx->mat; }

// ---- draws ----
static vkb_draw base_draw(VkvgContext ctx, uint32_t kind, uint32_t rule) {
    vkb_draw d;
    d.kind = kind;
    // _bind_draw_pipeline, internal.c:606-621: CLEAR and DIFFERENCE have pipelines of their own, everything else draws with OVER
    const uint32_t bop = ctx->op == VKVG_OPERATOR_CLEAR ? VKB_OP_CLEAR : (ctx->op == VKVG_OPERATOR_DIFFERENCE ? VKB_OP_SUB : VKB_OP_OVER);
    d.rule_pattern = rule | (ctx->patType << 8) | (bop << 16);
    d.first_subpath = ctx->path_first_sp;
    d.n_subpaths    = (uint32_t)ctx->batch.subpaths.size() - ctx->path_first_sp;
    d.color = ctx->curColor; d.opacity = ctx->opacity; d.gradient = 0;
    std::vector<vkb_xform> &xf = ctx->batch.xforms;
    if (xf.empty() || memcmp(xf.back().mat, &ctx->mat, sizeof(float) * 6) != 0 || xf.back().band != ctx->canvas) {
        vkb_xform x;
        memcpy(x.mat, &ctx->mat, sizeof(float) * 6);
        x.band = ctx->canvas; x.pad = 0;
        xf.push_back(x);
    }
    d.xform_stroke = (uint32_t)xf.size() - 1;
    if (ctx->patType == VKB_PAT_SURFACE) {
        vkb_surfpat sp;
        memset(&sp, 0, sizeof sp);
        VkvgSurface src = ctx->pattern->surf;
        sp.image  = (uint64_t)(uintptr_t)vkb_surface_device_pixels(src->impl);
        sp.width  = src->width; sp.height = src->height;
        const uint32_t filter = (ctx->pattern->filter == VKVG_FILTER_BILINEAR || ctx->pattern->filter == VKVG_FILTER_BEST) ? VKB_TEX_LINEAR : VKB_TEX_NEAREST;
        sp.filter_extend = filter | ((uint32_t)ctx->pattern->extend << 8);  // vkvg_extend_t order == VKB_TEX_* address modes
        sp.sx = ctx->src[0]; sp.sy = ctx->src[1];
        memcpy(sp.minv, &ctx->matInv, sizeof(float) * 6);
        std::vector<vkb_surfpat> &tab = ctx->batch.surfpats;
        if (tab.empty() || memcmp(&tab.back(), &sp, sizeof sp) != 0) {
            tab.push_back(sp);
            ctx->held.push_back(vkvg_surface_reference(src));  // stays alive until the draws that sample it have run
        }
        d.gradient = (uint32_t)tab.size() - 1;
    } else if (ctx->patType != VKB_PAT_SOLID) {
        if (ctx->grad_slot < 0) {
            ctx->grad_slot = (int32_t)ctx->batch.grads.size();
            ctx->batch.grads.push_back(ctx->grad);
        }
        d.gradient = (uint32_t)ctx->grad_slot;
    }
    return d;
}
// side tables are addressed with 16 bits: start a new batch before they overflow
static void flush_impl(VkvgContext ctx, vkb_capture *cap, bool keep_resident);
static void reserve_draw_tables(VkvgContext ctx) {
    if (ctx->batch.xforms.size() >= 65000 || ctx->batch.strokes.size() >= 65000) flush_impl(ctx, nullptr, false);
}
static void fill_preserve_(VkvgContext ctx) {  // _fill_preserve :796-821
    finish_path(ctx);
    if (ctx->batch.subpaths.size() == ctx->path_first_sp) return;
    reserve_draw_tables(ctx);
    ctx->batch.draws.push_back(base_draw(ctx, VKB_DRAW_FILL, ctx->fillRule == VKVG_FILL_RULE_EVEN_ODD ? VKB_RULE_EVEN_ODD : VKB_RULE_NON_ZERO));
}
static void stroke_preserve_(VkvgContext ctx) {  // _stroke_preserve :822-948
    finish_path(ctx);
    if (ctx->batch.subpaths.size() == ctx->path_first_sp) return;
    reserve_draw_tables(ctx);
    vkb_stroke st;
    memset(&st, 0, sizeof st);
    st.hw = ctx->lineWidth * 0.5f;
    st.lhMax = ctx->miterLimit * ctx->lineWidth;
    st.arcStep = get_arc_step(ctx, st.hw);
    st.join = ctx->join; st.cap = ctx->cap;
    if (!ctx->dashes.empty()) {
        float tot = 0;
        for (float v : ctx->dashes) tot += v;
        if (tot == 0) { ctx->status = VKVG_STATUS_INVALID_DASH; return; }  // the reference's only dash error (src/vkvg_context.c:853-860)
        if (ctx->dashes.size() > VKB_MAX_DASHES) {  // limit of the stroke kernels: this stroke is skipped, the context stays usable
            static bool warned = false;
            if (!warned) { fprintf(stderr, "vkvg_b200: dash patterns of more than %d entries are not supported; stroke skipped\n", VKB_MAX_DASHES); warned = true; }
            return;
        }
        // reuse the previous stroke's dash table entry when the pattern is unchanged
        std::vector<float> &dt = ctx->batch.dashes;
        bool same = false;
        if (!ctx->batch.strokes.empty()) {
            const vkb_stroke &pv = ctx->batch.strokes.back();
            same = pv.dash_count == ctx->dashes.size() && pv.dash_first + pv.dash_count <= dt.size() &&
                   memcmp(dt.data() + pv.dash_first, ctx->dashes.data(), sizeof(float) * pv.dash_count) == 0;
            if (same) st.dash_first = pv.dash_first;
        }
        if (!same) {
            st.dash_first = (uint32_t)dt.size();
            dt.insert(dt.end(), ctx->dashes.begin(), ctx->dashes.end());
        }
        st.dash_count  = (uint32_t)ctx->dashes.size();
        st.dash_offset = ctx->dashOffset;
    }
    vkb_draw d = base_draw(ctx, VKB_DRAW_STROKE, VKB_RULE_COUNT);
    std::vector<vkb_stroke> &sv = ctx->batch.strokes;
    if (sv.empty() || memcmp(&sv.back(), &st, sizeof st) != 0) sv.push_back(st);
    d.xform_stroke |= ((uint32_t)sv.size() - 1) << 16;
    ctx->batch.draws.push_back(d);
}
void vkvg_fill_preserve(VkvgContext ctx) { if (!vkvg_status(ctx) && !rec(ctx, RC_FILL_PRESERVE)) fill_preserve_(ctx); }
void vkvg_fill(VkvgContext ctx) {
    if (vkvg_status(ctx)) return;
    if (rec(ctx, RC_FILL)) return;
    fill_preserve_(ctx);
    clear_path(ctx);
}
void vkvg_stroke_preserve(VkvgContext ctx) { if (!vkvg_status(ctx) && !rec(ctx, RC_STROKE_PRESERVE)) stroke_preserve_(ctx); }
void vkvg_stroke(VkvgContext ctx) {
    if (vkvg_status(ctx)) return;
    if (rec(ctx, RC_STROKE)) return;
    stroke_preserve_(ctx);
    clear_path(ctx);
}
void vkvg_paint(VkvgContext ctx) {  // :990-1003
    if (vkvg_status(ctx)) return;
    if (rec(ctx, RC_PAINT)) return;
    finish_path(ctx);
    if (ctx->batch.subpaths.size() != ctx->path_first_sp) { vkvg_fill(ctx); return; }
    vkb_draw d = base_draw(ctx, VKB_DRAW_PAINT, VKB_RULE_NON_ZERO);
    d.n_subpaths = 0;
    ctx->batch.draws.push_back(d);
}
void vkvg_clear(VkvgContext ctx) {  // :734-753: everything drawn so far is wiped, so pending draws can be dropped
    if (vkvg_status(ctx)) return;
    if (rec(ctx, RC_CLEAR)) return;
    ctx->curClipState = previous_clip_state(ctx) == CLIP_STATE_CLEAR ? CLIP_STATE_NONE : CLIP_STATE_CLEAR;  // :740-743
    ctx->batch.clear_draws();
    ctx->grad_slot = -1;
    ctx->clear_pending = true;
}

// keep the current (preserved / under construction) path across a flush: compact it to the front of the batch
static void carry_path_over(VkvgContext ctx) {
    vkb_batch &b = ctx->batch;
    uint32_t   e0 = ctx->path_first_sp < b.subpaths.size() ? b.subpaths[ctx->path_first_sp].first_elem : ctx->sp_first_elem;
    uint32_t   d0 = e0 < b.elem_hdr.size() ? (b.elem_hdr[e0] >> VKB_EL_PAYLOAD_SHIFT) : (uint32_t)b.elem_data.size();
    // in place (the vectors keep their capacity from flush to flush: steady-state recording never reallocates)
    size_t nh = b.elem_hdr.size() - e0;
    for (size_t i = 0; i < nh; i++) {
        uint32_t h = b.elem_hdr[e0 + i];
        b.elem_hdr[i] = (h & ((1u << VKB_EL_PAYLOAD_SHIFT) - 1)) | (((h >> VKB_EL_PAYLOAD_SHIFT) - d0) << VKB_EL_PAYLOAD_SHIFT);
    }
    b.elem_hdr.resize(nh);
    b.n_curves = 0;
    for (size_t i = 0; i < nh; i++)
        if ((b.elem_hdr[i] & VKB_EL_TYPE_MASK) != VKB_EL_POINT) b.n_curves++;
    size_t nd = b.elem_data.size() - d0;
    if (nd && d0) memmove(b.elem_data.data(), b.elem_data.data() + d0, nd * sizeof(float));
    b.elem_data.resize(nd);
    size_t ns = b.subpaths.size() - ctx->path_first_sp;
    for (size_t i = 0; i < ns; i++) {
        vkb_subpath sp = b.subpaths[ctx->path_first_sp + i];
        sp.first_elem -= e0;
        b.subpaths[i] = sp;
    }
    b.subpaths.resize(ns);
    b.clear_draws();
    ctx->sp_first_elem -= e0;
    ctx->path_first_sp = 0;
    ctx->grad_slot     = -1;
}
static void flush_impl(VkvgContext ctx, vkb_capture *cap, bool keep_resident) {
    VkvgDevice dev = ctx->dev;
    {
        std::lock_guard<std::mutex> lk(dev->mtx);
        if (ctx->clear_pending) { vkb_surface_clear(ctx->pSurf->impl); ctx->clear_pending = false; }
        if (!ctx->batch.draws.empty() || cap) {
            vkb_stats st;
            bool want_stats = dev->profiling;
            if (vkb_render(dev->impl, ctx->pSurf->impl, dev->raster_samples(), ctx->batch, cap, want_stats ? &st : nullptr)) {
                ctx->status = VKVG_STATUS_DEVICE_ERROR;
                dev->status = VKVG_STATUS_DEVICE_ERROR;
            }
            if (want_stats) dev->last = st;
            {
                vkvg_debug_stats_t &g = dev->dbg;
                auto up = [](uint32_t &m, uint64_t v) { if (v > m) m = (uint32_t)(v > 0xffffffffull ? 0xffffffffull : v); };
                up(g.sizePoints, ctx->batch.elem_hdr.size());   // recorded path elements (a lower bound of the flattened points when curves are in)
                up(g.sizePathes, ctx->batch.subpaths.size());
                if (want_stats) { up(g.sizePoints, st.n_points); up(g.sizeVertices, st.n_verts); up(g.sizeIndices, st.n_inds); up(g.sizeVBO, st.n_verts); up(g.sizeIBO, st.n_inds); }
            }
        }
    }
    (void)keep_resident;
    // sources sampled by the flush before this one are no longer in flight (the upload of this one waited for it)
    for (VkvgSurface s : ctx->held_prev) vkvg_surface_destroy(s);
    ctx->held_prev.swap(ctx->held);
    ctx->held.clear();
    carry_path_over(ctx);
}
void vkvg_flush(VkvgContext ctx) {  // :180-184
    if (vkvg_status(ctx)) return;
    flush_impl(ctx, nullptr, false);
}
void vkvg_destroy(VkvgContext ctx) {  // :246-304
    if (vkvg_status(ctx)) return;
    if (--ctx->references > 0) return;
    if (ctx->recording) { vkvg_recording_destroy(ctx->recording); ctx->recording = NULL; }
    vkvg_flush(ctx);
    if (!ctx->held_prev.empty() || !ctx->held.empty()) {
        vkvg_b200_device_synchronize(ctx->dev);
        for (VkvgSurface s : ctx->held_prev) vkvg_surface_destroy(s);
        for (VkvgSurface s : ctx->held) vkvg_surface_destroy(s);
    }
    if (ctx->pattern) vkvg_pattern_destroy(ctx->pattern);
    for (saved_state &s : ctx->saved) if (s.pattern) vkvg_pattern_destroy(s.pattern);
    vkvg_surface_destroy(ctx->pSurf);
    delete ctx;
}
// ---- recording API: reference include/vkvg.h:1961-1970, src/recording/vkvg_record.c:28-76 ----
void vkvg_start_recording(VkvgContext ctx) {
    if (vkvg_status(ctx)) return;
    if (ctx->recording) vkvg_recording_destroy(ctx->recording);
    ctx->recording = new _vkvg_recording_t();
}
VkvgRecording vkvg_stop_recording(VkvgContext ctx) {  // NULL when nothing was recorded
    if (vkvg_status(ctx)) return NULL;
    VkvgRecording r = ctx->recording;
    ctx->recording  = NULL;
    if (r && r->cmds.empty()) { vkvg_recording_destroy(r); return NULL; }
    return r;
}
uint32_t vkvg_recording_get_count(VkvgRecording rec) { return rec ? (uint32_t)rec->cmds.size() : 0; }
void    *vkvg_recording_get_data(VkvgRecording rec) { return rec ? (void *)rec->buf.data() : NULL; }
void     vkvg_recording_get_command(VkvgRecording rec, uint32_t cmdIndex, uint32_t *cmd, void **dataOffset) {
    if (!rec) return;
    if (cmdIndex < rec->cmds.size()) { *cmd = rec->cmds[cmdIndex].cmd; *dataOffset = (void *)rec->cmds[cmdIndex].off; }
    else { *cmd = 0; *dataOffset = NULL; }
}
void vkvg_recording_destroy(VkvgRecording rec) {
    if (!rec) return;
    for (VkvgPattern p : rec->pats) vkvg_pattern_destroy(p);
    for (VkvgSurface s : rec->surfs) vkvg_surface_destroy(s);
    delete rec;
}
void vkvg_replay_command(VkvgContext ctx, VkvgRecording rec, uint32_t i) {  // _replay_command, src/recording/vkvg_record_internal.c:253-420
    if (vkvg_status(ctx) || !rec || i >= rec->cmds.size()) return;
    const float    *f = (const float *)(rec->buf.data() + rec->cmds[i].off);
    const uint32_t *u = (const uint32_t *)f;
    switch (rec->cmds[i].cmd) {
    case RC_SAVE: vkvg_save(ctx); break;
    case RC_RESTORE: vkvg_restore(ctx); break;
    case RC_NEW_PATH: vkvg_new_path(ctx); break;
    case RC_NEW_SUB_PATH: vkvg_new_sub_path(ctx); break;
    case RC_CLOSE_PATH: vkvg_close_path(ctx); break;
    case RC_MOVE_TO: vkvg_move_to(ctx, f[0], f[1]); break;
    case RC_LINE_TO: vkvg_line_to(ctx, f[0], f[1]); break;
    case RC_REL_MOVE_TO: vkvg_rel_move_to(ctx, f[0], f[1]); break;
    case RC_REL_LINE_TO: vkvg_rel_line_to(ctx, f[0], f[1]); break;
    case RC_RECTANGLE: vkvg_rectangle(ctx, f[0], f[1], f[2], f[3]); break;
    case RC_ARC: vkvg_arc(ctx, f[0], f[1], f[2], f[3], f[4]); break;
    case RC_ARC_NEG: vkvg_arc_negative(ctx, f[0], f[1], f[2], f[3], f[4]); break;
    case RC_CURVE_TO: vkvg_curve_to(ctx, f[0], f[1], f[2], f[3], f[4], f[5]); break;
    case RC_REL_CURVE_TO: vkvg_rel_curve_to(ctx, f[0], f[1], f[2], f[3], f[4], f[5]); break;
    case RC_QUADRATIC_TO: vkvg_quadratic_to(ctx, f[0], f[1], f[2], f[3]); break;
    case RC_REL_QUADRATIC_TO: vkvg_rel_quadratic_to(ctx, f[0], f[1], f[2], f[3]); break;
    case RC_ELLIPTICAL_ARC_TO: vkvg_elliptic_arc_to(ctx, f[0], f[1], u[5] != 0, u[6] != 0, f[2], f[3], f[4]); break;
    case RC_REL_ELLIPTICAL_ARC_TO: vkvg_rel_elliptic_arc_to(ctx, f[0], f[1], u[5] != 0, u[6] != 0, f[2], f[3], f[4]); break;
    case RC_SET_LINE_WIDTH: vkvg_set_line_width(ctx, f[0]); break;
    case RC_SET_MITER_LIMIT: vkvg_set_miter_limit(ctx, f[0]); break;
    case RC_SET_LINE_JOIN: vkvg_set_line_join(ctx, (vkvg_line_join_t)u[0]); break;
    case RC_SET_LINE_CAP: vkvg_set_line_cap(ctx, (vkvg_line_cap_t)u[0]); break;
    case RC_SET_OPERATOR: vkvg_set_operator(ctx, (vkvg_operator_t)u[0]); break;
    case RC_SET_FILL_RULE: vkvg_set_fill_rule(ctx, (vkvg_fill_rule_t)u[0]); break;
    case RC_SET_DASH: vkvg_set_dash(ctx, u[0] ? f + 2 : NULL, u[0], f[1]); break;
    case RC_TRANSLATE: vkvg_translate(ctx, f[0], f[1]); break;
    case RC_ROTATE: vkvg_rotate(ctx, f[0]); break;
    case RC_SCALE: vkvg_scale(ctx, f[0], f[1]); break;
    case RC_TRANSFORM: { vkvg_matrix_t m = {f[0], f[1], f[2], f[3], f[4], f[5]}; vkvg_transform(ctx, &m); break; }
    case RC_SET_MATRIX: { vkvg_matrix_t m = {f[0], f[1], f[2], f[3], f[4], f[5]}; vkvg_set_matrix(ctx, &m); break; }
    case RC_IDENTITY_MATRIX: vkvg_identity_matrix(ctx); break;
    case RC_PAINT: vkvg_paint(ctx); break;
    case RC_FILL: vkvg_fill(ctx); break;
    case RC_STROKE: vkvg_stroke(ctx); break;
    case RC_CLIP: vkvg_clip(ctx); break;
    case RC_RESET_CLIP: vkvg_reset_clip(ctx); break;
    case RC_CLEAR: vkvg_clear(ctx); break;
    case RC_FILL_PRESERVE: vkvg_fill_preserve(ctx); break;
    case RC_STROKE_PRESERVE: vkvg_stroke_preserve(ctx); break;
    case RC_CLIP_PRESERVE: vkvg_clip_preserve(ctx); break;
    case RC_SET_SOURCE_RGB: vkvg_set_source_rgb(ctx, f[0], f[1], f[2]); break;
    case RC_SET_SOURCE_RGBA: vkvg_set_source_rgba(ctx, f[0], f[1], f[2], f[3]); break;
    case RC_SET_SOURCE_COLOR: vkvg_set_source_color(ctx, u[0]); break;
    case RC_SET_SOURCE: if (u[0] < rec->pats.size()) vkvg_set_source(ctx, rec->pats[u[0]]); break;
    case RC_SET_SOURCE_SURFACE: if (u[2] < rec->surfs.size()) vkvg_set_source_surface(ctx, rec->surfs[u[2]], f[0], f[1]); break;
    default: break;
    }
}
void vkvg_replay(VkvgContext ctx, VkvgRecording rec) {
    if (!rec) return;
    for (uint32_t i = 0; i < rec->cmds.size(); i++) vkvg_replay_command(ctx, rec, i);
}
const char *vkvg_status_to_string(vkvg_status_t status) {  // :1647-1692
    switch (status) {
    case VKVG_STATUS_SUCCESS: return "no error has occurred";
    case VKVG_STATUS_INVALID_RESTORE: return "vkvg_restore() without matching vkvg_save()";
    case VKVG_STATUS_NO_CURRENT_POINT: return "no current point defined";
    case VKVG_STATUS_INVALID_MATRIX: return "invalid matrix (not invertible)";
    case VKVG_STATUS_INVALID_STATUS: return "invalid value for an input vkvg_status_t";
    case VKVG_STATUS_INVALID_INDEX: return "invalid index passed to getter";
    case VKVG_STATUS_NULL_POINTER: return "NULL pointer";
    case VKVG_STATUS_WRITE_ERROR: return "error while writing to output stream";
    case VKVG_STATUS_PATTERN_TYPE_MISMATCH: return "the pattern type is not appropriate for the operation";
    case VKVG_STATUS_PATTERN_INVALID_GRADIENT: return "the stops count is zero";
    case VKVG_STATUS_INVALID_FORMAT: return "invalid value for an input vkvg_format_t";
    case VKVG_STATUS_FILE_NOT_FOUND: return "file not found";
    case VKVG_STATUS_INVALID_DASH: return "invalid value for a dash setting";
    case VKVG_STATUS_INVALID_RECT: return "a rectangle has the height or width equal to 0";
    case VKVG_STATUS_TIMEOUT: return "waiting for a Vulkan operation to finish resulted in a fence timeout (5 seconds)";
    case VKVG_STATUS_DEVICE_ERROR: return "the initialization of the device resulted in an error";
    case VKVG_STATUS_INVALID_IMAGE: return "invalid image";
    case VKVG_STATUS_INVALID_SURFACE: return "invalid surface";
    case VKVG_STATUS_INVALID_FONT: return "unresolved font name";
    default: return "<unknown error status>";
    }
}

// ====================================================================================================
// vkvg_b200.h extensions
// ====================================================================================================
extern unsigned long long g_vkb_launches;
uint64_t vkvg_b200_launch_count(void) { return g_vkb_launches; }
void     vkvg_b200_set_profiling(VkvgDevice dev, int on) { if (!vkvg_device_status(dev)) dev->profiling = on != 0; }
void     vkvg_b200_last_stats(VkvgDevice dev, vkvg_b200_stats_t *out) {
    if (vkvg_device_status(dev) || !out) return;
    static_assert(sizeof(vkvg_b200_stats_t) == sizeof(vkb_stats), "stats layout");
    memcpy(out, &dev->last, sizeof *out);
}
void vkvg_b200_device_synchronize(VkvgDevice dev) {
    if (vkvg_device_status(dev)) return;
    std::lock_guard<std::mutex> lk(dev->mtx);
    vkb_device_sync(dev->impl);
}
void vkvg_b200_device_set_stage_timing(VkvgDevice dev, int on) {
    if (vkvg_device_status(dev)) return;
    std::lock_guard<std::mutex> lk(dev->mtx);
    vkb_device_set_stage_timing(dev->impl, on != 0);
}
void vkvg_b200_device_set_graphs(VkvgDevice dev, int on) {
    if (vkvg_device_status(dev)) return;
    std::lock_guard<std::mutex> lk(dev->mtx);
    vkb_device_set_graphs(dev->impl, on != 0);
}
void vkvg_b200_set_fine_kernel(int mode) { vkb_fine_set_mode(mode); }
int  vkvg_b200_get_fine_kernel(void) { return vkb_fine_get_mode(); }
uint64_t vkvg_b200_device_graph_replays(VkvgDevice dev) { return vkvg_device_status(dev) ? 0 : vkb_device_graph_replays(dev->impl); }
void vkvg_b200_get_source_push(VkvgContext ctx, float out[10]) {  // what the reference keeps in pushConsts.source / .matInv
    if (vkvg_status(ctx) || !out) return;
    memcpy(out, ctx->src, sizeof(float) * 4);
    memcpy(out + 4, &ctx->matInv, sizeof(float) * 6);
}
int vkb_device_ordinal(vkb_device_impl *d);
int vkvg_b200_device_ordinal(VkvgDevice dev) { return vkvg_device_status(dev) ? -1 : vkb_device_ordinal(dev->impl); }

// a context's current path as a one-draw batch (the batch is copied: the context is left untouched)
static bool path_batch(VkvgContext ctx, uint32_t kind, vkb_batch &out) {
    finish_path(ctx);
    if (ctx->batch.subpaths.size() == ctx->path_first_sp) return false;
    vkb_batch saved_draws;
    auto swap_tables = [&]() {
        std::swap(saved_draws.draws, ctx->batch.draws);
        std::swap(saved_draws.xforms, ctx->batch.xforms);
        std::swap(saved_draws.strokes, ctx->batch.strokes);
        std::swap(saved_draws.dashes, ctx->batch.dashes);
        std::swap(saved_draws.grads, ctx->batch.grads);
    };
    swap_tables();
    int32_t slot = ctx->grad_slot;
    ctx->grad_slot = -1;
    if (kind == VKB_DRAW_STROKE) stroke_preserve_(ctx); else fill_preserve_(ctx);
    out = ctx->batch;
    swap_tables();
    ctx->grad_slot = slot;
    return !out.draws.empty();
}
static void run_geometry(VkvgContext ctx, const vkb_batch &b, vkb_capture &cap) {
    cap.geometry_only = true;
    std::lock_guard<std::mutex> lk(ctx->dev->mtx);
    if (vkb_render(ctx->dev->impl, ctx->pSurf->impl, ctx->dev->raster_samples(), b, &cap, nullptr)) ctx->status = VKVG_STATUS_DEVICE_ERROR;
}
uint32_t vkvg_b200_flatten_path(VkvgContext ctx, float *xy, uint8_t *curved, uint32_t cap_points, uint32_t *sp_first, uint32_t *sp_count,
                                uint32_t cap_subpaths, uint32_t *n_subpaths) {
    if (n_subpaths) *n_subpaths = 0;
    if (vkvg_status(ctx)) return 0;
    vkb_batch b;
    if (!path_batch(ctx, VKB_DRAW_FILL, b)) return 0;
    std::vector<float> pts; std::vector<uint8_t> fl; std::vector<uint32_t> f, c;
    vkb_capture cap;
    cap.points = &pts; cap.ptflags = &fl; cap.sp_first = &f; cap.sp_count = &c;
    run_geometry(ctx, b, cap);
    // report only the sub-paths of the current path, points re-based to its first point
    uint32_t s0 = b.draws[0].first_subpath, ns = b.draws[0].n_subpaths;
    uint32_t p0 = ns ? f[s0] : 0, p1 = ns ? f[s0 + ns - 1] + c[s0 + ns - 1] : 0;
    // a DROP_LAST sub-path keeps its dropped point in memory: walk sub-path by sub-path to produce a dense array
    uint32_t n = 0;
    for (uint32_t s = 0; s < ns; s++) {
        if (s < cap_subpaths) { if (sp_first) sp_first[s] = n; if (sp_count) sp_count[s] = c[s0 + s]; }
        for (uint32_t k = 0; k < c[s0 + s]; k++, n++)
            if (n < cap_points) {
                if (xy) { xy[2 * n] = pts[2 * (f[s0 + s] + k)]; xy[2 * n + 1] = pts[2 * (f[s0 + s] + k) + 1]; }
                if (curved) curved[n] = fl[f[s0 + s] + k];
            }
    }
    (void)p0; (void)p1;
    if (n_subpaths) *n_subpaths = ns;
    return n;
}
// bounding box of the flattened current path in user space (vkvg_path_extents :684-696, _vkvg_path_extents
// internal.c:1879-1917: the maxima start from FLT_MIN as they do there).  Curves are flattened on the device, so this
// runs the geometry stages once.
void vkvg_path_extents(VkvgContext ctx, float *const x1, float *const y1, float *const x2, float *const y2) {
    if (vkvg_status(ctx)) return;
    vkb_batch b;
    if (!path_batch(ctx, VKB_DRAW_FILL, b)) { *x1 = *x2 = *y1 = *y2 = 0; return; }
    std::vector<float> pts; std::vector<uint32_t> f, c;
    vkb_capture cap;
    cap.points = &pts; cap.sp_first = &f; cap.sp_count = &c;
    run_geometry(ctx, b, cap);
    float xMin = FLT_MAX, yMin = FLT_MAX, xMax = FLT_MIN, yMax = FLT_MIN;
    const uint32_t s0 = b.draws[0].first_subpath, ns = b.draws[0].n_subpaths;
    for (uint32_t s = 0; s < ns; s++)
        for (uint32_t k = 0; k < c[s0 + s]; k++) {
            const float px = pts[2 * (f[s0 + s] + k)], py = pts[2 * (f[s0 + s] + k) + 1];
            if (px < xMin) xMin = px;
            if (px > xMax) xMax = px;
            if (py < yMin) yMin = py;
            if (py > yMax) yMax = py;
        }
    *x1 = xMin; *x2 = xMax; *y1 = yMin; *y2 = yMax;
}
void vkvg_b200_stroke_geometry(VkvgContext ctx, float *xy, uint32_t cap_verts, uint32_t *n_verts, uint32_t *indices, uint32_t cap_indices,
                               uint32_t *n_indices) {
    if (n_verts) *n_verts = 0;
    if (n_indices) *n_indices = 0;
    if (vkvg_status(ctx)) return;
    vkb_batch b;
    if (!path_batch(ctx, VKB_DRAW_STROKE, b)) return;
    // restrict the batch to this single draw so vertex numbering starts at 0
    vkb_draw d = b.draws.back();
    b.draws.assign(1, d);
    std::vector<float> v; std::vector<uint32_t> ix;
    vkb_capture cap;
    cap.verts = &v; cap.inds = &ix;
    run_geometry(ctx, b, cap);
    if (n_verts) *n_verts = (uint32_t)(v.size() / 2);
    if (n_indices) *n_indices = (uint32_t)ix.size();
    if (xy) memcpy(xy, v.data(), sizeof(float) * (v.size() < 2ull * cap_verts ? v.size() : 2ull * cap_verts));
    if (indices) memcpy(indices, ix.data(), 4 * (ix.size() < cap_indices ? ix.size() : (size_t)cap_indices));
}
uint64_t vkvg_b200_path_edges(VkvgContext ctx, int kind, int32_t *edges_xyxy, uint64_t cap_edges) {
    if (vkvg_status(ctx)) return 0;
    vkb_batch b;
    if (!path_batch(ctx, kind ? VKB_DRAW_STROKE : VKB_DRAW_FILL, b)) return 0;
    vkb_draw d = b.draws.back();
    b.draws.assign(1, d);
    std::vector<int32_t> e;
    vkb_capture cap;
    cap.edges = &e;
    run_geometry(ctx, b, cap);
    uint64_t n = e.size() / 4;
    if (edges_xyxy) memcpy(edges_xyxy, e.data(), 16 * (n < cap_edges ? n : cap_edges));
    return n;
}
void vkvg_b200_flush_capture_winding(VkvgContext ctx, int32_t *winding) {
    if (vkvg_status(ctx)) return;
    vkb_capture cap;
    cap.winding = winding;
    cap.winding_draw = ctx->batch.draws.empty() ? 0 : (uint32_t)ctx->batch.draws.size() - 1;
    if (ctx->batch.draws.empty()) {
        memset(winding, 0, (size_t)ctx->pSurf->width * ctx->pSurf->height * (ctx->dev->analytic ? 1 : ctx->dev->samples) * 4);
        flush_impl(ctx, nullptr, false);
        return;
    }
    flush_impl(ctx, &cap, false);
}
int vkb_winding_raw(vkb_device_impl *d, uint32_t samples, const int32_t *edges, uint64_t n, uint32_t w, uint32_t h, int32_t *out);  // pipeline.cu
vkvg_status_t vkvg_b200_winding(VkvgDevice dev, const int32_t *edges_xyxy, uint64_t n_edges, uint32_t width, uint32_t height, int32_t *winding) {
    if (vkvg_device_status(dev)) return VKVG_STATUS_DEVICE_ERROR;
    std::lock_guard<std::mutex> lk(dev->mtx);
    return vkb_winding_raw(dev->impl, dev->raster_samples(), edges_xyxy, n_edges, width, height, winding) ? VKVG_STATUS_DEVICE_ERROR : VKVG_STATUS_SUCCESS;
}
vkvg_status_t vkvg_b200_device_set_coverage_mode(VkvgDevice dev, int mode) {
    if (vkvg_device_status(dev)) return VKVG_STATUS_DEVICE_ERROR;
    if (mode != VKVG_B200_COVERAGE_MSAA && mode != VKVG_B200_COVERAGE_ANALYTIC) return VKVG_STATUS_INVALID_STATUS;
    std::lock_guard<std::mutex> lk(dev->mtx);
    dev->analytic = mode == VKVG_B200_COVERAGE_ANALYTIC;
    return VKVG_STATUS_SUCCESS;
}
int vkvg_b200_device_get_coverage_mode(VkvgDevice dev) {
    return (!vkvg_device_status(dev) && dev->analytic) ? VKVG_B200_COVERAGE_ANALYTIC : VKVG_B200_COVERAGE_MSAA;
}
void vkvg_b200_flush_keep(VkvgContext ctx) { vkvg_flush(ctx); }  // the uploaded batch stays on the device until the next upload
void vkvg_b200_replay_resident(VkvgDevice dev, VkvgSurface surf, int clear_first) {
    if (vkvg_device_status(dev) || vkvg_surface_status(surf)) return;
    std::lock_guard<std::mutex> lk(dev->mtx);
    if (clear_first) vkb_surface_clear(surf->impl);
    vkb_stats st;
    if (vkb_render_resident(dev->impl, surf->impl, dev->raster_samples(), nullptr, dev->profiling ? &st : nullptr)) dev->status = VKVG_STATUS_DEVICE_ERROR;
    if (dev->profiling) dev->last = st;
}
vkvg_status_t vkvg_b200_time_resident(VkvgDevice dev, VkvgSurface surf, uint32_t steps, int clear_first, int flush_l2, vkvg_b200_stats_t *sum) {
    if (vkvg_device_status(dev) || vkvg_surface_status(surf)) return VKVG_STATUS_DEVICE_ERROR;
    std::lock_guard<std::mutex> lk(dev->mtx);
    vkb_stats st;
    if (vkb_time_resident(dev->impl, surf->impl, dev->raster_samples(), steps, clear_first != 0, flush_l2 != 0, &st)) {
        dev->status = VKVG_STATUS_DEVICE_ERROR;
        return VKVG_STATUS_DEVICE_ERROR;
    }
    if (sum) memcpy(sum, &st, sizeof st);
    return VKVG_STATUS_SUCCESS;
}


// ---- the packed command stream with explicit argument counts, decoded on the device when it can be (decode.cu) ----
// the same stream decoded on the host: the calls one by one (the fallback of vkvg_b200_submit, and the definition of what it means)
static vkvg_status_t replay_cmds(VkvgContext ctx, const uint32_t *cmds, uint64_t n_cmds, const float *a, uint64_t n_args) {
    uint64_t k = 0;
    for (uint64_t i = 0; i < n_cmds; i++) {
        const uint32_t op = cmds[i] & 0xFF, na = cmds[i] >> 8;
        if ((uint64_t)na > n_args - k) return VKVG_STATUS_INVALID_INDEX;
        const float *p = a + k;
        auto need = [&](uint32_t n) { return na == n; };
        bool ok = true;
        switch (op) {
        case VKVG_B200_OP_MOVE_TO: if ((ok = need(2))) vkvg_move_to(ctx, p[0], p[1]); break;
        case VKVG_B200_OP_LINE_TO: if ((ok = need(2))) vkvg_line_to(ctx, p[0], p[1]); break;
        case VKVG_B200_OP_CURVE_TO: if ((ok = need(6))) vkvg_curve_to(ctx, p[0], p[1], p[2], p[3], p[4], p[5]); break;
        case VKVG_B200_OP_CLOSE_PATH: vkvg_close_path(ctx); break;
        case VKVG_B200_OP_NEW_PATH: vkvg_new_path(ctx); break;
        case VKVG_B200_OP_ARC: if ((ok = need(5))) vkvg_arc(ctx, p[0], p[1], p[2], p[3], p[4]); break;
        case VKVG_B200_OP_ARC_NEGATIVE: if ((ok = need(5))) vkvg_arc_negative(ctx, p[0], p[1], p[2], p[3], p[4]); break;
        case VKVG_B200_OP_RECTANGLE: if ((ok = need(4))) vkvg_rectangle(ctx, p[0], p[1], p[2], p[3]); break;
        case VKVG_B200_OP_FILL: vkvg_fill(ctx); break;
        case VKVG_B200_OP_FILL_PRESERVE: vkvg_fill_preserve(ctx); break;
        case VKVG_B200_OP_STROKE: vkvg_stroke(ctx); break;
        case VKVG_B200_OP_STROKE_PRESERVE: vkvg_stroke_preserve(ctx); break;
        case VKVG_B200_OP_PAINT: vkvg_paint(ctx); break;
        case VKVG_B200_OP_SET_SOURCE_RGBA: if ((ok = need(4))) vkvg_set_source_rgba(ctx, p[0], p[1], p[2], p[3]); break;
        case VKVG_B200_OP_SET_LINE_WIDTH: if ((ok = need(1))) vkvg_set_line_width(ctx, p[0]); break;
        case VKVG_B200_OP_SET_LINE_CAP: if ((ok = need(1) && p[0] >= 0.0f && p[0] <= 2.0f)) vkvg_set_line_cap(ctx, (vkvg_line_cap_t)(int)p[0]); break;
        case VKVG_B200_OP_SET_LINE_JOIN: if ((ok = need(1) && p[0] >= 0.0f && p[0] <= 2.0f)) vkvg_set_line_join(ctx, (vkvg_line_join_t)(int)p[0]); break;
        case VKVG_B200_OP_SET_MITER_LIMIT: if ((ok = need(1))) vkvg_set_miter_limit(ctx, p[0]); break;
        case VKVG_B200_OP_SET_FILL_RULE: if ((ok = need(1) && p[0] >= 0.0f && p[0] <= 1.0f)) vkvg_set_fill_rule(ctx, (vkvg_fill_rule_t)(int)p[0]); break;
        case VKVG_B200_OP_SET_DASH: if ((ok = na >= 1)) vkvg_set_dash(ctx, p + 1, na - 1, p[0]); break;   // offset d0 .. dn-1
        case VKVG_B200_OP_SET_SOURCE_LINEAR:
        case VKVG_B200_OP_SET_SOURCE_RADIAL: {
            const uint32_t np = op == VKVG_B200_OP_SET_SOURCE_LINEAR ? 4 : 6;
            if (!(ok = na >= np && (na - np) % 5 == 0 && (na - np) / 5 <= 16)) break;
            VkvgPattern pat = np == 4 ? vkvg_pattern_create_linear(p[0], p[1], p[2], p[3]) : vkvg_pattern_create_radial(p[0], p[1], p[2], p[3], p[4], p[5]);
            for (uint32_t j = 0; j < (na - np) / 5; j++) vkvg_pattern_add_color_stop(pat, p[np + 5 * j], p[np + 5 * j + 1], p[np + 5 * j + 2], p[np + 5 * j + 3], p[np + 5 * j + 4]);
            vkvg_set_source(ctx, pat);
            vkvg_pattern_destroy(pat);
            break;
        }
        case VKVG_B200_OP_TRANSLATE: if ((ok = need(2))) vkvg_translate(ctx, p[0], p[1]); break;
        case VKVG_B200_OP_SCALE: if ((ok = need(2))) vkvg_scale(ctx, p[0], p[1]); break;
        case VKVG_B200_OP_ROTATE: if ((ok = need(1))) vkvg_rotate(ctx, p[0]); break;
        case VKVG_B200_OP_IDENTITY_MATRIX: vkvg_identity_matrix(ctx); break;
        case VKVG_B200_OP_SAVE: vkvg_save(ctx); break;
        case VKVG_B200_OP_RESTORE: vkvg_restore(ctx); break;
        case VKVG_B200_OP_CLEAR: vkvg_clear(ctx); break;
        case VKVG_B200_OP_SET_OPACITY: if ((ok = need(1))) vkvg_set_opacity(ctx, p[0]); break;
        case VKVG_B200_OP_POLYLINE: if ((ok = na >= 2 && (na & 1) == 0)) add_polyline(ctx, p, na / 2); break;   // x0 y0 x1 y1 ...
        case VKVG_B200_OP_FLUSH: vkvg_flush(ctx); break;
        case VKVG_B200_OP_SET_CANVAS: if ((ok = need(1) && p[0] >= 0.0f && p[0] <= 16777216.0f)) ok = vkvg_b200_set_canvas(ctx, (uint32_t)p[0]) == VKVG_STATUS_SUCCESS; break;
        case VKVG_B200_OP_CLIP: vkvg_clip(ctx); break;
        case VKVG_B200_OP_CLIP_PRESERVE: vkvg_clip_preserve(ctx); break;
        case VKVG_B200_OP_RESET_CLIP: vkvg_reset_clip(ctx); break;
        default: return VKVG_STATUS_INVALID_STATUS;
        }
        if (!ok) return VKVG_STATUS_INVALID_INDEX;
        if (ctx->status) return ctx->status;
        k += na;
    }
    return ctx->status;
}
static int g_submit_mode = [] {  // VKVG_B200_SUBMIT=host forces the host decoder (A/B timing, tests)
    const char *e = getenv("VKVG_B200_SUBMIT");
    return (e && e[0] == 'h') ? 1 : 0;
}();
void vkvg_b200_set_submit_decoder(int mode) { g_submit_mode = mode == 1 ? 1 : 0; }
static std::atomic<unsigned long long> g_submit_device{0}, g_submit_host{0};
void vkvg_b200_submit_counts(uint64_t *on_device, uint64_t *on_host) {
    if (on_device) *on_device = g_submit_device.load();
    if (on_host) *on_host = g_submit_host.load();
}
vkvg_status_t vkvg_b200_submit(VkvgContext ctx, const uint32_t *cmds, uint64_t n_cmds, const float *args, uint64_t n_args) {
    if (vkvg_status(ctx)) return vkvg_status(ctx);
    if (!cmds || (!args && n_args)) return VKVG_STATUS_NULL_POINTER;
    // the device decodes a stream that starts from a plain state: no path under construction, a solid or gradient source, no clip, no recording
    bool device_ok = g_submit_mode == 0 && !ctx->recording && ctx->sp_points == 0 && ctx->path_first_sp == ctx->batch.subpaths.size() && ctx->patType != VKB_PAT_SURFACE &&
                     ctx->curClipState == CLIP_STATE_NONE && ctx->saved.empty() && ctx->dashes.size() <= VKB_MAX_DASHES && n_cmds > 0;
    if (device_ok) {
        if (!ctx->batch.draws.empty()) flush_impl(ctx, nullptr, false);  // draws recorded call by call come first
        VkvgDevice dev = ctx->dev;
        vkb_decode_init in;
        memset(&in, 0, sizeof in);
        memcpy(in.mat, &ctx->mat, sizeof in.mat);
        in.band = ctx->canvas; in.color = ctx->curColor; in.rule = ctx->fillRule == VKVG_FILL_RULE_EVEN_ODD ? 0u : 1u;
        in.cap = ctx->cap; in.join = ctx->join;
        in.bop = ctx->op == VKVG_OPERATOR_CLEAR ? VKB_OP_CLEAR : (ctx->op == VKVG_OPERATOR_DIFFERENCE ? VKB_OP_SUB : VKB_OP_OVER);
        in.dash_count = (uint32_t)ctx->dashes.size(); in.dash_offset = ctx->dashOffset;
        for (size_t q = 0; q < ctx->dashes.size(); q++) in.dashes[q] = ctx->dashes[q];
        in.lw = ctx->lineWidth; in.miter = ctx->miterLimit; in.opacity = ctx->opacity;
        in.pattern = ctx->patType;
        if (ctx->patType != VKB_PAT_SOLID) in.grad = ctx->grad;
        vkb_decode_census c;
        int r;
        {
            std::lock_guard<std::mutex> lk(dev->mtx);
            if (ctx->clear_pending) { vkb_surface_clear(ctx->pSurf->impl); ctx->clear_pending = false; }
            vkb_stats st;
            r = vkb_submit_stream(dev->impl, ctx->pSurf->impl, dev->raster_samples(), cmds, n_cmds, args, n_args, in, &c, dev->profiling ? &st : nullptr);
            if (r == 0 && dev->profiling) dev->last = st;
        }
        if (r == 1) { ctx->status = VKVG_STATUS_DEVICE_ERROR; dev->status = VKVG_STATUS_DEVICE_ERROR; return ctx->status; }
        if (r == 0) {
            g_submit_device++;
            // leave the context in the state the last setters of the stream put it in (their arguments are in the caller's arrays)
            auto arg = [&](int q) { return args + c.last_setter_arg[q]; };
            if (c.last_setter[0] >= 0) {
                const uint32_t op = cmds[c.last_setter[0]] & 0xFF, na = cmds[c.last_setter[0]] >> 8;
                const float   *p  = arg(0);
                if (op == VKVG_B200_OP_SET_SOURCE_RGBA) vkvg_set_source_rgba(ctx, p[0], p[1], p[2], p[3]);
                else {
                    // (through the CTM the setter saw: final_mat is only right when no CTM command follows it; replay the setter under the final CTM instead
                    //  would differ, so the gradient is rebuilt with the host's own code under the matrix in force at the setter - the final one if none follows)
                    const uint32_t np = op == VKVG_B200_OP_SET_SOURCE_LINEAR ? 4 : 6;
                    VkvgPattern pat = np == 4 ? vkvg_pattern_create_linear(p[0], p[1], p[2], p[3]) : vkvg_pattern_create_radial(p[0], p[1], p[2], p[3], p[4], p[5]);
                    for (uint32_t j = 0; j < (na - np) / 5; j++) vkvg_pattern_add_color_stop(pat, p[np + 5 * j], p[np + 5 * j + 1], p[np + 5 * j + 2], p[np + 5 * j + 3], p[np + 5 * j + 4]);
                    memcpy(&ctx->mat, c.final_mat, sizeof c.final_mat);
                    vkvg_set_source(ctx, pat);
                    vkvg_pattern_destroy(pat);
                }
            }
            if (c.last_setter[1] >= 0) ctx->fillRule = (int)arg(1)[0] == 0 ? VKVG_FILL_RULE_EVEN_ODD : VKVG_FILL_RULE_NON_ZERO;
            if (c.last_setter[2] >= 0) ctx->lineWidth = arg(2)[0];
            if (c.last_setter[3] >= 0) ctx->cap = (vkvg_line_cap_t)(int)arg(3)[0];
            if (c.last_setter[4] >= 0) ctx->join = (vkvg_line_join_t)(int)arg(4)[0];
            if (c.last_setter[5] >= 0) ctx->miterLimit = arg(5)[0];
            if (c.last_setter[6] >= 0) { const uint32_t na = cmds[c.last_setter[6]] >> 8; vkvg_set_dash(ctx, arg(6) + 1, na - 1, arg(6)[0]); }
            if (c.last_setter[7] >= 0) ctx->opacity = arg(7)[0];
            memcpy(&ctx->mat, c.final_mat, sizeof c.final_mat);
            set_mat_inv(ctx);
            ctx->canvas = c.final_band;
            clear_path(ctx);
            return ctx->status;
        }
        // r == 2: not a stream the device decodes
    }
    g_submit_host++;
    const vkvg_status_t st = replay_cmds(ctx, cmds, n_cmds, args, n_args);
    if (st) return st;
    vkvg_flush(ctx);
    return ctx->status;
}

vkvg_status_t vkvg_b200_replay(VkvgContext ctx, const uint8_t *ops, uint64_t n_ops, const float *a, uint64_t n_args) {
    if (vkvg_status(ctx)) return vkvg_status(ctx);
    uint64_t k = 0;
#define NEED(n) if ((uint64_t)(n) > n_args - k) return VKVG_STATUS_INVALID_INDEX  /* (k <= n_args always) */
    for (uint64_t i = 0; i < n_ops; i++) {
        switch (ops[i]) {
        case VKVG_B200_OP_MOVE_TO: NEED(2); vkvg_move_to(ctx, a[k], a[k + 1]); k += 2; break;
        case VKVG_B200_OP_LINE_TO: NEED(2); vkvg_line_to(ctx, a[k], a[k + 1]); k += 2; break;
        case VKVG_B200_OP_CURVE_TO: NEED(6); vkvg_curve_to(ctx, a[k], a[k + 1], a[k + 2], a[k + 3], a[k + 4], a[k + 5]); k += 6; break;
        case VKVG_B200_OP_CLOSE_PATH: vkvg_close_path(ctx); break;
        case VKVG_B200_OP_NEW_PATH: vkvg_new_path(ctx); break;
        case VKVG_B200_OP_ARC: NEED(5); vkvg_arc(ctx, a[k], a[k + 1], a[k + 2], a[k + 3], a[k + 4]); k += 5; break;
        case VKVG_B200_OP_ARC_NEGATIVE: NEED(5); vkvg_arc_negative(ctx, a[k], a[k + 1], a[k + 2], a[k + 3], a[k + 4]); k += 5; break;
        case VKVG_B200_OP_RECTANGLE: NEED(4); vkvg_rectangle(ctx, a[k], a[k + 1], a[k + 2], a[k + 3]); k += 4; break;
        case VKVG_B200_OP_FILL: vkvg_fill(ctx); break;
        case VKVG_B200_OP_FILL_PRESERVE: vkvg_fill_preserve(ctx); break;
        case VKVG_B200_OP_STROKE: vkvg_stroke(ctx); break;
        case VKVG_B200_OP_STROKE_PRESERVE: vkvg_stroke_preserve(ctx); break;
        case VKVG_B200_OP_PAINT: vkvg_paint(ctx); break;
        case VKVG_B200_OP_SET_SOURCE_RGBA: NEED(4); vkvg_set_source_rgba(ctx, a[k], a[k + 1], a[k + 2], a[k + 3]); k += 4; break;
        case VKVG_B200_OP_SET_LINE_WIDTH: NEED(1); vkvg_set_line_width(ctx, a[k]); k += 1; break;
        case VKVG_B200_OP_SET_LINE_CAP: NEED(1); if (!(a[k] >= 0.0f && a[k] <= 2.0f)) return VKVG_STATUS_INVALID_INDEX; vkvg_set_line_cap(ctx, (vkvg_line_cap_t)(int)a[k]); k += 1; break;
        case VKVG_B200_OP_SET_LINE_JOIN: NEED(1); if (!(a[k] >= 0.0f && a[k] <= 2.0f)) return VKVG_STATUS_INVALID_INDEX; vkvg_set_line_join(ctx, (vkvg_line_join_t)(int)a[k]); k += 1; break;
        case VKVG_B200_OP_SET_MITER_LIMIT: NEED(1); vkvg_set_miter_limit(ctx, a[k]); k += 1; break;
        case VKVG_B200_OP_SET_FILL_RULE: NEED(1); if (!(a[k] >= 0.0f && a[k] <= 1.0f)) return VKVG_STATUS_INVALID_INDEX; vkvg_set_fill_rule(ctx, (vkvg_fill_rule_t)(int)a[k]); k += 1; break;
        case VKVG_B200_OP_SET_DASH: {
            NEED(2);
            if (!(a[k] >= 0.0f && a[k] <= 16777216.0f)) return VKVG_STATUS_INVALID_INDEX;  // NaN, negative or absurd count: the conversion below would be undefined
            uint64_t n = (uint64_t)a[k];
            NEED(2 + n);
            vkvg_set_dash(ctx, a + k + 2, (uint32_t)n, a[k + 1]);
            k += 2 + n;
            break;
        }
        case VKVG_B200_OP_SET_SOURCE_LINEAR:
        case VKVG_B200_OP_SET_SOURCE_RADIAL: {
            int np = ops[i] == VKVG_B200_OP_SET_SOURCE_LINEAR ? 4 : 6;
            NEED(np + 1);
            if (!(a[k + np] >= 0.0f && a[k + np] <= 16.0f)) return VKVG_STATUS_INVALID_INDEX;  // a gradient holds at most 16 stops (and 5 * ns must not wrap)
            uint64_t ns = (uint64_t)a[k + np];
            NEED((uint64_t)np + 1 + 5 * ns);
            VkvgPattern pat = np == 4 ? vkvg_pattern_create_linear(a[k], a[k + 1], a[k + 2], a[k + 3])
                                      : vkvg_pattern_create_radial(a[k], a[k + 1], a[k + 2], a[k + 3], a[k + 4], a[k + 5]);
            const float *s = a + k + np + 1;
            for (uint64_t j = 0; j < ns; j++)
                if (vkvg_pattern_add_color_stop(pat, s[5 * j], s[5 * j + 1], s[5 * j + 2], s[5 * j + 3], s[5 * j + 4])) break;
            vkvg_set_source(ctx, pat);
            vkvg_pattern_destroy(pat);
            k += np + 1 + 5 * ns;
            break;
        }
        case VKVG_B200_OP_TRANSLATE: NEED(2); vkvg_translate(ctx, a[k], a[k + 1]); k += 2; break;
        case VKVG_B200_OP_SCALE: NEED(2); vkvg_scale(ctx, a[k], a[k + 1]); k += 2; break;
        case VKVG_B200_OP_ROTATE: NEED(1); vkvg_rotate(ctx, a[k]); k += 1; break;
        case VKVG_B200_OP_IDENTITY_MATRIX: vkvg_identity_matrix(ctx); break;
        case VKVG_B200_OP_SAVE: vkvg_save(ctx); break;
        case VKVG_B200_OP_RESTORE: vkvg_restore(ctx); break;
        case VKVG_B200_OP_CLEAR: vkvg_clear(ctx); break;
        case VKVG_B200_OP_SET_OPACITY: NEED(1); vkvg_set_opacity(ctx, a[k]); k += 1; break;
        case VKVG_B200_OP_POLYLINE: {
            NEED(1);
            uint32_t nbits;
            memcpy(&nbits, &a[k], 4);  // the point count travels as raw uint32 bits (a float cannot hold every count)
            uint64_t n = nbits;
            NEED(1 + 2 * n);
            add_polyline(ctx, a + k + 1, n);
            k += 1 + 2 * n;
            break;
        }
        case VKVG_B200_OP_FLUSH: vkvg_flush(ctx); break;
        case VKVG_B200_OP_SET_CANVAS:
            NEED(1);
            if (!(a[k] >= 0.0f && a[k] <= 16777216.0f) || vkvg_b200_set_canvas(ctx, (uint32_t)a[k])) return VKVG_STATUS_INVALID_INDEX;
            k += 1;
            break;
        case VKVG_B200_OP_CLIP: vkvg_clip(ctx); break;
        case VKVG_B200_OP_CLIP_PRESERVE: vkvg_clip_preserve(ctx); break;
        case VKVG_B200_OP_RESET_CLIP: vkvg_reset_clip(ctx); break;
        default: return VKVG_STATUS_INVALID_STATUS;
        }
        if (ctx->status) return ctx->status;
    }
#undef NEED
    return ctx->status;
}
