/*
 * vkvg.h — drop-in C API for the path-rendering hot path of vkvg, implemented by libvkvg_b200.so
 * (hand-written sm_100a CUDA behind a C host).  Source compatible with the subset of the reference's
 * include/vkvg.h that SURVEY.md §8(b) lists: same names, argument meaning, enum values, struct layouts,
 * refcount and sticky-status conventions.  Each group cites the reference declaration it replaces.
 *
 * Provided: everything SURVEY.md 8(a)/(b) names plus the 8(f) widening - clipping and save / restore, surfaces as paint
 * (vkvg_set_source_surface, vkvg_pattern_create_for_surface, vkvg_surface_create_from_image / _from_bitmap, PNG only), the recording
 * API, the CLEAR and DIFFERENCE operators, vkvg_set_source_color_name, vkvg_device_get_stats.
 * Not provided (out of scope, SURVEY.md 2 / DESIGN.md 7): text and fonts, JPEG decoding, the Vulkan interop entry points
 * (vkvg_surface_get_vk_image ...).  A program that only uses the calls below links unchanged.
 */
#ifndef VKVG_H
#define VKVG_H

#ifdef __cplusplus
extern "C" {
#endif

#include <stdbool.h>
#include <stdint.h>

/* The reference includes <vulkan/vulkan.h> (include/vkvg.h:69) only for these names.  Layout-compatible
 * stand-ins are supplied when the Vulkan headers are absent; handle fields are accepted and ignored. */
#ifndef VULKAN_CORE_H_
typedef uint32_t VkFlags;
typedef VkFlags  VkSampleCountFlags;
typedef struct VkInstance_T       *VkInstance;
typedef struct VkPhysicalDevice_T *VkPhysicalDevice;
typedef struct VkDevice_T         *VkDevice;
typedef struct VkImage_T          *VkImage;
typedef int32_t                    VkFormat;
#define VK_SAMPLE_COUNT_1_BIT 1
#define VK_SAMPLE_COUNT_2_BIT 2
#define VK_SAMPLE_COUNT_4_BIT 4
#define VK_SAMPLE_COUNT_8_BIT 8
#define VK_SAMPLE_COUNT_16_BIT 16
#define VK_FORMAT_B8G8R8A8_UNORM 44
#define VK_FORMAT_R8G8B8A8_UNORM 37
#endif

#if defined(__GNUC__)
#define vkvg_public __attribute__((visibility("default")))
#else
#define vkvg_public
#endif

/* ---- status codes: reference include/vkvg.h:125-150 (same order => same values) ---- */
typedef enum {
    VKVG_STATUS_SUCCESS = 0,
    VKVG_STATUS_NO_MEMORY,
    VKVG_STATUS_NULL_POINTER,
    VKVG_STATUS_INVALID_RESTORE,
    VKVG_STATUS_NO_CURRENT_POINT,
    VKVG_STATUS_INVALID_MATRIX,
    VKVG_STATUS_INVALID_STATUS,
    VKVG_STATUS_INVALID_INDEX,
    VKVG_STATUS_WRITE_ERROR,
    VKVG_STATUS_PATTERN_TYPE_MISMATCH,
    VKVG_STATUS_PATTERN_INVALID_GRADIENT,
    VKVG_STATUS_INVALID_FORMAT,
    VKVG_STATUS_FILE_NOT_FOUND,
    VKVG_STATUS_INVALID_DASH,
    VKVG_STATUS_INVALID_RECT,
    VKVG_STATUS_TIMEOUT,
    VKVG_STATUS_DEVICE_ERROR,
    VKVG_STATUS_INVALID_DEVICE_CREATE_INFO,
    VKVG_STATUS_INVALID_IMAGE,
    VKVG_STATUS_INVALID_SURFACE,
    VKVG_STATUS_INVALID_FONT,
    VKVG_STATUS_IN_CACHE,
    VKVG_STATUS_ENUM_MAX = 0x7FFFFFFF
} vkvg_status_t;

/* ---- enums: reference include/vkvg.h:161-225, :855-891 ---- */
typedef enum { VKVG_EXTEND_NONE, VKVG_EXTEND_REPEAT, VKVG_EXTEND_REFLECT, VKVG_EXTEND_PAD } vkvg_extend_t;
typedef enum {
    VKVG_FILTER_FAST, VKVG_FILTER_GOOD, VKVG_FILTER_BEST, VKVG_FILTER_NEAREST, VKVG_FILTER_BILINEAR, VKVG_FILTER_GAUSSIAN
} vkvg_filter_t;
typedef enum {
    VKVG_PATTERN_TYPE_SOLID, VKVG_PATTERN_TYPE_SURFACE, VKVG_PATTERN_TYPE_LINEAR, VKVG_PATTERN_TYPE_RADIAL,
    VKVG_PATTERN_TYPE_MESH, VKVG_PATTERN_TYPE_RASTER_SOURCE
} vkvg_pattern_type_t;
typedef enum { VKVG_LINE_CAP_BUTT, VKVG_LINE_CAP_ROUND, VKVG_LINE_CAP_SQUARE } vkvg_line_cap_t;
typedef enum { VKVG_LINE_JOIN_MITER, VKVG_LINE_JOIN_ROUND, VKVG_LINE_JOIN_BEVEL } vkvg_line_join_t;
typedef enum { VKVG_FILL_RULE_EVEN_ODD, VKVG_FILL_RULE_NON_ZERO } vkvg_fill_rule_t;
typedef enum _vkvg_operator {
    VKVG_OPERATOR_CLEAR, VKVG_OPERATOR_SOURCE, VKVG_OPERATOR_OVER, VKVG_OPERATOR_DIFFERENCE, VKVG_OPERATOR_MAX
} vkvg_operator_t;

typedef struct { float r, g, b, a; } vkvg_color_t;

/* ---- opaque handles: reference include/vkvg.h:294-329 ---- */
typedef struct _vkvg_context_t *VkvgContext;
typedef struct _vkvg_surface_t *VkvgSurface;
typedef struct _vkvg_device_t  *VkvgDevice;
typedef struct _vkvg_pattern_t *VkvgPattern;

/* ---- matrices: reference include/vkvg.h:358-514 (x' = xx*x + xy*y + x0 ; y' = yx*x + yy*y + y0) ---- */
typedef struct { float xx, yx, xy, yy, x0, y0; } vkvg_matrix_t;
#define VKVG_IDENTITY_MATRIX (vkvg_matrix_t){1, 0, 0, 1, 0, 0}
vkvg_public void vkvg_matrix_init_identity(vkvg_matrix_t *matrix);
vkvg_public void vkvg_matrix_init(vkvg_matrix_t *matrix, float xx, float yx, float xy, float yy, float x0, float y0);
vkvg_public void vkvg_matrix_init_translate(vkvg_matrix_t *matrix, float tx, float ty);
vkvg_public void vkvg_matrix_init_scale(vkvg_matrix_t *matrix, float sx, float sy);
vkvg_public void vkvg_matrix_init_rotate(vkvg_matrix_t *matrix, float radians);
vkvg_public void vkvg_matrix_translate(vkvg_matrix_t *matrix, float tx, float ty);
vkvg_public void vkvg_matrix_scale(vkvg_matrix_t *matrix, float sx, float sy);
vkvg_public void vkvg_matrix_rotate(vkvg_matrix_t *matrix, float radians);
vkvg_public void vkvg_matrix_multiply(vkvg_matrix_t *result, const vkvg_matrix_t *a, const vkvg_matrix_t *b);
vkvg_public void vkvg_matrix_transform_distance(const vkvg_matrix_t *matrix, float *dx, float *dy);
vkvg_public void vkvg_matrix_transform_point(const vkvg_matrix_t *matrix, float *x, float *y);
vkvg_public vkvg_status_t vkvg_matrix_invert(vkvg_matrix_t *matrix);
vkvg_public void          vkvg_matrix_get_scale(const vkvg_matrix_t *matrix, float *sx, float *sy);

/* ---- device: reference include/vkvg.h:555-664.  `samples` selects the MSAA-matching sample count
 *      (1, 2, 4, 8, 16); the Vulkan handle fields are ignored (the device is CUDA device 0 or the one
 *      named by VKVG_B200_DEVICE / LOCAL_RANK). ---- */
typedef struct {
    VkSampleCountFlags samples;
    bool               deferredResolve;
    VkInstance         inst;
    VkPhysicalDevice   phy;
    VkDevice           vkdev;
    uint32_t           qFamIdx;
    uint32_t           qIndex;
    bool               threadAware;
} vkvg_device_create_info_t;
vkvg_public VkvgDevice    vkvg_device_create(vkvg_device_create_info_t *info);
vkvg_public void          vkvg_device_destroy(VkvgDevice dev);
vkvg_public vkvg_status_t vkvg_device_status(VkvgDevice dev);
vkvg_public VkvgDevice    vkvg_device_reference(VkvgDevice dev);
vkvg_public uint32_t      vkvg_device_get_reference_count(VkvgDevice dev);
vkvg_public void          vkvg_device_set_dpy(VkvgDevice dev, int hdpy, int vdpy);
vkvg_public void          vkvg_device_get_dpy(VkvgDevice dev, int *hdpy, int *vdpy);
vkvg_public void          vkvg_device_set_context_cache_size(VkvgDevice dev, uint32_t maxCount);
/* reference include/vkvg.h:331-349 (its VKVG_DBG_STATS build): high-water marks of the path / vertex arrays */
#define VKVG_HAS_DBG_STATS
typedef struct {
    uint32_t sizePoints, sizePathes, sizeVertices, sizeIndices, sizeVBO, sizeIBO;
} vkvg_debug_stats_t;
vkvg_public vkvg_debug_stats_t vkvg_device_get_stats(VkvgDevice dev);
vkvg_public void               vkvg_device_reset_stats(VkvgDevice dev);

/* ---- surface: reference include/vkvg.h:725-845 ---- */
vkvg_public VkvgSurface   vkvg_surface_create(VkvgDevice dev, uint32_t width, uint32_t height);
vkvg_public vkvg_status_t vkvg_surface_status(VkvgSurface surf);
vkvg_public VkvgSurface   vkvg_surface_reference(VkvgSurface surf);
vkvg_public uint32_t      vkvg_surface_get_reference_count(VkvgSurface surf);
vkvg_public void          vkvg_surface_destroy(VkvgSurface surf);
vkvg_public void          vkvg_surface_clear(VkvgSurface surf);
vkvg_public VkImage       vkvg_surface_get_vk_image(VkvgSurface surf);  /* always NULL: there is no VkImage */
vkvg_public VkFormat      vkvg_surface_get_vk_format(VkvgSurface surf); /* VK_FORMAT_B8G8R8A8_UNORM, as the reference */
vkvg_public uint32_t      vkvg_surface_get_width(VkvgSurface surf);
vkvg_public uint32_t      vkvg_surface_get_height(VkvgSurface surf);
vkvg_public vkvg_status_t vkvg_surface_write_to_png(VkvgSurface surf, const char *path);
vkvg_public vkvg_status_t vkvg_surface_write_to_memory(VkvgSurface surf, unsigned char *const bitmap);
/* reference include/vkvg.h:735, :752: a surface initialised from an image file (PNG here; the reference uses stb_image) or
 * from width*height RGBA8 pixels, taken as they are (premultiplied when used as a source) */
vkvg_public VkvgSurface vkvg_surface_create_from_image(VkvgDevice dev, const char *filePath);
vkvg_public VkvgSurface vkvg_surface_create_from_bitmap(VkvgDevice dev, unsigned char *img, uint32_t width, uint32_t height);
vkvg_public void          vkvg_surface_resolve(VkvgSurface surf);

/* ---- context life cycle: reference include/vkvg.h:906-952 ---- */
vkvg_public VkvgContext   vkvg_create(VkvgSurface surf);
vkvg_public void          vkvg_destroy(VkvgContext ctx);
vkvg_public vkvg_status_t vkvg_status(VkvgContext ctx);
vkvg_public const char   *vkvg_status_to_string(vkvg_status_t status);
vkvg_public VkvgContext   vkvg_reference(VkvgContext ctx);
vkvg_public uint32_t      vkvg_get_reference_count(VkvgContext ctx);
vkvg_public void          vkvg_flush(VkvgContext ctx);

/* ---- path construction: reference include/vkvg.h:961-1198 ---- */
vkvg_public void vkvg_new_path(VkvgContext ctx);
vkvg_public void vkvg_close_path(VkvgContext ctx);
vkvg_public void vkvg_new_sub_path(VkvgContext ctx);
vkvg_public void vkvg_get_current_point(VkvgContext ctx, float *x, float *y);
vkvg_public bool vkvg_has_current_point(VkvgContext ctx);
vkvg_public void vkvg_line_to(VkvgContext ctx, float x, float y);
vkvg_public void vkvg_rel_line_to(VkvgContext ctx, float dx, float dy);
vkvg_public void vkvg_move_to(VkvgContext ctx, float x, float y);
vkvg_public void vkvg_rel_move_to(VkvgContext ctx, float x, float y);
vkvg_public void vkvg_arc(VkvgContext ctx, float xc, float yc, float radius, float a1, float a2);
vkvg_public void vkvg_arc_negative(VkvgContext ctx, float xc, float yc, float radius, float a1, float a2);
vkvg_public void vkvg_curve_to(VkvgContext ctx, float x1, float y1, float x2, float y2, float x3, float y3);
vkvg_public void vkvg_rel_curve_to(VkvgContext ctx, float x1, float y1, float x2, float y2, float x3, float y3);
vkvg_public void vkvg_quadratic_to(VkvgContext ctx, float x1, float y1, float x2, float y2);
vkvg_public void vkvg_rel_quadratic_to(VkvgContext ctx, float x1, float y1, float x2, float y2);
vkvg_public vkvg_status_t vkvg_rectangle(VkvgContext ctx, float x, float y, float w, float h);
vkvg_public vkvg_status_t vkvg_rounded_rectangle(VkvgContext ctx, float x, float y, float w, float h, float radius);
vkvg_public void vkvg_ellipse(VkvgContext ctx, float radiusX, float radiusY, float x, float y, float rotationAngle);
/* reference include/vkvg.h:1184: rectangle whose corners are quarter ellipses of radii rx, ry */
vkvg_public void vkvg_rounded_rectangle2(VkvgContext ctx, float x, float y, float w, float h, float rx, float ry);
/* reference include/vkvg.h:1219, :1235: SVG-style elliptical arc from the current point (phi in radians) */
vkvg_public void vkvg_elliptic_arc_to(VkvgContext ctx, float x, float y, bool large_arc_flag, bool sweep_flag, float rx, float ry, float phi);
vkvg_public void vkvg_rel_elliptic_arc_to(VkvgContext ctx, float x, float y, bool large_arc_flag, bool sweep_flag, float rx, float ry, float phi);
/* reference include/vkvg.h:989: user-space bounding box of the flattened current path (0,0,0,0 without a path) */
vkvg_public void vkvg_path_extents(VkvgContext ctx, float *const x1, float *const y1, float *const x2, float *const y2);

/* ---- drawing: reference include/vkvg.h:1246-1291 ---- */
vkvg_public void vkvg_stroke(VkvgContext ctx);
vkvg_public void vkvg_stroke_preserve(VkvgContext ctx);
vkvg_public void vkvg_fill(VkvgContext ctx);
vkvg_public void vkvg_fill_preserve(VkvgContext ctx);
vkvg_public void vkvg_paint(VkvgContext ctx);
vkvg_public void vkvg_clear(VkvgContext ctx);
/* a surface as paint: reference include/vkvg.h:1386 (source drawn with its origin at x, y in device pixels) */
vkvg_public void vkvg_set_source_surface(VkvgContext ctx, VkvgSurface surf, float x, float y);
/* clipping: reference include/vkvg.h:1299-1325.  The clip region is the intersection of every path clipped so far (with
 * the fill rule current at each call); it is part of the state vkvg_save / vkvg_restore stack. */
vkvg_public void vkvg_reset_clip(VkvgContext ctx);
vkvg_public void vkvg_clip(VkvgContext ctx);
vkvg_public void vkvg_clip_preserve(VkvgContext ctx);

/* ---- sources and stroke/fill state: reference include/vkvg.h:1334-1555 ---- */
vkvg_public void  vkvg_set_opacity(VkvgContext ctx, float opacity);
vkvg_public float vkvg_get_opacity(VkvgContext ctx);
vkvg_public void  vkvg_set_source_color(VkvgContext ctx, uint32_t c);
vkvg_public void  vkvg_set_source_color_name(VkvgContext ctx, const char *color); /* reference include/vkvg.h:1958 (declared there, never defined) */
vkvg_public void  vkvg_set_source_rgba(VkvgContext ctx, float r, float g, float b, float a);
vkvg_public void  vkvg_set_source_rgb(VkvgContext ctx, float r, float g, float b);
vkvg_public void  vkvg_set_source(VkvgContext ctx, VkvgPattern pat);
vkvg_public void  vkvg_set_line_width(VkvgContext ctx, float width);
vkvg_public void  vkvg_set_miter_limit(VkvgContext ctx, float limit);
vkvg_public float vkvg_get_miter_limit(VkvgContext ctx);
vkvg_public void  vkvg_set_line_cap(VkvgContext ctx, vkvg_line_cap_t cap);
vkvg_public void  vkvg_set_line_join(VkvgContext ctx, vkvg_line_join_t join);
vkvg_public void  vkvg_set_operator(VkvgContext ctx, vkvg_operator_t op);
vkvg_public void  vkvg_set_fill_rule(VkvgContext ctx, vkvg_fill_rule_t fr);
vkvg_public void  vkvg_set_dash(VkvgContext ctx, const float *dashes, uint32_t num_dashes, float offset);
vkvg_public void  vkvg_get_dash(VkvgContext ctx, const float *dashes, uint32_t *num_dashes, float *offset);
vkvg_public float            vkvg_get_line_width(VkvgContext ctx);
vkvg_public vkvg_line_cap_t  vkvg_get_line_cap(VkvgContext ctx);
vkvg_public vkvg_line_join_t vkvg_get_line_join(VkvgContext ctx);
vkvg_public vkvg_operator_t  vkvg_get_operator(VkvgContext ctx);
vkvg_public vkvg_fill_rule_t vkvg_get_fill_rule(VkvgContext ctx);
vkvg_public VkvgPattern      vkvg_get_source(VkvgContext ctx);
vkvg_public VkvgSurface      vkvg_get_target(VkvgContext ctx);

/* ---- save/restore and the CTM: reference include/vkvg.h:1565-1637 ---- */
vkvg_public void vkvg_save(VkvgContext ctx);
vkvg_public void vkvg_restore(VkvgContext ctx);
vkvg_public void vkvg_translate(VkvgContext ctx, float dx, float dy);
vkvg_public void vkvg_scale(VkvgContext ctx, float sx, float sy);
vkvg_public void vkvg_rotate(VkvgContext ctx, float radians);
vkvg_public void vkvg_transform(VkvgContext ctx, const vkvg_matrix_t *matrix);
vkvg_public void vkvg_set_matrix(VkvgContext ctx, const vkvg_matrix_t *matrix);
vkvg_public void vkvg_get_matrix(VkvgContext ctx, vkvg_matrix_t *const matrix);
vkvg_public void vkvg_identity_matrix(VkvgContext ctx);

/* ---- gradient patterns: reference include/vkvg.h:1763-1953 ---- */
vkvg_public vkvg_status_t vkvg_pattern_status(VkvgPattern pat);
vkvg_public VkvgPattern   vkvg_pattern_reference(VkvgPattern pat);
vkvg_public uint32_t      vkvg_pattern_get_reference_count(VkvgPattern pat);
vkvg_public VkvgPattern   vkvg_pattern_create_linear(float x0, float y0, float x1, float y1);
vkvg_public vkvg_status_t vkvg_pattern_edit_linear(VkvgPattern pat, float x0, float y0, float x1, float y1);
vkvg_public vkvg_status_t vkvg_pattern_get_linear_points(VkvgPattern pat, float *x0, float *y0, float *x1, float *y1);
vkvg_public VkvgPattern   vkvg_pattern_create_radial(float cx0, float cy0, float radius0, float cx1, float cy1, float radius1);
/* reference include/vkvg.h:1790: the surface as a paint; extend NONE / REPEAT / REFLECT / PAD and nearest or bilinear filtering
 * through vkvg_pattern_set_extend / set_filter, placed by the pattern matrix and the CTM */
vkvg_public VkvgPattern   vkvg_pattern_create_for_surface(VkvgSurface surf);
vkvg_public vkvg_status_t vkvg_pattern_edit_radial(VkvgPattern pat, float cx0, float cy0, float radius0, float cx1, float cy1,
                                                   float radius1);
vkvg_public vkvg_status_t vkvg_pattern_get_color_stop_count(VkvgPattern pat, uint32_t *count);
vkvg_public vkvg_status_t vkvg_pattern_get_color_stop_rgba(VkvgPattern pat, uint32_t index, float *offset, float *r, float *g,
                                                           float *b, float *a);
vkvg_public void          vkvg_pattern_destroy(VkvgPattern pat);
vkvg_public vkvg_status_t vkvg_pattern_add_color_stop(VkvgPattern pat, float offset, float r, float g, float b, float a);
vkvg_public void          vkvg_pattern_set_extend(VkvgPattern pat, vkvg_extend_t extend);
vkvg_public void          vkvg_pattern_set_filter(VkvgPattern pat, vkvg_filter_t filter);
vkvg_public vkvg_extend_t vkvg_pattern_get_extend(VkvgPattern pat);
vkvg_public vkvg_filter_t vkvg_pattern_get_filter(VkvgPattern pat);
vkvg_public vkvg_pattern_type_t vkvg_pattern_get_type(VkvgPattern pat);
vkvg_public void          vkvg_pattern_set_matrix(VkvgPattern pat, const vkvg_matrix_t *matrix);
vkvg_public void          vkvg_pattern_get_matrix(VkvgPattern pat, vkvg_matrix_t *matrix);

/* ---- recording: reference include/vkvg.h:1961-1970 (there only in builds with VKVG_RECORDING) ----
 * Between vkvg_start_recording and vkvg_stop_recording the context stores drawing calls instead of executing them;
 * vkvg_replay issues them on any context.  Command codes reported by vkvg_recording_get_command are the reference's
 * (src/recording/vkvg_record_internal.h:28-95). */
typedef struct _vkvg_recording_t *VkvgRecording;
vkvg_public void          vkvg_start_recording(VkvgContext ctx);
vkvg_public VkvgRecording vkvg_stop_recording(VkvgContext ctx);
vkvg_public void          vkvg_replay(VkvgContext ctx, VkvgRecording rec);
vkvg_public void          vkvg_replay_command(VkvgContext ctx, VkvgRecording rec, uint32_t cmdIndex);
vkvg_public void          vkvg_recording_get_command(VkvgRecording rec, uint32_t cmdIndex, uint32_t *cmd, void **dataOffset);
vkvg_public uint32_t      vkvg_recording_get_count(VkvgRecording rec);
vkvg_public void         *vkvg_recording_get_data(VkvgRecording rec);
vkvg_public void          vkvg_recording_destroy(VkvgRecording rec);

#ifdef __cplusplus
}
#endif
#endif /* VKVG_H */
