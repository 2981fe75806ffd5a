/* TEST INFRASTRUCTURE ONLY — CPU restatement of vkvg's path-rendering hot path.
 *
 * This is the parity oracle for the CUDA implementation in vkvg_b200/csrc.  It restates, in plain
 * sequential C, the algorithms of the reference (file:line cited on each function in vkvg_oracle.c):
 *   path building + flattening   src/vkvg_context.c:350-683, src/vkvg_context_internal.c:136-252,1313-1472
 *   stroke expansion             src/vkvg_context.c:822-948, src/vkvg_context_internal.c:924-1282
 *   fill (even-odd fan / non-zero) src/vkvg_context_internal.c:1583-1655,1720-1793
 *   paint evaluation             shaders/vkvg_main.frag:68-157, src/vkvg_context_internal.c:774-826
 *   stencil / blend semantics    src/vkvg_device_internal.c:200-246
 * and a scalar restatement of the Vulkan rasterisation the reference delegates to its ICD (top-left
 * rule, 8 sub-pixel bits, standard sample positions, per-sample stencil + blend, box resolve).
 *
 * PINNING: the reference holds no golden vectors for this path.  The oracle is pinned
 * against the reference's OWN object code instead: tests/test_oracle_vs_ref.py drives identical call
 * sequences through oracle/_ref/libvkvg_ref.so (the unmodified reference sources) and compares
 * flattened points, stroke vertices/indices and the rasterised draw lists.  The rasterisation rules
 * themselves live in Mesa, which is not in the reference tree: that part is "parity unpinned" and is
 * defined by this file.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load
 * this library.  The product path never does.
 */
#ifndef VKVG_ORACLE_H
#define VKVG_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct ovk_ctx ovk_ctx;

/* enums use the reference's numeric values (include/vkvg.h:197-225) */
enum { OVK_CAP_BUTT = 0, OVK_CAP_ROUND = 1, OVK_CAP_SQUARE = 2 };
enum { OVK_JOIN_MITER = 0, OVK_JOIN_ROUND = 1, OVK_JOIN_BEVEL = 2 };
enum { OVK_FILL_EVEN_ODD = 0, OVK_FILL_NON_ZERO = 1 };
enum { OVK_RULE_EVEN_ODD = 0, OVK_RULE_NON_ZERO = 1, OVK_RULE_COUNT = 2 };

/* ---- context: surface + state ---- */
ovk_ctx *ovk_create(uint32_t width, uint32_t height, uint32_t samples);
/* logical W x H surface of which only the window [x0, x0+w) x [y0, y0+h) is stored / rasterised (MSAA mode only) */
ovk_ctx *ovk_create_window(uint32_t W, uint32_t H, uint32_t S, uint32_t x0, uint32_t y0, uint32_t w, uint32_t h);
void     ovk_destroy(ovk_ctx *c);
void     ovk_clear(ovk_ctx *c);
int      ovk_status(ovk_ctx *c);

/* ---- path construction (same semantics as the vkvg_* calls of the same name) ---- */
void ovk_new_path(ovk_ctx *c);
void ovk_new_sub_path(ovk_ctx *c);
void ovk_close_path(ovk_ctx *c);
void ovk_move_to(ovk_ctx *c, float x, float y);
void ovk_line_to(ovk_ctx *c, float x, float y);
void ovk_rel_move_to(ovk_ctx *c, float x, float y);
void ovk_rel_line_to(ovk_ctx *c, float x, float y);
void ovk_curve_to(ovk_ctx *c, float x1, float y1, float x2, float y2, float x3, float y3);
void ovk_rel_curve_to(ovk_ctx *c, float x1, float y1, float x2, float y2, float x3, float y3);
void ovk_quadratic_to(ovk_ctx *c, float x1, float y1, float x2, float y2);
void ovk_arc(ovk_ctx *c, float xc, float yc, float radius, float a1, float a2);
void ovk_arc_negative(ovk_ctx *c, float xc, float yc, float radius, float a1, float a2);
int  ovk_rectangle(ovk_ctx *c, float x, float y, float w, float h);
void ovk_get_current_point(ovk_ctx *c, float *x, float *y);

/* ---- state ---- */
void ovk_set_line_width(ovk_ctx *c, float w);
void ovk_set_miter_limit(ovk_ctx *c, float l);
void ovk_set_line_cap(ovk_ctx *c, int cap);
void ovk_set_line_join(ovk_ctx *c, int join);
void ovk_set_dash(ovk_ctx *c, const float *dashes, uint32_t n, float offset);
void ovk_set_fill_rule(ovk_ctx *c, int rule);
void ovk_set_opacity(ovk_ctx *c, float o);
void ovk_set_operator(ovk_ctx *c, int vkvg_operator); /* CLEAR and DIFFERENCE have pipelines of their own, the rest is OVER */
void ovk_set_source_rgba(ovk_ctx *c, float r, float g, float b, float a);
void ovk_set_source_color(ovk_ctx *c, uint32_t rgba);
/* gradient sources: stops = n x {offset, r, g, b, a}; control points in user space at the time of the call */
void ovk_set_source_linear(ovk_ctx *c, float x0, float y0, float x1, float y1, const float *stops, uint32_t n);
void ovk_set_source_radial(ovk_ctx *c, float cx0, float cy0, float r0, float cx1, float cy1, float r1, const float *stops,
                           uint32_t n);
void ovk_set_source_surface(ovk_ctx *c, const uint32_t *rgba_premultiplied, uint32_t w, uint32_t h, float x, float y, int extend, int linear,
                            const float *pattern_matrix, int keep_offset);
void ovk_get_source_push(ovk_ctx *c, float out[10]);
void ovk_translate(ovk_ctx *c, float dx, float dy);
void ovk_scale(ovk_ctx *c, float sx, float sy);
void ovk_rotate(ovk_ctx *c, float radians);
void ovk_set_matrix(ovk_ctx *c, const float m[6]); /* xx yx xy yy x0 y0 */
void ovk_get_matrix(ovk_ctx *c, float m[6]);
void ovk_identity_matrix(ovk_ctx *c);

/* ---- drawing ---- */
void ovk_fill(ovk_ctx *c);
void ovk_fill_preserve(ovk_ctx *c);
void ovk_stroke(ovk_ctx *c);
void ovk_stroke_preserve(ovk_ctx *c);
void ovk_paint(ovk_ctx *c);
/* clipping (src/vkvg_context.c:698-795) and the graphics-state stack (:1251-1512) with the reference's clip-state bookkeeping */
void ovk_clip(ovk_ctx *c);
void ovk_clip_preserve(ovk_ctx *c);
void ovk_reset_clip(ovk_ctx *c);
void ovk_save(ovk_ctx *c);
void ovk_restore(ovk_ctx *c);
void ovk_clear_ctx(ovk_ctx *c); /* vkvg_clear as a context call (ovk_clear only wipes the surface) */

/* ---- introspection used by the parity tests ---- */
/* current path as the reference stores it: points (user space) and the `pathes` table; finishes the open sub-path */
uint32_t ovk_path_points(ovk_ctx *c, const float **pts);
uint32_t ovk_path_table(ovk_ctx *c, const uint32_t **pathes);
/* geometry produced by the LAST stroke / non-zero-convex call: reference vertex order, indices relative to the call */
uint32_t ovk_last_vertices(ovk_ctx *c, const float **xy);
uint32_t ovk_last_indices(ovk_ctx *c, const uint32_t **idx);
/* resolved RGBA8 (premultiplied) pixels, row-major, R in byte 0 */
const uint8_t *ovk_pixels(ovk_ctx *c);
/* per-sample colours (height*width*samples RGBA8) */
const uint8_t *ovk_sample_pixels(ovk_ctx *c);
/* un-premultiplied copy as vkvg_surface_write_to_memory produces (src/vkvg_surface.c:371-382) */
void ovk_write_to_memory(ovk_ctx *c, uint8_t *out);
/* when set, every draw also records per-sample coverage counts of that draw only (for winding parity) */
void           ovk_set_capture_coverage(ovk_ctx *c, int on);
const int32_t *ovk_last_coverage(ovk_ctx *c); /* height*width*samples, value per rule: EO parity, NZ winding, COUNT triangles */

/* ---- stand-alone pieces ---- */
/* adaptive cubic flattening exactly as _curve_to + _recursive_bezier; returns number of points written
 * (excluding p0, including the end point); out may be NULL to count only */
uint32_t ovk_flatten_cubic(float x0, float y0, float x1, float y1, float x2, float y2, float x3, float y3, float tolerance,
                           float *out_xy, uint32_t cap);
/* exact per-sample integer winding of a set of directed fixed-point (24.8) edges, brute force over all edges:
 * out[(y*width+x)*samples+s] = sum over edges of sign * [edge crosses the -x ray from the sample].
 * This is the definition the tile-binned CUDA rasteriser must reproduce bit-for-bit. */
void ovk_winding_brute(const int32_t *edges_xyxy, uint64_t n_edges, uint32_t width, uint32_t height, uint32_t samples,
                       int32_t *out);
/* analytic-coverage mode: A = integral of the winding number of the directed 24.8 edges over each pixel square,
 * out[y*width+x], evaluated edge by edge in double precision over the whole surface (no tiles, no backdrop) */
void ovk_area_brute(const int32_t *edges_xyxy, uint64_t n_edges, uint32_t width, uint32_t height, double *out);
/* analytic != 0: every later draw uses exact area coverage (one colour per pixel; create the context with 1 sample) */
void          ovk_set_coverage_mode(ovk_ctx *c, int analytic);
const double *ovk_last_area(ovk_ctx *c); /* A of the last draw, height*width */
/* the vertex-shader + viewport + snap chain: user space -> 24.8 window coordinates */
void ovk_transform_snap(const float m[6], uint32_t width, uint32_t height, const float *xy, uint64_t n, int32_t *out_xy);
/* sample positions (in 1/16 pixel) for a sample count; returns 0 if unsupported */
int ovk_sample_positions(uint32_t samples, int32_t *xy16);

/* rasterise a draw list recorded by oracle/_ref (ref_shim.h layout) onto the context's surface */
void ovk_raster_ref_drawlist(ovk_ctx *c, const void *draws, uint32_t n_draws, const uint8_t *blob);

#ifdef __cplusplus
}
#endif
#endif
