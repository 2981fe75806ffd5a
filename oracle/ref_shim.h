/* TEST INFRASTRUCTURE ONLY — interface of oracle/_ref/libvkvg_ref.so beyond the
 * reference's own vkvg.h entry points (vkvg_create, vkvg_move_to, ... which the
 * library exports unchanged because they are the reference's own object code).
 * See oracle/ref_shim.c. */
#ifndef ORACLE_REF_SHIM_H
#define ORACLE_REF_SHIM_H
#include <stdint.h>

enum { REF_PIPE_OVER = 1, REF_PIPE_SUB, REF_PIPE_CLEAR, REF_PIPE_POLYFILL, REF_PIPE_CLIPPING };
enum { REF_RP_LOAD = 1, REF_RP_CLEAR_STENCIL, REF_RP_CLEAR_ALL };
enum { REF_DRAW_ARRAYS = 1, REF_DRAW_INDEXED, REF_DRAW_BEGIN_PASS, REF_DRAW_END_PASS, REF_DRAW_CLEAR };

/* One resolved Vulkan command.  vbo/ibo/ubo are offsets into `blob` in units of 16 bytes:
 *   vbo -> array of reference `Vertex` (24 B: float x,y; u32 colour; float uv[3])
 *   ibo -> array of u32 indices
 *   ubo -> vkvg_gradient_t (356 B scalar layout: 16 vec4 colours, 16 float stops, 2 vec4 cp, u32 count)
 * push = the 80-byte push-constant block (src/vkvg_context_internal.h:74-81). */
typedef struct {
    int32_t  kind;
    uint32_t pipeline; /* REF_PIPE_* (or REF_RP_* for BEGIN_PASS) */
    uint32_t cmpMask, ref, writeMask;
    int32_t  sc_x, sc_y;
    uint32_t sc_w, sc_h;
    uint32_t count, first;
    int32_t  vertexOffset;
    uint32_t vbo, ibo, ubo;
    uint8_t  push[80];
} ref_draw_t;

typedef struct {
    ref_draw_t *draws;
    uint32_t    n_draws, cap_draws;
    uint8_t    *blob;
    uint64_t    n_blob, cap_blob;
} ref_drawlist_t;

#ifndef ORACLE_REF_SHIM_NO_PROTOS
struct _vkvg_device_t;
struct _vkvg_surface_t;
struct _vkvg_context_t;
struct _vkvg_device_t  *ref_device_create(uint32_t samples);
void                    ref_device_destroy(struct _vkvg_device_t *dev);
struct _vkvg_surface_t *ref_surface_create(struct _vkvg_device_t *dev, uint32_t w, uint32_t h);
ref_drawlist_t         *ref_drawlist(void);
void                    ref_drawlist_reset(void);
void                    ref_set_recording(int on);
uint32_t                ref_ctx_points(struct _vkvg_context_t *ctx, const float **pts);
uint32_t                ref_ctx_pathes(struct _vkvg_context_t *ctx, const uint32_t **pathes);
void                    ref_ctx_finish_path(struct _vkvg_context_t *ctx);
uint32_t                ref_ctx_vertices(struct _vkvg_context_t *ctx, const void **v);
uint32_t                ref_ctx_indices(struct _vkvg_context_t *ctx, const uint32_t **idx);
#endif
#endif
