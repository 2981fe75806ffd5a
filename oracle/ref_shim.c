/* TEST INFRASTRUCTURE ONLY — recording back end for the UNMODIFIED reference
 * tessellation sources.
 *
 * oracle/Makefile compiles, from where they lie under /root/reference:
 *   src/vkvg_context.c, src/vkvg_context_internal.c, src/vkvg_pattern.c,
 *   src/vkvg_matrix.c, external/glutess/src/*.c
 * against oracle/ref_stubs/ and links them with this file into
 * oracle/_ref/libvkvg_ref.so.  This file supplies
 *   - the 17 Vulkan function-pointer globals (src/vkvg_device_internal.c:38-58)
 *     as recorders,
 *   - host-memory fakes of the vkh_* helpers the context uses,
 *   - a fake device / surface (struct layouts come from the reference headers),
 *   - no-op font cache entry points,
 * and turns every submitted command buffer into a flat, self-contained draw
 * list (`ref_draw_t`) with snapshots of the vertex / index / gradient buffers.
 * oracle/vkvg_oracle.c rasterises that list (ovk_raster_*), which gives
 * "reference geometry + restated Vulkan rasterisation" pixels.
 *
 * Nothing in the product library links or loads this.
 */
#include "vkvg_device_internal.h"
#include "vkvg_context_internal.h"
#include "vkvg_surface_internal.h"
#include "vkvg_pattern.h"
#include "ref_shim.h"

/* ------------------------------------------------------------------ */
/* command recording                                                  */
/* ------------------------------------------------------------------ */
enum { C_BIND_PIPE, C_DRAW, C_DRAW_IDX, C_CMP, C_REF, C_WRITE, C_BEGIN_RP, C_END_RP, C_SCISSOR, C_PUSH, C_CLEAR_ATT };

typedef struct {
    int      op;
    uint32_t a, b, c;
    int32_t  d;
    VkRect2D rect;
    uint32_t push_off, push_size;
    uint8_t  push[80];
} rec_cmd;

typedef struct {
    rec_cmd *cmds;
    uint32_t count, cap;
} rec_buf;

static void rec_push(VkCommandBuffer cb, rec_cmd c) {
    rec_buf *rb = (rec_buf *)cb;
    if (rb->count == rb->cap) {
        rb->cap  = rb->cap ? rb->cap * 2 : 256;
        rb->cmds = (rec_cmd *)realloc(rb->cmds, rb->cap * sizeof(rec_cmd));
    }
    rb->cmds[rb->count++] = c;
}

/* pipeline / renderpass handles are small integers */
#define H(x) ((void *)(uintptr_t)(x))

/* the single live fake device + the context currently recording (for buffer snapshots) */
static ref_drawlist_t g_list;
static vkvg_device   *g_dev;

/* buffers: remember the three context buffers by usage so that submit can snapshot them */
static vkh_buffer_t *g_vbo, *g_ibo, *g_ubo;
static int g_recording = 1; /* 0 while timing the reference tessellation as the CPU baseline */
void ref_set_recording(int on) { g_recording = on; }

static void f_CmdBindPipeline(VkCommandBuffer cb, VkPipelineBindPoint bp, VkPipeline p) {
    rec_cmd c = {.op = C_BIND_PIPE, .a = (uint32_t)(uintptr_t)p};
    rec_push(cb, c);
}
static void f_CmdBindDescriptorSets(VkCommandBuffer cb, VkPipelineBindPoint bp, VkPipelineLayout l, uint32_t a, uint32_t b,
                                    const VkDescriptorSet *s, uint32_t c, const uint32_t *d) {}
static void f_CmdBindIndexBuffer(VkCommandBuffer cb, VkBuffer b, VkDeviceSize o, VkIndexType t) {}
static void f_CmdBindVertexBuffers(VkCommandBuffer cb, uint32_t a, uint32_t b, const VkBuffer *p, const VkDeviceSize *o) {}
static void f_CmdDrawIndexed(VkCommandBuffer cb, uint32_t indexCount, uint32_t inst, uint32_t firstIndex, int32_t vertexOffset,
                             uint32_t fi) {
    rec_cmd c = {.op = C_DRAW_IDX, .a = indexCount, .b = firstIndex, .d = vertexOffset};
    rec_push(cb, c);
}
static void f_CmdDraw(VkCommandBuffer cb, uint32_t vertexCount, uint32_t inst, uint32_t firstVertex, uint32_t fi) {
    rec_cmd c = {.op = C_DRAW, .a = vertexCount, .b = firstVertex};
    rec_push(cb, c);
}
static void f_CmdSetStencilCompareMask(VkCommandBuffer cb, VkStencilFaceFlags f, uint32_t m) {
    rec_cmd c = {.op = C_CMP, .a = m};
    rec_push(cb, c);
}
static void f_CmdSetStencilReference(VkCommandBuffer cb, VkStencilFaceFlags f, uint32_t m) {
    rec_cmd c = {.op = C_REF, .a = m};
    rec_push(cb, c);
}
static void f_CmdSetStencilWriteMask(VkCommandBuffer cb, VkStencilFaceFlags f, uint32_t m) {
    rec_cmd c = {.op = C_WRITE, .a = m};
    rec_push(cb, c);
}
static void f_CmdBeginRenderPass(VkCommandBuffer cb, const VkRenderPassBeginInfo *bi, VkSubpassContents sc) {
    rec_cmd c = {.op = C_BEGIN_RP, .a = (uint32_t)(uintptr_t)bi->renderPass};
    rec_push(cb, c);
}
static void f_CmdEndRenderPass(VkCommandBuffer cb) {
    rec_cmd c = {.op = C_END_RP};
    rec_push(cb, c);
}
static void f_CmdSetViewport(VkCommandBuffer cb, uint32_t a, uint32_t b, const VkViewport *v) {}
static void f_CmdSetScissor(VkCommandBuffer cb, uint32_t a, uint32_t b, const VkRect2D *r) {
    rec_cmd c = {.op = C_SCISSOR, .rect = *r};
    rec_push(cb, c);
}
static void f_CmdPushConstants(VkCommandBuffer cb, VkPipelineLayout l, VkShaderStageFlags s, uint32_t off, uint32_t size,
                               const void *data) {
    rec_cmd c = {.op = C_PUSH, .push_off = off, .push_size = size};
    memcpy(c.push, data, size);
    rec_push(cb, c);
}
static VkResult f_WaitForFences(VkDevice d, uint32_t n, const VkFence *f, VkBool32 all, uint64_t t) { return VK_SUCCESS; }
static VkResult f_ResetFences(VkDevice d, uint32_t n, const VkFence *f) { return VK_SUCCESS; }
static VkResult f_ResetCommandBuffer(VkCommandBuffer cb, VkCommandBufferResetFlags f) {
    ((rec_buf *)cb)->count = 0;
    return VK_SUCCESS;
}

PFN_vkCmdBindPipeline          CmdBindPipeline          = f_CmdBindPipeline;
PFN_vkCmdBindDescriptorSets    CmdBindDescriptorSets    = f_CmdBindDescriptorSets;
PFN_vkCmdBindIndexBuffer       CmdBindIndexBuffer       = f_CmdBindIndexBuffer;
PFN_vkCmdBindVertexBuffers     CmdBindVertexBuffers     = f_CmdBindVertexBuffers;
PFN_vkCmdDrawIndexed           CmdDrawIndexed           = f_CmdDrawIndexed;
PFN_vkCmdDraw                  CmdDraw                  = f_CmdDraw;
PFN_vkCmdSetStencilCompareMask CmdSetStencilCompareMask = f_CmdSetStencilCompareMask;
PFN_vkCmdSetStencilReference   CmdSetStencilReference   = f_CmdSetStencilReference;
PFN_vkCmdSetStencilWriteMask   CmdSetStencilWriteMask   = f_CmdSetStencilWriteMask;
PFN_vkCmdBeginRenderPass       CmdBeginRenderPass       = f_CmdBeginRenderPass;
PFN_vkCmdEndRenderPass         CmdEndRenderPass         = f_CmdEndRenderPass;
PFN_vkCmdSetViewport           CmdSetViewport           = f_CmdSetViewport;
PFN_vkCmdSetScissor            CmdSetScissor            = f_CmdSetScissor;
PFN_vkCmdPushConstants         CmdPushConstants         = f_CmdPushConstants;
PFN_vkWaitForFences            WaitForFences            = f_WaitForFences;
PFN_vkResetFences              ResetFences              = f_ResetFences;
PFN_vkResetCommandBuffer       ResetCommandBuffer       = f_ResetCommandBuffer;

void vkCmdClearAttachments(VkCommandBuffer cb, uint32_t n, const VkClearAttachment *att, uint32_t nr, const VkClearRect *r) {
    uint32_t aspects = 0;
    for (uint32_t i = 0; i < n; i++)
        aspects |= att[i].aspectMask;
    rec_cmd c = {.op = C_CLEAR_ATT, .a = aspects};
    rec_push(cb, c);
}
void vkCmdCopyImage(VkCommandBuffer cb, VkImage a, VkImageLayout la, VkImage b, VkImageLayout lb, uint32_t n,
                    const VkImageCopy *r) {}
void vkUpdateDescriptorSets(VkDevice d, uint32_t n, const VkWriteDescriptorSet *w, uint32_t m, const VkCopyDescriptorSet *c) {}
VkResult vkCreateDescriptorPool(VkDevice d, const VkDescriptorPoolCreateInfo *ci, const VkAllocationCallbacks *a,
                                VkDescriptorPool *p) {
    *p = (VkDescriptorPool)H(1);
    return VK_SUCCESS;
}
VkResult vkAllocateDescriptorSets(VkDevice d, const VkDescriptorSetAllocateInfo *ai, VkDescriptorSet *s) {
    *s = (VkDescriptorSet)H(1);
    return VK_SUCCESS;
}
VkResult vkFreeDescriptorSets(VkDevice d, VkDescriptorPool p, uint32_t n, const VkDescriptorSet *s) { return VK_SUCCESS; }
void     vkDestroyDescriptorPool(VkDevice d, VkDescriptorPool p, const VkAllocationCallbacks *a) {}
void     vkDestroyFence(VkDevice d, VkFence f, const VkAllocationCallbacks *a) {}
void     vkFreeCommandBuffers(VkDevice d, VkCommandPool p, uint32_t n, const VkCommandBuffer *cbs) {
    for (uint32_t i = 0; i < n; i++) {
        rec_buf *rb = (rec_buf *)cbs[i];
        free(rb->cmds);
        free(rb);
    }
}
void     vkDestroyCommandPool(VkDevice d, VkCommandPool p, const VkAllocationCallbacks *a) {}
VkResult vkEndCommandBuffer(VkCommandBuffer cb) { return VK_SUCCESS; }

/* ------------------------------------------------------------------ */
/* vkh fakes                                                          */
/* ------------------------------------------------------------------ */
void vkh_buffer_init(VkhDevice dev, VkBufferUsageFlags usage, VkhMemoryUsage mem, VkDeviceSize size, vkh_buffer_t *buff,
                     bool mapped) {
    buff->pDev   = dev;
    buff->size   = size;
    buff->mapped = malloc(size);
    buff->buffer = (VkBuffer)buff;
    if (usage & VK_BUFFER_USAGE_VERTEX_BUFFER_BIT)
        g_vbo = buff;
    else if (usage & VK_BUFFER_USAGE_INDEX_BUFFER_BIT)
        g_ibo = buff;
    else
        g_ubo = buff;
}
void vkh_buffer_reset(vkh_buffer_t *buff) {
    free(buff->mapped);
    buff->mapped = NULL;
    if (g_vbo == buff) g_vbo = NULL;
    if (g_ibo == buff) g_ibo = NULL;
    if (g_ubo == buff) g_ubo = NULL;
}
void vkh_buffer_resize(vkh_buffer_t *buff, VkDeviceSize newSize, bool mapped) {
    buff->mapped = realloc(buff->mapped, newSize);
    buff->size   = newSize;
}
void *vkh_buffer_get_mapped_pointer(vkh_buffer_t *buff) { return buff->mapped; }
void  vkh_buffer_flush(vkh_buffer_t *buff) {}

void vkh_cmd_begin(VkCommandBuffer cmd, VkCommandBufferUsageFlags flags) { ((rec_buf *)cmd)->count = 0; }
void vkh_cmd_end(VkCommandBuffer cmd) {}
void vkh_cmd_buffs_create(VkhDevice dev, VkCommandPool pool, VkCommandBufferLevel level, uint32_t count, VkCommandBuffer *cmds) {
    for (uint32_t i = 0; i < count; i++)
        cmds[i] = (VkCommandBuffer)calloc(1, sizeof(rec_buf));
}
VkCommandPool vkh_cmd_pool_create(VkhDevice dev, uint32_t q, VkCommandPoolCreateFlags f) { return (VkCommandPool)H(1); }
void          vkh_cmd_label_start(VkCommandBuffer cmd, const char *name, const float color[4]) {}
void          vkh_cmd_label_end(VkCommandBuffer cmd) {}
void     vkh_cmd_submit_timelined(VkhQueue q, VkCommandBuffer *cmd, VkSemaphore s, uint64_t w, uint64_t sig) {}
void     vkh_cmd_submit_timelined2(VkhQueue q, VkCommandBuffer *cmd, VkSemaphore s[2], uint64_t w[2], uint64_t sig[2]) {}
VkResult vkh_timeline_wait(VkhDevice dev, VkSemaphore s, uint64_t v) { return VK_SUCCESS; }
void     vkh_device_set_object_name(VkhDevice dev, VkObjectType t, uint64_t h, const char *name) {}
VkFence  vkh_fence_create_signaled(VkhDevice dev) { return (VkFence)H(1); }
void     vkh_image_set_layout(VkCommandBuffer cmd, VkhImage img, VkImageAspectFlags aspect, VkImageLayout oldL, VkImageLayout newL,
                              VkPipelineStageFlags src, VkPipelineStageFlags dst) {
    if (img)
        img->layout = newL;
}
void     vkh_image_destroy(VkhImage img) {}
VkImage  vkh_image_get_vkimage(VkhImage img) { return img ? img->image : NULL; }
VkhImage vkh_image_ms_create(VkhDevice dev, VkFormat format, VkSampleCountFlags samples, uint32_t w, uint32_t h, VkhMemoryUsage mem,
                             VkImageUsageFlags usage) {
    return (VkhImage)calloc(1, sizeof(struct _vkh_image_t));
}
void vkh_image_create_sampler(VkhImage img, VkFilter mag, VkFilter min, VkSamplerMipmapMode mip, VkSamplerAddressMode addr) {}
VkDescriptorImageInfo vkh_image_get_descriptor(VkhImage img, VkImageLayout layout) {
    VkDescriptorImageInfo d = {0};
    return d;
}

/* tinycthread mutexes are never taken (threadAware == false) but must link */
int mtx_lock(mtx_t *m) { return 0; }
int mtx_unlock(mtx_t *m) { return 0; }

/* ------------------------------------------------------------------ */
/* font cache: text is out of scope             */
/* ------------------------------------------------------------------ */
_vkvg_font_identity_t *_font_cache_add_font_identity(VkvgContext ctx, const char *fontFile, const char *name) { return NULL; }
bool _font_cache_load_font_file_in_memory(_vkvg_font_identity_t *fontId) { return false; }
void _font_cache_show_text(VkvgContext ctx, const char *text) {}
void _font_cache_text_extents(VkvgContext ctx, const char *text, int length, vkvg_text_extents_t *extents) {}
void _font_cache_font_extents(VkvgContext ctx, vkvg_font_extents_t *extents) {}
void _font_cache_create_text_run(VkvgContext ctx, const char *text, int length, VkvgText textRun) {}
void _font_cache_destroy_text_run(VkvgText textRun) {}
void _font_cache_show_text_run(VkvgContext ctx, VkvgText tr) {}
void _font_cache_update_context_descset(VkvgContext ctx) {}

/* ------------------------------------------------------------------ */
/* device / surface fakes                                             */
/* ------------------------------------------------------------------ */
vkvg_status_t vkvg_device_status(VkvgDevice dev) { return dev ? dev->status : VKVG_STATUS_NULL_POINTER; }
vkvg_status_t vkvg_surface_status(VkvgSurface surf) { return surf ? surf->status : VKVG_STATUS_NULL_POINTER; }
VkvgSurface   vkvg_surface_reference(VkvgSurface surf) {
    surf->references++;
    return surf;
}
void vkvg_surface_destroy(VkvgSurface surf) {
    if (--surf->references > 0)
        return;
    free(surf->img);
    free(surf->imgMS);
    free(surf->stencil);
    free(surf);
}
bool _device_try_get_cached_context(VkvgDevice dev, VkvgContext *pCtx) { return false; }
void _device_store_context(VkvgContext ctx) {}

static void list_push_draw(ref_draw_t d) {
    if (g_list.n_draws == g_list.cap_draws) {
        g_list.cap_draws = g_list.cap_draws ? g_list.cap_draws * 2 : 256;
        g_list.draws     = (ref_draw_t *)realloc(g_list.draws, g_list.cap_draws * sizeof(ref_draw_t));
    }
    g_list.draws[g_list.n_draws++] = d;
}
static uint32_t list_push_bytes(const void *src, size_t n) {
    if (g_list.n_blob + n > g_list.cap_blob) {
        while (g_list.n_blob + n > g_list.cap_blob)
            g_list.cap_blob = g_list.cap_blob ? g_list.cap_blob * 2 : (1u << 20);
        g_list.blob = (uint8_t *)realloc(g_list.blob, g_list.cap_blob);
    }
    size_t off = g_list.n_blob;
    memcpy(g_list.blob + off, src, n);
    g_list.n_blob += (n + 15) & ~(size_t)15;
    if (g_list.n_blob > g_list.cap_blob) {
        g_list.cap_blob = g_list.n_blob;
        g_list.blob     = (uint8_t *)realloc(g_list.blob, g_list.cap_blob);
    }
    return (uint32_t)(off / 16);
}

/* CPU→GPU boundary of the reference (src/vkvg_device_internal.c:486-490): resolve the
 * recorded command buffer against the buffer contents it would have been executed with. */
void _device_submit_cmd(VkvgDevice dev, VkCommandBuffer *cmd, VkFence fence) {
    rec_buf *rb = (rec_buf *)*cmd;
    if (!g_recording)
        return;
    /* state carried inside one command buffer */
    uint32_t pipe = REF_PIPE_OVER, cmpMask = STENCIL_CLIP_BIT, ref = 0, writeMask = 0;
    VkRect2D scissor = {{0, 0}, {0, 0}};
    uint8_t  pc[80];
    memset(pc, 0, sizeof pc);

    /* lazily snapshot buffers only if the command buffer draws */
    uint32_t vb_off = 0, ib_off = 0, ub_off = 0;
    bool     snap = false;

    for (uint32_t i = 0; i < rb->count; i++) {
        rec_cmd *c = &rb->cmds[i];
        switch (c->op) {
        case C_BIND_PIPE: pipe = c->a; break;
        case C_CMP: cmpMask = c->a; break;
        case C_REF: ref = c->a; break;
        case C_WRITE: writeMask = c->a; break;
        case C_SCISSOR: scissor = c->rect; break;
        case C_PUSH: memcpy(pc + c->push_off, c->push, c->push_size); break;
        case C_BEGIN_RP: {
            ref_draw_t d = {.kind = REF_DRAW_BEGIN_PASS, .pipeline = c->a};
            list_push_draw(d);
            break;
        }
        case C_END_RP: {
            ref_draw_t d = {.kind = REF_DRAW_END_PASS};
            list_push_draw(d);
            break;
        }
        case C_CLEAR_ATT: {
            ref_draw_t d = {.kind = REF_DRAW_CLEAR, .first = c->a};
            list_push_draw(d);
            break;
        }
        case C_DRAW:
        case C_DRAW_IDX: {
            if (!snap) {
                snap = true;
                /* the reference memcpy's vertexCache/indexCache into the mapped buffers in
                 * _flush_vertices_caches (src/vkvg_context_internal.c:516-524) before submit */
                vb_off = list_push_bytes(g_vbo->mapped, g_vbo->size);
                ib_off = list_push_bytes(g_ibo->mapped, g_ibo->size);
                ub_off = list_push_bytes(g_ubo->mapped, sizeof(vkvg_gradient_t));
            }
            ref_draw_t d = {.kind      = c->op == C_DRAW ? REF_DRAW_ARRAYS : REF_DRAW_INDEXED,
                            .pipeline  = pipe,
                            .cmpMask   = cmpMask,
                            .ref       = ref,
                            .writeMask = writeMask,
                            .sc_x      = scissor.offset.x,
                            .sc_y      = scissor.offset.y,
                            .sc_w      = scissor.extent.width,
                            .sc_h      = scissor.extent.height,
                            .count     = c->a,
                            .first     = c->b,
                            .vertexOffset = c->d,
                            .vbo       = vb_off,
                            .ibo       = ib_off,
                            .ubo       = ub_off};
            memcpy(d.push, pc, 80);
            list_push_draw(d);
            break;
        }
        }
    }
}

/* ------------------------------------------------------------------ */
/* public shim API (oracle/ref_shim.h)                                */
/* ------------------------------------------------------------------ */
VkvgDevice ref_device_create(uint32_t samples) {
    vkvg_device *dev = (vkvg_device *)calloc(1, sizeof(vkvg_device));
    dev->status      = VKVG_STATUS_SUCCESS;
    dev->references  = 1;
    dev->samples     = samples;
    dev->threadAware = false;
    dev->hdpi = dev->vdpi    = 96;
    dev->pipe_OVER           = (VkPipeline)H(REF_PIPE_OVER);
    dev->pipe_SUB            = (VkPipeline)H(REF_PIPE_SUB);
    dev->pipe_CLEAR          = (VkPipeline)H(REF_PIPE_CLEAR);
    dev->pipelinePolyFill    = (VkPipeline)H(REF_PIPE_POLYFILL);
    dev->pipelineClipping    = (VkPipeline)H(REF_PIPE_CLIPPING);
    dev->renderPass          = (VkRenderPass)H(REF_RP_LOAD);
    dev->renderPass_ClearStencil = (VkRenderPass)H(REF_RP_CLEAR_STENCIL);
    dev->renderPass_ClearAll     = (VkRenderPass)H(REF_RP_CLEAR_ALL);
    dev->gQueue                  = (VkhQueue)calloc(1, sizeof(struct _vkh_queue_t));
    dev->emptyImg                = (VkhImage)calloc(1, sizeof(struct _vkh_image_t));
    /* do not park destroyed contexts in the per-thread cache (src/vkvg_context.c:295) */
    dev->cachedContextCount = VKVG_MAX_CACHED_CONTEXT_COUNT;
    g_dev                   = dev;
    return dev;
}
void ref_device_destroy(VkvgDevice dev) {
    free(dev->gQueue);
    free(dev->emptyImg);
    free(dev);
}
VkvgSurface ref_surface_create(VkvgDevice dev, uint32_t w, uint32_t h) {
    vkvg_surface *s = (vkvg_surface *)calloc(1, sizeof(vkvg_surface));
    s->status       = VKVG_STATUS_SUCCESS;
    s->references   = 1;
    s->dev          = dev;
    s->width        = w;
    s->height       = h;
    s->format       = VK_FORMAT_B8G8R8A8_UNORM;
    s->newSurf      = true; /* src/vkvg_surface.c:38-52: a new surface starts cleared */
    s->img          = (VkhImage)calloc(1, sizeof(struct _vkh_image_t));
    s->imgMS        = dev->samples > 1 ? (VkhImage)calloc(1, sizeof(struct _vkh_image_t)) : NULL;
    s->stencil      = (VkhImage)calloc(1, sizeof(struct _vkh_image_t));
    return s;
}
ref_drawlist_t *ref_drawlist(void) { return &g_list; }
void            ref_drawlist_reset(void) {
    g_list.n_draws = 0;
    g_list.n_blob  = 0;
}
/* host-array peek: flattened points + path table of the live context, before fill/stroke clears them */
uint32_t ref_ctx_points(VkvgContext ctx, const float **pts) {
    *pts = (const float *)ctx->points;
    return ctx->pointCount;
}
uint32_t ref_ctx_pathes(VkvgContext ctx, const uint32_t **pathes) {
    *pathes = ctx->pathes;
    /* make the open (unfinished) sub-path visible the way fill/stroke would (they call _finish_path) */
    return ctx->pathPtr;
}
void ref_ctx_finish_path(VkvgContext ctx) { _finish_path(ctx); }
/* vertex/index caches of the live context (stroke / non-zero output before any flush) */
uint32_t ref_ctx_vertices(VkvgContext ctx, const void **v) {
    *v = ctx->vertexCache;
    return ctx->vertCount;
}
uint32_t ref_ctx_indices(VkvgContext ctx, const uint32_t **idx) {
    *idx = ctx->indexCache;
    return ctx->indCount;
}
