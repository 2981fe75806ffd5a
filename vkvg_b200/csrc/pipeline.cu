// Orchestration of one flush: upload -> flatten -> stroke / fill edges -> binning -> fine pass.
// Everything runs on one CUDA stream per device; device buffers are grow-only and reused between flushes.  A flush is queued
// without any host round trip (counts that only the device knows stay there: dev_util.cuh vkb_counts; overflowing a capacity
// makes the rest of the flush a no-op and the host replays it with room), and flushes that repeat the structure of the previous
// one are replayed as a CUDA graph (FlushKey).
#include "pipeline.h"
#include "renderer.h"
#include <string.h>
#include <vector>
#include <chrono>

unsigned long long g_vkb_launches = 0;
unsigned long long g_vkb_alloc_generation = 0;
// A CUDA error is charged to the device whose entry point is running on this thread (an out-of-memory on one device, or one
// oversized surface, must not put every other device of the process into VKVG_STATUS_DEVICE_ERROR); errors outside any entry
// point land in a process-wide flag that only vkb_device_open looks at.
struct vkb_device_impl;
static thread_local vkb_device_impl *t_dev = nullptr;
static int                           g_failed_nodev = 0;
static int  dev_failed(const vkb_device_impl *d);
static void dev_set_failed(vkb_device_impl *d);
void        vkb_note_cuda_error(cudaError_t) { if (t_dev) dev_set_failed(t_dev); else g_failed_nodev = 1; }
#define g_cuda_failed (t_dev ? dev_failed(t_dev) : g_failed_nodev)

#define VKB_MAX_BANDS 8
struct SurfFlags { bool known_clear, stencil_live; uint32_t stencil_samples; bool readback_valid; };  // surface state a flush attempt changes
struct vkb_device_impl;
static int finish_pending(vkb_device_impl *d);

struct vkb_device_impl {
    int          ordinal = 0;
    int          failed  = 0;  // sticky: a CUDA call made on behalf of this device failed
    cudaStream_t stream  = nullptr;
    cudaEvent_t  ev_begin = nullptr, ev_end = nullptr, ev_fine0 = nullptr, ev_fine1 = nullptr;
    cudaEvent_t  ev_stage[VKB_N_STAGES + 1] = {};  // boundaries between pipeline stages (profiling)
    cudaStream_t copy_stream = nullptr;             // read-back of finished bands of tile rows while later bands render (vkb_surface_set_readback)
    cudaEvent_t  ev_band[VKB_MAX_BANDS] = {}, ev_copied = nullptr;
    DevBuf       band_counters;
    DevBuf       l2_flush;
    // pinned staging
    uint8_t *stage     = nullptr;
    size_t   stage_cap = 0;
    // batch (device)
    DevBuf   elem_hdr, elem_data, subpaths, draws, xforms, strokes, grads, dashes, paints, fcnt, scnt, pcnt, srank, surfpats;
    cudaEvent_t ev_h2d = nullptr;
    DevBuf   fjob_draw, fjob_sp, sjob_draw, sjob_sp, sdraw_id, sdraw_first_job, sdraw_first_item, extra_edges, extra_edge_draw;
    uint32_t n_elems = 0, n_sp = 0, n_draws = 0, n_fjobs = 0, n_sjobs = 0, n_sdraws = 0, n_extra = 0;
    bool     any_dash = false;
    bool     has_clip_draws = false, has_stencil_ops = false;  // the uploaded batch holds VKB_DRAW_CLIP / any stencil-writing draw
    int      stencil_after = 0;                                // 0: batch leaves the stencil as it found it, 1: possibly non-zero, 2: all zero
    uint64_t h2d_bytes = 0;
    float    ms_host_upload = 0;
    // intermediates
    DevBuf elem_cnt, totals, pts, ptflags, sp_first, sp_count;
    DevBuf fjob_base, sjob_base, seglen, cum, item_counts, verts, inds, job_inverse;
    DevBuf edges, edge_draw;
    DevBuf draw_bbox, draw_rect, draw_counts, draw_ptbase, draw_rowbase;
    DevBuf pt_count, pt_backdrop, pt_flags, pt_draw, keys, vals, sorted_cnt, pt_slot, cursor, hdr, tile_first, tile_end, tile_edges;
    DevBuf winding, tmp_image, cursor2, flat_cache, pt_owner, row_owner, gprep, long_edges, snapped, wscratch, nz_mode, sp_bbox;
    DevBuf   dc_cmds, dc_args, dc_S, dc_blocksum, dc_lists, dc_small, dc_xfscale;  // command stream decoded on the device (decode.cu)
    uint32_t *dc_host = nullptr;  // pinned: totals row of the scan, irregular flag, census
    DevBuf   long_sp;  // per sub-path: the first 256-element block of the long ones (sp_bounds_long_k)
    bool   nz_any = false;  // the batch holds NON_ZERO fills / clips: they go through nz_classify / nz_split (raster.cu)
    uint32_t n_grads = 0;
    uint32_t n_curves = 0;  // cubic / arc elements in the resident batch
    uint32_t max_sp_elems = 0;  // elements of the longest sub-path of the resident batch
    ScanScratch scan;
    SortScratch sort;
    // counts that live on the device (dev_util.cuh: vkb_counts) and the flush that may still be in flight
    DevBuf            counts;
    vkb_counts       *counts_host = nullptr;  // pinned
    uint32_t          capv[16] = {};           // grow-only capacities
    bool              pending = false;
    vkb_surface_impl *pending_surf = nullptr;
    uint32_t          pending_samples = 0;
    SurfFlags         pending_before = {};
    // CUDA graph of one whole flush, reused while everything that shapes the launches stays the same (FlushKey)
    bool              capturing = false;
    bool              begin_recorded = false;  // ev_begin already sits in the stream (vkb_time_resident records it before the clear)
    bool              stage_timing = true;   // stats carry per-stage times (forces plain launches: events between kernels)
    bool              graphs_enabled = true;
    uint8_t           last_key[256] = {}, graph_key[256] = {};
    bool              have_last_key = false;
    cudaGraphExec_t   graph_exec = nullptr;
    unsigned long long graph_launches = 0;   // kernels in the cached graph (bench: gpu_launches)
    unsigned long long n_graph_replays = 0;
    SurfFlags         graph_after = {};       // surface flags a flush of the cached graph leaves behind
    bool              graph_fine_events = false;  // the cached graph records ev_fine0 / ev_fine1 as external event nodes
};
static int  dev_failed(const vkb_device_impl *d) { return d->failed; }
static void dev_set_failed(vkb_device_impl *d) { d->failed = 1; }
static inline void dev_enter(vkb_device_impl *d) { t_dev = d; cudaSetDevice(d->ordinal); }
#define VKB_EVENT_RECORD(d, ev) do { if (!(d)->capturing) VKB_CUDA_OK(cudaEventRecord((ev), (d)->stream)); } while (0)
struct vkb_surface_impl {
    vkb_device_impl *dev;
    uint32_t         w, h;
    uint32_t         full_h, origin_y;  // logical surface this one is a stripe of (full_h == h, origin_y == 0 otherwise)
    uint32_t         band_h = 0;        // batch surface: height of one canvas (full_h == band_h), 0 otherwise
    DevBuf           image;
    DevBuf           ms_image, tile_ms, ms_mask;  // per-sample plane + per-tile validity flags (allocated by the first render)
    DevBuf           stencil;                     // per-sample clip / save bits (allocated by the first flush that clips)
    bool             stencil_live = false;        // the plane may hold non-zero bytes
    uint32_t         stencil_samples = 0;         // sample count the plane was laid out for
    std::vector<DevBuf> stencil_spills;           // whole-plane copies, one per six nested clip saves
    bool             known_clear;
    uint8_t         *readback = nullptr;   // host memory every flush copies the finished image to, band by band (vkb_surface_set_readback)
    bool             readback_valid = false;  // *readback holds the image as the last flush left it
};

vkb_device_impl *vkb_device_open(int ordinal) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0) return nullptr;
    if (ordinal < 0 || ordinal >= n) ordinal = 0;
    if (cudaSetDevice(ordinal) != cudaSuccess) return nullptr;
    vkb_device_impl *d = new vkb_device_impl();
    d->ordinal         = ordinal;
    t_dev              = d;
    VKB_CUDA_OK(cudaStreamCreateWithFlags(&d->stream, cudaStreamNonBlocking));
    VKB_CUDA_OK(cudaEventCreate(&d->ev_begin));
    VKB_CUDA_OK(cudaEventCreate(&d->ev_end));
    VKB_CUDA_OK(cudaEventCreate(&d->ev_fine0));
    VKB_CUDA_OK(cudaEventCreate(&d->ev_fine1));
    VKB_CUDA_OK(cudaEventCreateWithFlags(&d->ev_h2d, cudaEventDisableTiming));
    VKB_CUDA_OK(cudaStreamCreateWithFlags(&d->copy_stream, cudaStreamNonBlocking));
    for (cudaEvent_t &e : d->ev_band) VKB_CUDA_OK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    VKB_CUDA_OK(cudaEventCreateWithFlags(&d->ev_copied, cudaEventDisableTiming));
    for (cudaEvent_t &e : d->ev_stage) VKB_CUDA_OK(cudaEventCreate(&e));
    VKB_CUDA_OK(cudaHostAlloc((void **)&d->counts_host, sizeof(vkb_counts), cudaHostAllocDefault));
    if (d->counts_host) memset(d->counts_host, 0, sizeof(vkb_counts));
    if (d->failed) { t_dev = nullptr; delete d; return nullptr; }
    return d;
}
void vkb_device_close(vkb_device_impl *d) {
    if (!d) return;
    dev_enter(d);
    finish_pending(d);
    cudaStreamSynchronize(d->stream);
    d->counts.release(); d->cursor2.release(); d->flat_cache.release(); d->pt_owner.release(); d->row_owner.release(); d->gprep.release(); d->surfpats.release(); d->long_edges.release(); d->snapped.release(); d->wscratch.release(); d->long_sp.release();
    d->dc_cmds.release(); d->dc_args.release(); d->dc_S.release(); d->dc_blocksum.release(); d->dc_lists.release(); d->dc_small.release(); d->dc_xfscale.release();
    if (d->dc_host) cudaFreeHost(d->dc_host); d->nz_mode.release(); d->sp_bbox.release();
    if (d->counts_host) cudaFreeHost(d->counts_host);
    DevBuf *bufs[] = {&d->sdraw_first_job, &d->xforms, &d->strokes, &d->fcnt, &d->scnt, &d->pcnt, &d->srank, &d->elem_hdr, &d->elem_data, &d->subpaths, &d->draws, &d->grads, &d->dashes, &d->paints, &d->fjob_draw, &d->fjob_sp, &d->sjob_draw,
                      &d->sjob_sp, &d->sdraw_id, &d->sdraw_first_item, &d->extra_edges, &d->extra_edge_draw, &d->elem_cnt, &d->totals, &d->pts, &d->ptflags,
                      &d->sp_first, &d->sp_count, &d->fjob_base, &d->sjob_base, &d->seglen, &d->cum, &d->item_counts, &d->verts, &d->inds, &d->job_inverse,
                      &d->edges, &d->edge_draw, &d->draw_bbox, &d->draw_rect, &d->draw_counts, &d->draw_ptbase, &d->draw_rowbase, &d->pt_count,
                      &d->pt_backdrop, &d->pt_flags, &d->pt_draw, &d->keys, &d->vals, &d->sorted_cnt, &d->pt_slot, &d->cursor, &d->hdr, &d->tile_first,
                      &d->tile_end, &d->tile_edges, &d->winding, &d->tmp_image, &d->scan.sums, &d->scan.ticket, &d->sort.hist, &d->sort.k2, &d->sort.v2, &d->sort.scan.sums, &d->sort.scan.ticket};
    for (DevBuf *b : bufs) b->release();
    if (d->stage) cudaFreeHost(d->stage);
    for (cudaEvent_t &e : d->ev_stage) cudaEventDestroy(e);
    if (d->graph_exec) cudaGraphExecDestroy(d->graph_exec);
    d->l2_flush.release();
    cudaEventDestroy(d->ev_h2d);
    for (cudaEvent_t &e : d->ev_band) cudaEventDestroy(e);
    cudaEventDestroy(d->ev_copied);
    if (d->copy_stream) cudaStreamDestroy(d->copy_stream);
    d->band_counters.release();
    cudaEventDestroy(d->ev_begin); cudaEventDestroy(d->ev_end); cudaEventDestroy(d->ev_fine0); cudaEventDestroy(d->ev_fine1);
    cudaStreamDestroy(d->stream);
    if (t_dev == d) t_dev = nullptr;
    delete d;
}
int  vkb_device_failed(vkb_device_impl *d) { return d->failed; }
void vkb_device_sync(vkb_device_impl *d) {
    dev_enter(d);
    finish_pending(d);
    VKB_CUDA_OK(cudaStreamSynchronize(d->stream));
}

vkb_surface_impl *vkb_surface_new(vkb_device_impl *d, uint32_t w, uint32_t h, uint32_t full_h, uint32_t origin_y) {
    dev_enter(d);
    vkb_surface_impl *s = new vkb_surface_impl();
    s->dev = d; s->w = w; s->h = h;
    s->full_h = full_h ? full_h : h; s->origin_y = origin_y;
    s->image.ensure((size_t)w * h * 4 + 16, d->stream);
    VKB_CUDA_OK(cudaMemsetAsync(s->image.p, 0, (size_t)w * h * 4, d->stream));
    s->known_clear = true;
    return s;
}
void vkb_surface_free(vkb_surface_impl *s) {
    if (!s) return;
    dev_enter(s->dev);
    finish_pending(s->dev);
    cudaStreamSynchronize(s->dev->stream);
    s->image.release();
    s->ms_image.release();
    s->tile_ms.release();
    s->ms_mask.release();
    s->stencil.release();
    for (DevBuf &b : s->stencil_spills) b.release();
    delete s;
}
void vkb_surface_clear(vkb_surface_impl *s) {
    dev_enter(s->dev);
    finish_pending(s->dev);
    if (!s->known_clear) VKB_CUDA_OK(cudaMemsetAsync(s->image.p, 0, (size_t)s->w * s->h * 4, s->dev->stream));
    s->known_clear = true;
    s->readback_valid = false;
    s->stencil_live = false;  // vkvg_clear wipes the stencil attachment too (src/vkvg_context.c:745-752)
}
void vkb_surface_set_band_height(vkb_surface_impl *s, uint32_t band_h) { s->band_h = band_h; }
void vkb_surface_stencil_reset(vkb_surface_impl *s) {
    if (s->stencil_live) finish_pending(s->dev);
    s->stencil_live = false;
}
static size_t stencil_bytes(const vkb_surface_impl *s, uint32_t samples) {
    const size_t tiles = (size_t)((s->w + VKB_TILE - 1) / VKB_TILE) * ((s->h + VKB_TILE - 1) / VKB_TILE);
    return tiles * 256 * ((((size_t)(samples ? samples : 1)) + 3) / 4) * 4;
}
// the reference parks the whole stencil image in a spare one every six nested clip saves and copies it back on the matching
// restore (src/vkvg_context.c:1268-1318, :1425-1470); same here with the stencil plane.  Both run after a flush.
int vkb_surface_stencil_push(vkb_surface_impl *s, uint32_t samples) {
    dev_enter(s->dev);
    finish_pending(s->dev);
    const size_t bytes = stencil_bytes(s, samples);
    s->stencil_spills.emplace_back();
    DevBuf &b = s->stencil_spills.back();
    b.ensure(bytes, s->dev->stream);
    if (s->stencil_live && s->stencil.p) VKB_CUDA_OK(cudaMemcpyAsync(b.p, s->stencil.p, bytes, cudaMemcpyDeviceToDevice, s->dev->stream));
    else VKB_CUDA_OK(cudaMemsetAsync(b.p, 0, bytes, s->dev->stream));
    return g_cuda_failed;
}
int vkb_surface_stencil_pop(vkb_surface_impl *s, uint32_t samples) {
    dev_enter(s->dev);
    finish_pending(s->dev);
    if (s->stencil_spills.empty()) return 1;
    const size_t bytes = stencil_bytes(s, samples);
    s->stencil.ensure(bytes, s->dev->stream);
    VKB_CUDA_OK(cudaMemcpyAsync(s->stencil.p, s->stencil_spills.back().p, bytes, cudaMemcpyDeviceToDevice, s->dev->stream));
    VKB_CUDA_OK(cudaStreamSynchronize(s->dev->stream));
    s->stencil_spills.back().release();
    s->stencil_spills.pop_back();
    s->stencil_live = true;
    return g_cuda_failed;
}
const uint32_t *vkb_surface_device_pixels(vkb_surface_impl *s) { return s->image.as<uint32_t>(); }
// replace the contents of the surface by width*height premultiplied RGBA8 pixels from the host (vkvg_surface_create_from_bitmap)
int vkb_surface_upload(vkb_surface_impl *s, const uint8_t *rgba) {
    dev_enter(s->dev);
    finish_pending(s->dev);
    VKB_CUDA_OK(cudaMemcpyAsync(s->image.p, rgba, (size_t)s->w * s->h * 4, cudaMemcpyHostToDevice, s->dev->stream));
    VKB_CUDA_OK(cudaStreamSynchronize(s->dev->stream));
    s->known_clear = false;
    if (s->tile_ms.p) VKB_CUDA_OK(cudaMemsetAsync(s->tile_ms.p, 0, (size_t)((s->w + VKB_TILE - 1) / VKB_TILE) * ((s->h + VKB_TILE - 1) / VKB_TILE), s->dev->stream));
    return g_cuda_failed;
}
// device-to-device copy of the premultiplied pixels (e.g. into a tensor handed to an NCCL gather); synchronous
int vkb_surface_copy_to_device(vkb_surface_impl *s, void *dst) {
    dev_enter(s->dev);
    finish_pending(s->dev);
    VKB_CUDA_OK(cudaMemcpyAsync(dst, s->image.p, (size_t)s->w * s->h * 4, cudaMemcpyDeviceToDevice, s->dev->stream));
    VKB_CUDA_OK(cudaStreamSynchronize(s->dev->stream));
    return g_cuda_failed;
}
// Host memory (pinned, width * height * 4 bytes) that every later flush onto s also delivers the premultiplied image to: the fine pass then
// runs in bands of tile rows and each finished band is copied on a second stream while the next ones render, so that reading the surface
// back costs the copy of the last band instead of the whole image after the whole frame.  vkb_surface_download(s, that pointer, false)
// then only waits.  NULL: off.
void vkb_surface_set_readback(vkb_surface_impl *s, uint8_t *host) {
    dev_enter(s->dev);
    finish_pending(s->dev);
    s->readback = host;
    s->readback_valid = false;
}
// The image of a surface as an inter-process handle (cudaIpcMemHandle_t, 64 bytes): another process - another GPU of the node - opens it and
// names the returned pointer as the read-back target of its own stripe surface, whose bands then travel over NVLink while later bands render.
int vkb_surface_ipc_export(vkb_surface_impl *s, void *handle64) {
    dev_enter(s->dev);
    if (finish_pending(s->dev)) return 1;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size");
    cudaIpcMemHandle_t h;
    if (cudaIpcGetMemHandle(&h, s->image.p) != cudaSuccess) { cudaGetLastError(); return 1; }
    memcpy(handle64, &h, 64);
    return 0;
}
void *vkb_ipc_open(vkb_device_impl *d, const void *handle64) {
    dev_enter(d);
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    void *p = nullptr;
    if (cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return p;
}
int vkb_ipc_close(vkb_device_impl *d, void *p) {
    dev_enter(d);
    finish_pending(d);
    cudaStreamSynchronize(d->copy_stream);
    if (cudaIpcCloseMemHandle(p) != cudaSuccess) { cudaGetLastError(); return 1; }
    return 0;
}
int vkb_surface_download(vkb_surface_impl *s, uint8_t *out, bool unpremultiply) {
    vkb_device_impl *d = s->dev;
    dev_enter(d);
    if (finish_pending(d)) return 1;
    if (!unpremultiply && out && out == s->readback && s->readback_valid) {  // delivered by the flush itself
        VKB_CUDA_OK(cudaStreamSynchronize(d->stream));
        return g_cuda_failed;
    }
    size_t          bytes = (size_t)s->w * s->h * 4;
    const uint32_t *src   = s->image.as<uint32_t>();
    if (unpremultiply) {
        d->tmp_image.ensure(bytes, d->stream);
        vkb_launch_unpremultiply(src, (uint64_t)s->w * s->h, d->tmp_image.as<uint32_t>(), d->stream);
        src = d->tmp_image.as<uint32_t>();
    }
    VKB_CUDA_OK(cudaMemcpyAsync(out, src, bytes, cudaMemcpyDefault, d->stream));  // (out: host memory, or device memory of any GPU in reach)
    VKB_CUDA_OK(cudaStreamSynchronize(d->stream));
    return g_cuda_failed;
}

// ---- pending (asynchronous) flush bookkeeping ----
static SurfFlags surf_flags(const vkb_surface_impl *s) { return SurfFlags{s->known_clear, s->stencil_live, s->stencil_samples, s->readback_valid}; }
static void      surf_restore(vkb_surface_impl *s, SurfFlags f) { s->known_clear = f.known_clear; s->stencil_live = f.stencil_live; s->stencil_samples = f.stencil_samples; s->readback_valid = f.readback_valid; }
static int  run_flush(vkb_device_impl *d, vkb_surface_impl *surf, uint32_t samples, vkb_capture *cap, vkb_stats *stats, bool allow_async);

// ---- upload ----
static uint8_t *stage_reserve(vkb_device_impl *d, size_t bytes) {
    if (bytes > d->stage_cap) {
        if (d->stage) {
            cudaStreamSynchronize(d->stream);
            cudaFreeHost(d->stage);
        }
        size_t cap = d->stage_cap ? d->stage_cap : (1u << 20);
        while (cap < bytes) cap *= 2;
        VKB_CUDA_OK(cudaHostAlloc((void **)&d->stage, cap, cudaHostAllocDefault));
        d->stage_cap = cap;
    }
    return d->stage;
}

// ---- per-draw tables built on the device (the host only uploads what it recorded) ----
__global__ void draw_tables_k(const vkb_draw *draws, uint32_t n, vkb_paint *paints, uint32_t *fcnt, uint32_t *scnt, uint32_t *pcnt, uint32_t *is_sdraw) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    vkb_draw d = draws[i];
    paints[i]  = vkb_paint{d.rule_pattern, d.color, d.opacity, d.gradient};
    fcnt[i]    = (d.kind == VKB_DRAW_FILL || d.kind == VKB_DRAW_CLIP) ? d.n_subpaths : 0u;
    scnt[i]    = d.kind == VKB_DRAW_STROKE ? d.n_subpaths : 0u;
    pcnt[i]    = (d.kind == VKB_DRAW_PAINT || d.kind == VKB_DRAW_STENCIL) ? 1u : 0u;  // both cover the surface with a rectangle
    is_sdraw[i] = (d.kind == VKB_DRAW_STROKE && d.n_subpaths) ? 1u : 0u;
}
// job j of a kind = (draw, sub-path): the draw is the last one whose exclusive job base is <= j and that has jobs
__global__ void expand_jobs_k(const vkb_draw *draws, const uint32_t *base, uint32_t n_draws, uint32_t n_jobs, uint32_t *job_draw, uint32_t *job_sp) {
    uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_jobs) return;
    uint32_t lo = 0, hi = n_draws;
    while (hi - lo > 1) {
        uint32_t mid = (lo + hi) >> 1;
        if (base[mid] <= j) lo = mid; else hi = mid;
    }
    job_draw[j] = lo;
    job_sp[j]   = draws[lo].first_subpath + (j - base[lo]);
}
// stroke draws in order (ids + their first job) and whole-surface draws (ids, four rectangle edges each)
__global__ void list_draws_k(const vkb_draw *draws, const uint32_t *sbase, const uint32_t *pbase, uint32_t n, uint32_t *sdraw_id, uint32_t *sdraw_first_job,
                             const uint32_t *sdraw_rank, uint32_t *extra_edge_draw) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    vkb_draw d = draws[i];
    if (d.kind == VKB_DRAW_STROKE && d.n_subpaths) {
        uint32_t r = sdraw_rank[i];
        sdraw_id[r] = i; sdraw_first_job[r] = sbase[i];
    }
    if (d.kind == VKB_DRAW_PAINT || d.kind == VKB_DRAW_STENCIL) {
        uint32_t r = pbase[i];
        for (int k = 0; k < 4; k++) extra_edge_draw[4 * r + k] = i;
    }
}

// per-draw tables built on the device from d->draws (d->n_draws, n_fjobs, n_sjobs, n_sdraws, n_extra already known to the host)
static void build_job_tables(vkb_device_impl *d) {
    cudaStream_t st = d->stream;
    // ---- job tables on the device ----
    const uint32_t nd = d->n_draws;
    d->paints.ensure((size_t)(nd + 1) * sizeof(vkb_paint), st);
    d->fcnt.ensure((size_t)(nd + 1) * 4, st); d->scnt.ensure((size_t)(nd + 1) * 4, st); d->pcnt.ensure((size_t)(nd + 1) * 4, st);
    d->srank.ensure((size_t)(nd + 1) * 4, st);
    d->totals.ensure(16 * 8, st);
    if (nd) {
        draw_tables_k<<<vkb_div_up(nd, 256), 256, 0, st>>>(d->draws.as<vkb_draw>(), nd, d->paints.as<vkb_paint>(), d->fcnt.as<uint32_t>(),
                                                          d->scnt.as<uint32_t>(), d->pcnt.as<uint32_t>(), d->srank.as<uint32_t>());
        VKB_LAUNCHED();
        if (d->n_fjobs) vkb_exclusive_scan<uint32_t, uint32_t>(d->fcnt.as<uint32_t>(), d->fcnt.as<uint32_t>(), nd, nullptr, d->scan, st);
        if (d->n_sjobs) {
            vkb_exclusive_scan<uint32_t, uint32_t>(d->scnt.as<uint32_t>(), d->scnt.as<uint32_t>(), nd, nullptr, d->scan, st);
            vkb_exclusive_scan<uint32_t, uint32_t>(d->srank.as<uint32_t>(), d->srank.as<uint32_t>(), nd, nullptr, d->scan, st);
        }
        if (d->n_extra) vkb_exclusive_scan<uint32_t, uint32_t>(d->pcnt.as<uint32_t>(), d->pcnt.as<uint32_t>(), nd, nullptr, d->scan, st);
        d->fjob_draw.ensure((size_t)(d->n_fjobs + 1) * 4, st); d->fjob_sp.ensure((size_t)(d->n_fjobs + 1) * 4, st);
        d->sjob_draw.ensure((size_t)(d->n_sjobs + 1) * 4, st); d->sjob_sp.ensure((size_t)(d->n_sjobs + 1) * 4, st);
        d->sdraw_id.ensure((size_t)(d->n_sdraws + 1) * 4, st); d->sdraw_first_job.ensure((size_t)(d->n_sdraws + 1) * 4, st);
        d->extra_edge_draw.ensure((size_t)d->n_extra * 4 + 16, st);
        if (d->n_fjobs) {
            expand_jobs_k<<<vkb_div_up(d->n_fjobs, 256), 256, 0, st>>>(d->draws.as<vkb_draw>(), d->fcnt.as<uint32_t>(), nd, d->n_fjobs,
                                                                      d->fjob_draw.as<uint32_t>(), d->fjob_sp.as<uint32_t>());
            VKB_LAUNCHED();
        }
        if (d->n_sjobs) {
            expand_jobs_k<<<vkb_div_up(d->n_sjobs, 256), 256, 0, st>>>(d->draws.as<vkb_draw>(), d->scnt.as<uint32_t>(), nd, d->n_sjobs,
                                                                      d->sjob_draw.as<uint32_t>(), d->sjob_sp.as<uint32_t>());
            VKB_LAUNCHED();
        }
        if (d->n_sdraws || d->n_extra) {
            list_draws_k<<<vkb_div_up(nd, 256), 256, 0, st>>>(d->draws.as<vkb_draw>(), d->scnt.as<uint32_t>(), d->pcnt.as<uint32_t>(), nd,
                                                             d->sdraw_id.as<uint32_t>(), d->sdraw_first_job.as<uint32_t>(), d->srank.as<uint32_t>(),
                                                             d->extra_edge_draw.as<uint32_t>());
            VKB_LAUNCHED();
        }
    }
}

int vkb_upload(vkb_device_impl *d, const vkb_batch &b) {
    const auto t_begin = std::chrono::steady_clock::now();
    dev_enter(d);
    cudaStream_t st = d->stream;
    finish_pending(d);
    // the previous flush may still be reading the staging area
    VKB_CUDA_OK(cudaStreamSynchronize(st));
    d->n_elems = (uint32_t)b.elem_hdr.size(); d->n_sp = (uint32_t)b.subpaths.size(); d->n_draws = (uint32_t)b.draws.size();
    d->n_curves = b.n_curves;
    d->max_sp_elems = 0;
    for (const vkb_subpath &sp : b.subpaths) if (sp.n_elems > d->max_sp_elems) d->max_sp_elems = sp.n_elems;
    d->n_grads  = (uint32_t)b.grads.size();
    // what the job tables will hold is a function of the recorded draws alone: counted here, no read-back
    d->has_clip_draws = d->has_stencil_ops = false;
    d->stencil_after = 0;
    d->n_fjobs = d->n_sjobs = d->n_sdraws = d->n_extra = 0;
    d->any_dash = false;
    d->nz_any   = false;
    for (const vkb_draw &dr : b.draws) {
        if (dr.n_subpaths && ((dr.kind == VKB_DRAW_FILL && (dr.rule_pattern & 0xFF) == VKB_RULE_NON_ZERO) || (dr.kind == VKB_DRAW_CLIP && (dr.rule_pattern & 0xFF) == VKB_RULE_CLIP_NZ))) d->nz_any = true;
        if (dr.kind == VKB_DRAW_CLIP) { d->has_clip_draws = d->has_stencil_ops = true; d->stencil_after = 1; }
        else if (dr.kind == VKB_DRAW_STENCIL) {
            d->has_stencil_ops = true;
            d->stencil_after   = (dr.rule_pattern & 0xFF) == VKB_RULE_ST_CLEAR ? 2 : 1;
        }
        if (dr.kind == VKB_DRAW_FILL || dr.kind == VKB_DRAW_CLIP) d->n_fjobs += dr.n_subpaths;
        else if (dr.kind == VKB_DRAW_STROKE) {
            d->n_sjobs += dr.n_subpaths;
            if (dr.n_subpaths) {
                d->n_sdraws++;
                if (b.strokes[dr.xform_stroke >> 16].dash_count) d->any_dash = true;
            }
        } else d->n_extra += 4;
    }

    struct Src { DevBuf *dst; const void *p; size_t bytes; bool pinned; };
    Src srcs[] = {
        {&d->elem_hdr, b.elem_hdr.data(), b.elem_hdr.size() * 4, true},   // recorded straight into pinned memory: no staging copy
        {&d->elem_data, b.elem_data.data(), b.elem_data.size() * 4, true},
        {&d->subpaths, b.subpaths.data(), b.subpaths.size() * sizeof(vkb_subpath), false},
        {&d->draws, b.draws.data(), b.draws.size() * sizeof(vkb_draw), false},
        {&d->xforms, b.xforms.data(), b.xforms.size() * sizeof(vkb_xform), false},
        {&d->strokes, b.strokes.data(), b.strokes.size() * sizeof(vkb_stroke), false},
        {&d->grads, b.grads.data(), b.grads.size() * sizeof(vkb_gradient), false},
        {&d->dashes, b.dashes.data(), b.dashes.size() * 4, false},
        {&d->surfpats, b.surfpats.data(), b.surfpats.size() * sizeof(vkb_surfpat), false},
    };
    size_t total = 0;
    for (Src &s : srcs) if (!s.pinned) total += (s.bytes + 255) & ~(size_t)255;
    uint8_t *stg = stage_reserve(d, total + 256);
    size_t   off = 0;
    d->h2d_bytes = 0;
    for (Src &s : srcs) {
        s.dst->ensure(s.bytes + 16, st);
        d->h2d_bytes += s.bytes;
        if (!s.bytes) continue;
        if (s.pinned) {
            VKB_CUDA_OK(cudaMemcpyAsync(s.dst->p, s.p, s.bytes, cudaMemcpyHostToDevice, st));
        } else {
            memcpy(stg + off, s.p, s.bytes);
            VKB_CUDA_OK(cudaMemcpyAsync(s.dst->p, stg + off, s.bytes, cudaMemcpyHostToDevice, st));
            off += (s.bytes + 255) & ~(size_t)255;
        }
    }
    VKB_CUDA_OK(cudaEventRecord(d->ev_h2d, st));

    build_job_tables(d);
    d->ms_host_upload = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t_begin).count();
    return g_cuda_failed;
}

// pinned host memory for the recorder's element arrays (PodVec in renderer.h)
void *vkb_host_alloc(size_t bytes) {
    void *p = nullptr;
    if (cudaHostAlloc(&p, bytes, cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); return malloc(bytes); }
    return p;
}
void vkb_host_free(void *p) {
    if (!p) return;
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) == cudaSuccess && at.type == cudaMemoryTypeHost) cudaFreeHost(p);
    else { cudaGetLastError(); free(p); }
}

template <class T> static void download(vkb_device_impl *d, std::vector<T> *out, const void *src, size_t n) {
    if (!out) return;
    out->resize(n);
    if (n) VKB_CUDA_OK(cudaMemcpyAsync(out->data(), src, n * sizeof(T), cudaMemcpyDeviceToHost, d->stream));
    VKB_CUDA_OK(cudaStreamSynchronize(d->stream));
}

// the four edges of a rectangle one tile larger than the surface for every whole-surface draw; they follow the fill edges
__global__ void extra_rect_edges_k(vkb_edge *edges, uint32_t *edge_draw, const uint32_t *extra_edge_draw, uint32_t n_rects, int32_t W, int32_t H,
                                   const vkb_counts *C, int32_t *bbox) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_rects || C->overflow) return;
    const uint32_t base = C->n[VKC_FEDGES];  // (before the stroke edges, whose stored number is only known after tri_edges_k)
    vkb_edge      *e    = edges + base;
    const int32_t  x0 = -VKB_TILE_FX, y0 = -VKB_TILE_FX, x1 = W * 256 + VKB_TILE_FX, y1 = H * 256 + VKB_TILE_FX;
    e[4 * i]     = vkb_edge{x0, y0, x1, y0};
    e[4 * i + 1] = vkb_edge{x1, y0, x1, y1};
    e[4 * i + 2] = vkb_edge{x1, y1, x0, y1};
    e[4 * i + 3] = vkb_edge{x0, y1, x0, y0};
    for (int k = 0; k < 4; k++) edge_draw[base + 4 * i + k] = extra_edge_draw[4 * i + k];
    int32_t *b = bbox + 4 * (size_t)extra_edge_draw[4 * i];   // (one rectangle per whole-surface draw: nobody else writes its box)
    b[0] = x0; b[1] = y0; b[2] = x1; b[3] = y1;
}

// ---- commit kernels: one thread turns the raw totals of a producing scan into checked counts (dev_util.cuh: vkb_counts) ----
__global__ void commit_flatten_k(vkb_counts *C, const uint64_t *totals, uint32_t has_fill, uint32_t has_stroke, uint32_t split) {
    if (C->overflow) return;
    vkc_commit(C, VKC_POINTS, (uint32_t)totals[0]);
    vkc_commit(C, VKC_FILL, has_fill ? (uint32_t)totals[1] : 0u);
    if (!split) vkc_commit(C, VKC_FEDGES, has_fill ? (uint32_t)totals[1] : 0u);  // (else the scan of the split counts commits it)
    vkc_commit(C, VKC_SITEMS, has_stroke ? (uint32_t)totals[2] : 0u);
}
// (its other threads gather the first work item of every stroke draw - what maps a triangle back to its draw in tri_edges_k)
__global__ void commit_stroke_k(vkb_counts *C, const uint64_t *totals, uint32_t has_stroke, uint32_t n_extra, const uint32_t *first_job, const uint32_t *job_base,
                                uint32_t n_sdraws, uint32_t *first_item) {
    if (C->overflow) return;
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_sdraws) first_item[i] = job_base[first_job[i]];
    if (i) return;
    const unsigned long long tot = has_stroke ? totals[3] : 0ull;
    const uint32_t nv = (uint32_t)(tot & 0xffffffffull), ni = (uint32_t)(tot >> 32);
    vkc_commit(C, VKC_VERTS, nv);
    vkc_commit(C, VKC_INDS, ni);
    vkc_commit(C, VKC_TRIS, ni / 3);
    vkc_commit(C, VKC_EDGES, C->n[VKC_FEDGES] + 3u * (ni / 3) + n_extra);
}
__global__ void set_edge_count_k(vkb_counts *C, uint32_t n) {  // raw edge lists (vkb_winding_raw)
    for (int i = 0; i < VKC_N; i++) { C->n[i] = 0; C->need[i] = 0; }
    C->overflow = 0;
    C->n[VKC_EDGES] = n;
}

// Capacities are grow-only per device: whatever an earlier flush needed (plus slack) is what the next one is launched
// with, so a steady stream of similar frames never overflows; the first flush of a new kind starts from guesses.
static void plan_caps(vkb_device_impl *d, const SurfaceDesc &sd) {
    uint32_t *c = d->capv;
    auto up = [&](int i, uint64_t v) { if (v > 0xfffffff0ull) v = 0xfffffff0ull; if (c[i] < v) c[i] = (uint32_t)v; };
    up(VKC_POINTS, (uint64_t)d->n_elems * 2 + 4096);
    if (d->n_fjobs) up(VKC_FILL, (uint64_t)c[VKC_POINTS]);
    if (d->n_sjobs) {
        up(VKC_SITEMS, (uint64_t)c[VKC_POINTS]);
        up(VKC_VERTS, (uint64_t)c[VKC_SITEMS] * 6 + 64);
        up(VKC_INDS, (uint64_t)c[VKC_SITEMS] * 18 + 192);
    }
    up(VKC_TRIS, (uint64_t)c[VKC_INDS] / 3);
    up(VKC_FEDGES, (uint64_t)c[VKC_FILL] + (d->nz_any ? (uint64_t)c[VKC_FILL] / 2 + 64 : 0));  // split NON_ZERO draws: a guess, grown on overflow
    up(VKC_EDGES, (uint64_t)c[VKC_FEDGES] + 3ull * c[VKC_TRIS] + d->n_extra);
    const uint64_t n_tiles = (uint64_t)sd.tiles_x * sd.tiles_y;
    up(VKC_PT, (uint64_t)d->n_draws * 8 + ((uint64_t)d->n_extra / 4 + (d->has_clip_draws ? 2 : 0)) * n_tiles + 1024);
    up(VKC_ROWS, c[VKC_PT]);
    up(VKC_NE, c[VKC_PT]);   // a subset of the path-tiles: cannot overflow on its own
    up(VKC_TE, (uint64_t)c[VKC_EDGES] * 2 + 1024);
}
static void grow_caps_from_need(vkb_device_impl *d, const vkb_counts &h) {
    for (int i = 0; i < VKC_N; i++)
        if (h.overflow & (1u << i)) {
            uint64_t v = (uint64_t)h.need[i] + h.need[i] / 4 + 64;
            if (v > 0xfffffff0ull) v = 0xfffffff0ull;
            if (d->capv[i] < v) d->capv[i] = (uint32_t)v;
        }
}

// binning + fine pass over d->edges / d->edge_draw (count C->n[VKC_EDGES], nd draws, paints / grads already on the device)
static void enqueue_bin_and_fine(vkb_device_impl *d, vkb_surface_impl *surf, SurfaceDesc sd, uint32_t nd, vkb_capture *cap, const vkb_draw *draws,
                                 DevBuf &wbuf, const uint32_t *live_edges = nullptr) {  // draws: device pointer, or null for a raw edge list (no clip draws then)
    cudaStream_t st     = d->stream;
    uint64_t    *totals = d->totals.as<uint64_t>();
    vkb_counts  *C      = d->counts.as<vkb_counts>();
    const uint32_t *cv  = d->capv;
    const uint32_t samples = sd.samples;
    vkb_edge *edges = d->edges.as<vkb_edge>();
    uint32_t *edraw = d->edge_draw.as<uint32_t>();
    // ---- 5. binning ----
    VKB_EVENT_RECORD(d, d->ev_stage[3]);
    const uint32_t n_tiles = sd.tiles_x * sd.tiles_y;
    d->draw_bbox.ensure((size_t)nd * 16, st);
    d->draw_rect.ensure((size_t)nd * 16, st);
    d->draw_counts.ensure((size_t)(nd + 1) * 8, st);
    d->draw_ptbase.ensure((size_t)(nd + 1) * 4, st);
    d->draw_rowbase.ensure((size_t)(nd + 1) * 4, st);
    if (d->failed) return;  // an allocation failed: nothing that would use the buffer is launched
    if (!draws) vkb_launch_draw_bbox(edges, edraw, cv[VKC_EDGES], C, nd, d->draw_bbox.as<int32_t>(), st);  // (else grown by the kernels that emitted the edges)
    unsigned long long *dc = d->draw_counts.as<unsigned long long>();
    d->gprep.ensure((size_t)(d->n_grads + 1) * 16 * 4, st);
    if (d->failed) return;  // an allocation failed: nothing that would use the buffer is launched
    const GradPrep gp = {d->grads.as<vkb_gradient>(), draws ? d->n_grads : 0u, (float)sd.width, (float)sd.full_height, d->gprep.as<float>()};
    const bool clip_draws = draws && d->has_clip_draws, stencil_ops = draws && d->has_stencil_ops;
    vkb_launch_draw_rects(d->draw_bbox.as<int32_t>(), (clip_draws || sd.band_tiles) ? draws : nullptr, d->xforms.as<vkb_xform>(), nd, sd,
                          d->draw_rect.as<int32_t>(), dc, C, live_edges, d->n_extra, gp, st);   // (+ the live stroke edges committed as the edge count, the gradients prepared)
    vkb_exclusive_scan<unsigned long long, unsigned long long>(dc, dc, nd, (unsigned long long *)(totals + 4), d->scan, st);
    vkb_launch_split_bases(dc, nd, d->draw_ptbase.as<uint32_t>(), d->draw_rowbase.as<uint32_t>(), C, (const unsigned long long *)(totals + 4), st);   // (+ VKC_PT / VKC_ROWS committed)

    const uint32_t cap_pt = cv[VKC_PT], cap_ne = cv[VKC_NE];
    d->pt_count.ensure((size_t)(cap_pt + 1) * 4, st);
    d->pt_backdrop.ensure((size_t)(cap_pt + 1) * 4, st);
    d->pt_flags.ensure((size_t)(cap_pt + 1) * 4, st);
    d->pt_slot.ensure((size_t)(cap_pt + 1) * 4, st);
    d->pt_owner.ensure((size_t)(cap_pt + 1) * 4, st);
    d->row_owner.ensure((size_t)(cv[VKC_ROWS] + 1) * 4, st);
    if (d->failed) return;  // an allocation failed: nothing that would use the buffer is launched
    // small frames: owners_k zeroes the counts and backdrops of its path-tiles itself (two graph nodes less); large ones keep the memsets, which run at
    // memory speed where a draw as large as the surface would have ONE warp zero a million path-tiles (C5a: binning 2.07 -> 2.46 ms without this)
    const bool zero_in_owners = cap_pt <= (1u << 18);
    vkb_launch_owners(d->draw_rect.as<int32_t>(), d->draw_ptbase.as<uint32_t>(), d->draw_rowbase.as<uint32_t>(), nd, C, d->pt_owner.as<uint32_t>(),
                      d->row_owner.as<uint32_t>(), zero_in_owners ? d->pt_count.as<uint32_t>() : nullptr, zero_in_owners ? d->pt_backdrop.as<int32_t>() : nullptr, st);
    if (!zero_in_owners) {
        VKB_CUDA_OK(cudaMemsetAsync(d->pt_count.p, 0, (size_t)cap_pt * 4, st));
        VKB_CUDA_OK(cudaMemsetAsync(d->pt_backdrop.p, 0, (size_t)cap_pt * 4, st));
    }
    d->long_edges.ensure(((size_t)cv[VKC_EDGES] + 1) * 4, st);
    if (d->failed) return;  // an allocation failed: nothing that would use the buffer is launched
    uint32_t *long_n = (uint32_t *)(totals + 8);  // (zeroed with the other totals when the flush starts)
    vkb_launch_bin_count(edges, edraw, cv[VKC_EDGES], C, d->draw_rect.as<int32_t>(), d->draw_ptbase.as<uint32_t>(), d->pt_count.as<uint32_t>(),
                         d->pt_backdrop.as<int32_t>(), d->long_edges.as<uint32_t>(), long_n, st);
    vkb_launch_backdrop_prefix(d->draw_rect.as<int32_t>(), d->draw_ptbase.as<uint32_t>(), d->draw_rowbase.as<uint32_t>(), d->row_owner.as<uint32_t>(),
                               cv[VKC_ROWS], C, d->pt_backdrop.as<int32_t>(), st);
    vkb_launch_pt_flags(d->pt_count.as<uint32_t>(), d->pt_backdrop.as<int32_t>(), cap_pt, C, draws, d->pt_owner.as<uint32_t>(), clip_draws,
                        d->pt_flags.as<uint32_t>(), st);
    // pt_slot doubles as the exclusive scan of the flags until the sorted slots overwrite it
    d->sorted_cnt.ensure((size_t)(cap_pt + 1) * 4, st);
    if (d->failed) return;  // an allocation failed: nothing that would use the buffer is launched
    uint32_t *flag_scan = d->sorted_cnt.as<uint32_t>();
    vkb_exclusive_scan<uint32_t, uint32_t>(d->pt_flags.as<uint32_t>(), flag_scan, 0, (uint32_t *)(totals + 5), d->scan, st, C, VKC_PT, cap_pt, VKC_NE);

    d->keys.ensure((size_t)(cap_ne + 1) * 4, st);
    d->vals.ensure((size_t)(cap_ne + 1) * 4, st);
    d->pt_draw.ensure((size_t)(cap_ne + 1) * 4, st);
    d->cursor.ensure((size_t)(cap_ne + 1) * 4, st);
    d->hdr.ensure((size_t)(cap_ne + 1) * 32, st);
    d->tile_first.ensure((size_t)n_tiles * 4 + 16, st);
    d->tile_end.ensure((size_t)n_tiles * 4 + 16, st);
    if (d->failed) return;  // an allocation failed: nothing that would use the buffer is launched
    vkb_launch_pt_compact(d->pt_flags.as<uint32_t>(), flag_scan, cap_pt, C, d->draw_rect.as<int32_t>(), d->draw_ptbase.as<uint32_t>(), d->pt_owner.as<uint32_t>(), sd,
                          d->keys.as<uint32_t>(), d->vals.as<uint32_t>(), d->pt_draw.as<uint32_t>(), st);
    int bits = 1;
    while ((1u << bits) < n_tiles && bits < 32) bits++;
    vkb_radix_sort(d->keys.as<uint32_t>(), d->vals.as<uint32_t>(), cap_ne, bits, d->sort, st, C, VKC_NE);
    // edge offsets in sorted (tile-major, draw-ordered) order
    uint32_t *eoff = d->cursor.as<uint32_t>();  // scan output; cursor proper is a separate zeroed array below
    DevBuf   &cur2 = d->cursor2;
    cur2.ensure((size_t)(cap_ne + 1) * 4, st);
    if (d->failed) return;  // an allocation failed: nothing that would use the buffer is launched
    uint32_t *tile_ms_words = nullptr;   // the per-tile multisample flags, when this frame must start from none set
    if (samples) {  // (analytic mode, samples == 0, keeps one colour per pixel: no per-sample plane)
        const bool fresh = surf->tile_ms.p == nullptr;
        surf->ms_image.ensure((size_t)n_tiles * 256 * samples * 4, st);
        surf->tile_ms.ensure((size_t)n_tiles + 16, st);
        surf->ms_mask.ensure((size_t)n_tiles * 32 + 16, st);
        if (d->failed) return;  // an allocation failed: nothing that would use the buffer is launched
        if (fresh || surf->known_clear) tile_ms_words = surf->tile_ms.as<uint32_t>();
    }
    {
        // sorted_cnt currently holds flag_scan which headers_k still needs: put the sorted counts in pt_flags instead
        uint32_t *scnt = d->pt_flags.as<uint32_t>();
        // (the same launch zeroes the scatter cursors, the tile list bounds and those flags)
        vkb_launch_sorted_counts(d->vals.as<uint32_t>(), cap_ne, C, d->pt_count.as<uint32_t>(), scnt, d->pt_slot.as<uint32_t>(), cur2.as<uint32_t>(), d->tile_first.as<uint32_t>(),
                                 d->tile_end.as<uint32_t>(), n_tiles, tile_ms_words, st);
        vkb_exclusive_scan<uint32_t, uint32_t>(scnt, eoff, 0, (uint32_t *)(totals + 6), d->scan, st, C, VKC_NE, cap_ne, VKC_TE);
    }
    d->tile_edges.ensure((size_t)(cv[VKC_TE] + 1) * 16, st);
    if (d->failed) return;  // an allocation failed: nothing that would use the buffer is launched
    vkb_launch_headers(d->keys.as<uint32_t>(), d->vals.as<uint32_t>(), cap_ne, C, d->pt_draw.as<uint32_t>(), flag_scan, d->pt_backdrop.as<int32_t>(),
                       d->pt_count.as<uint32_t>(), eoff, d->paints.as<vkb_paint>(), d->hdr.as<int4>(), d->tile_first.as<uint32_t>(), d->tile_end.as<uint32_t>(), st);
    vkb_launch_bin_scatter(edges, edraw, cv[VKC_EDGES], C, d->draw_rect.as<int32_t>(), d->draw_ptbase.as<uint32_t>(), d->pt_slot.as<uint32_t>(), eoff,
                           cur2.as<uint32_t>(), d->tile_edges.as<vkb_edge>(), d->long_edges.as<uint32_t>(), long_n, st);

    // ---- 6. fine pass ----
    FineArgs fa;
    fa.sd = sd;
    fa.counts = C;
    fa.tile_first = d->tile_first.as<uint32_t>(); fa.tile_end = d->tile_end.as<uint32_t>();
    fa.hdr = d->hdr.as<int4>(); fa.tile_edges = d->tile_edges.as<vkb_edge>();
    fa.paints = d->paints.as<vkb_paint>(); fa.grads = d->grads.as<vkb_gradient>();
    fa.gprep = d->gprep.as<float>();
    fa.tile_counter = (uint32_t *)(totals + 10);  // (zeroed with the other totals when the flush starts)
    fa.tile_lo = 0; fa.tile_hi = n_tiles;
    d->wscratch.ensure(vkb_fine_wscratch_words(samples) * 4, st);
    if (d->failed) return;  // an allocation failed: nothing that would use the buffer is launched
    fa.wscratch = d->wscratch.as<int32_t>();
    fa.surfpats = d->surfpats.as<vkb_surfpat>();
    fa.image = surf->image.as<uint32_t>();
    fa.ms_image = surf->ms_image.as<uint32_t>(); fa.tile_ms = surf->tile_ms.as<uint8_t>(); fa.ms_mask = surf->ms_mask.as<uint32_t>();
    fa.dst_is_clear = surf->known_clear ? 1 : 0;
    // clip / save bits: the stencil-aware kernel variant only runs when this batch writes them or earlier ones left some behind
    fa.stencil = nullptr; fa.stencil_in = 0;
    if (stencil_ops || surf->stencil_live) {
        if (surf->stencil_samples != samples) { surf->stencil_live = false; surf->stencil_samples = samples; }  // coverage mode changed: layout differs
        surf->stencil.ensure(stencil_bytes(surf, samples), st);
        if (d->failed) return;  // an allocation failed: nothing that would use the buffer is launched
        fa.stencil    = surf->stencil.as<uint32_t>();
        fa.stencil_in = surf->stencil_live ? 1 : 0;
        if (stencil_ops && d->stencil_after) surf->stencil_live = d->stencil_after == 1;
    }
    fa.winding_out = nullptr; fa.winding_draw = 0;
    if (cap && cap->winding) {
        size_t wb = (size_t)sd.width * sd.height * (samples ? samples : 1) * 4;  // analytic mode: one float area per pixel
        wbuf.ensure(wb, st);
        if (d->failed) return;  // an allocation failed: nothing that would use the buffer is launched
        VKB_CUDA_OK(cudaMemsetAsync(wbuf.p, 0, wb, st));
        fa.winding_out = wbuf.as<int32_t>(); fa.winding_draw = cap->winding_draw;
    }
    VKB_EVENT_RECORD(d, d->ev_stage[4]);
    if (d->capturing) d->graph_fine_events = cudaEventRecordWithFlags(d->ev_fine0, st, cudaEventRecordExternal) == cudaSuccess;
    else VKB_EVENT_RECORD(d, d->ev_fine0);
    // a surface with a read-back target: bands of tile rows, each copied to the host on the copy stream as soon as it is finished
    const bool     banded  = surf->readback && !(cap && cap->winding);
    const uint32_t n_bands = banded ? (sd.tiles_y >= 128 ? 4u : (sd.tiles_y >= 2 * VKB_MAX_BANDS ? 2u : 1u)) : 1u;  // (a small surface: one band, copied behind the frame)
    d->band_counters.ensure(VKB_MAX_BANDS * 4, st);
    if (banded) VKB_CUDA_OK(cudaMemsetAsync(d->band_counters.p, 0, VKB_MAX_BANDS * 4, st));
    for (uint32_t k = 0; k < n_bands; k++) {
        const uint32_t r0 = (uint32_t)((uint64_t)sd.tiles_y * k / n_bands), r1 = (uint32_t)((uint64_t)sd.tiles_y * (k + 1) / n_bands);
        fa.tile_lo = r0 * sd.tiles_x; fa.tile_hi = r1 * sd.tiles_x;
        if (banded) fa.tile_counter = d->band_counters.as<uint32_t>() + k;
        vkb_launch_fine(fa, st);
        if (banded) {
            const size_t y0 = (size_t)r0 * VKB_TILE, y1 = r1 * VKB_TILE < sd.height ? (size_t)r1 * VKB_TILE : sd.height;
            VKB_CUDA_OK(cudaEventRecord(d->ev_band[k], st));
            VKB_CUDA_OK(cudaStreamWaitEvent(d->copy_stream, d->ev_band[k], 0));
            // (cudaMemcpyDefault: the target is host memory for a read-back, this or a PEER device's memory for a stripe delivered to its root)
            VKB_CUDA_OK(cudaMemcpyAsync(surf->readback + y0 * sd.width * 4, surf->image.as<uint8_t>() + y0 * sd.width * 4, (y1 - y0) * sd.width * 4, cudaMemcpyDefault,
                                        d->copy_stream));
        }
    }
    if (banded) {  // the stream that renders goes on only when the image is out (the next flush may clear it) - and the capture, if any, is joined
        VKB_CUDA_OK(cudaEventRecord(d->ev_copied, d->copy_stream));
        VKB_CUDA_OK(cudaStreamWaitEvent(st, d->ev_copied, 0));
    }
    surf->readback_valid = banded;
    if (d->capturing) d->graph_fine_events = d->graph_fine_events && cudaEventRecordWithFlags(d->ev_fine1, st, cudaEventRecordExternal) == cudaSuccess;
    else VKB_EVENT_RECORD(d, d->ev_fine1);
    VKB_EVENT_RECORD(d, d->ev_stage[5]);
    surf->known_clear = false;
}

// the capacities of this attempt travel as kernel arguments (no staging copy to wait on); every count starts at zero
// The same launch zeroes the 16 raw totals of the flush and empties the per-draw boxes (bbox: n_draws x 4, or null) - one graph node
// instead of a kernel, a memset and another kernel at the head of every frame.
struct CapArgs { uint32_t v[16]; };
__global__ void counts_reset_k(vkb_counts *C, CapArgs caps, uint64_t *totals, int32_t *bbox, uint32_t n_draws) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < 16) { C->n[i] = 0; C->need[i] = 0; C->cap[i] = caps.v[i]; totals[i] = 0; }
    if (i == 0) C->overflow = 0;
    if (bbox && i < n_draws) { bbox[4 * i] = INT32_MAX; bbox[4 * i + 1] = INT32_MAX; bbox[4 * i + 2] = INT32_MIN; bbox[4 * i + 3] = INT32_MIN; }
}
static void enqueue_counts_reset(vkb_device_impl *d, int32_t *bbox, uint32_t n_draws) {
    d->counts.ensure(sizeof(vkb_counts), d->stream);
    d->totals.ensure(16 * 8, d->stream);
    if (d->failed) return;
    CapArgs a;
    memcpy(a.v, d->capv, sizeof a.v);
    counts_reset_k<<<bbox && n_draws > 256 ? vkb_div_up(n_draws, 256) : 1, 256, 0, d->stream>>>(d->counts.as<vkb_counts>(), a, d->totals.as<uint64_t>(), bbox, n_draws);
    VKB_LAUNCHED();
}

// one attempt at the whole pipeline for the resident batch; no host round trip
static void enqueue_flush(vkb_device_impl *d, vkb_surface_impl *surf, SurfaceDesc sd, vkb_capture *cap, DevBuf &wbuf) {
    cudaStream_t st = d->stream;
    plan_caps(d, sd);
    const uint32_t *cv = d->capv;
    d->draw_bbox.ensure((size_t)(d->n_draws + 1) * 16, st);
    if (d->failed) return;  // an allocation failed: nothing that would use the buffer is launched
    enqueue_counts_reset(d, d->draw_bbox.as<int32_t>(), d->n_draws);   // (+ the totals zeroed, the draw boxes emptied)
    if (d->failed) return;
    vkb_counts *C = d->counts.as<vkb_counts>();
    VKB_EVENT_RECORD(d, d->ev_stage[0]);
    uint64_t *totals = d->totals.as<uint64_t>();
    // ---- 1. flatten: count -> scan -> emit ----
    d->elem_cnt.ensure((size_t)(d->n_elems + 1) * 4, st);
    d->sp_first.ensure((size_t)(d->n_sp + 1) * 4, st);
    d->sp_count.ensure((size_t)(d->n_sp + 1) * 4, st);
    d->pts.ensure((size_t)(cv[VKC_POINTS] + 1) * 8, st);
    d->ptflags.ensure((size_t)cv[VKC_POINTS] + 16, st);
    d->fjob_base.ensure((size_t)(d->n_fjobs + 1) * 4, st);
    d->sjob_base.ensure((size_t)(d->n_sjobs + 1) * 4, st);
    if (d->failed) return;  // an allocation failed: nothing that would use the buffer is launched
    // curve-heavy batches keep the first 16 points of every element from the counting pass (16 * 8 B per element); batches of few elements
    // (a tiger: 2500) the first 64 - there the emitting pass lasts as long as its longest curve, and with 64 slots it only copies
    float2        *fcache = nullptr;
    const uint32_t fcache_n = d->n_elems <= 65536 ? 64u : 16u;
    if (d->n_curves && (uint64_t)d->n_curves * 8 >= d->n_elems) {
        d->flat_cache.ensure((size_t)d->n_elems * fcache_n * 8, st);
        if (d->failed) return;  // an allocation failed: nothing that would use the buffer is launched
        fcache = d->flat_cache.as<float2>();
    }
    if (d->n_elems) {
        vkb_launch_flatten_count(d->elem_hdr.as<uint32_t>(), d->elem_data.as<float>(), d->n_elems, d->elem_cnt.as<uint32_t>(), fcache, fcache_n, st);
        vkb_exclusive_scan<uint32_t, uint32_t>(d->elem_cnt.as<uint32_t>(), d->elem_cnt.as<uint32_t>(), d->n_elems, (uint32_t *)totals, d->scan, st);
        vkb_launch_subpath_ranges(d->subpaths.as<vkb_subpath>(), d->n_sp, d->elem_cnt.as<uint32_t>(), d->n_elems, (uint32_t *)totals,
                                  d->sp_first.as<uint32_t>(), d->sp_count.as<uint32_t>(), st);
    }
    // ---- 2. job sizes: fill jobs need > 2 points, stroke jobs >= 2 (point counts come from the count scan alone) ----
    d->sp_bbox.ensure((size_t)(d->n_sp + 1) * 16, st);
    d->long_sp.ensure((size_t)(d->n_sp + 1) * 4, st);
    if (d->failed) return;
    // (geometry captures - vkvg_b200_stroke_geometry, path_edges - report the whole tessellation, on or off the surface)
    const int4 *sp_bbox = (cap && cap->geometry_only) ? nullptr : d->sp_bbox.as<int4>();
    if (sp_bbox) vkb_launch_sp_bounds(d->subpaths.as<vkb_subpath>(), d->n_sp, d->elem_hdr.as<uint32_t>(), d->elem_data.as<float>(), d->n_elems, d->long_sp.as<uint32_t>(), (uint32_t *)(totals + 12), d->scan, d->sp_bbox.as<int4>(), d->max_sp_elems > VKB_SP_LONG, st);
    if (d->n_fjobs && d->n_sjobs && d->n_fjobs <= VKB_SCAN_SMALL && d->n_sjobs <= VKB_SCAN_SMALL) {   // a small batch that fills and strokes: two launches instead of four
        vkb_launch_job_counts2(d->fjob_sp.as<uint32_t>(), d->fjob_draw.as<uint32_t>(), d->n_fjobs, d->fjob_base.as<uint32_t>(), d->sjob_sp.as<uint32_t>(), d->sjob_draw.as<uint32_t>(),
                               d->n_sjobs, d->sjob_base.as<uint32_t>(), d->sp_count.as<uint32_t>(), sp_bbox, d->draws.as<vkb_draw>(), d->xforms.as<vkb_xform>(),
                               d->strokes.as<vkb_stroke>(), sd, st);
        scan_small2_k<uint32_t><<<2, 1024, 0, st>>>(d->fjob_base.as<uint32_t>(), d->n_fjobs, (uint32_t *)(totals + 1), d->sjob_base.as<uint32_t>(), d->n_sjobs, (uint32_t *)(totals + 2));
        VKB_LAUNCHED();
    } else {
    if (d->n_fjobs) {
        vkb_launch_job_counts(d->fjob_sp.as<uint32_t>(), d->fjob_draw.as<uint32_t>(), d->n_fjobs, d->sp_count.as<uint32_t>(), 3, sp_bbox, d->draws.as<vkb_draw>(),
                              d->xforms.as<vkb_xform>(), d->strokes.as<vkb_stroke>(), sd, d->fjob_base.as<uint32_t>(), st);
        vkb_exclusive_scan<uint32_t, uint32_t>(d->fjob_base.as<uint32_t>(), d->fjob_base.as<uint32_t>(), d->n_fjobs, (uint32_t *)(totals + 1), d->scan, st);
    }
    if (d->n_sjobs) {
        vkb_launch_job_counts(d->sjob_sp.as<uint32_t>(), d->sjob_draw.as<uint32_t>(), d->n_sjobs, d->sp_count.as<uint32_t>(), 2, sp_bbox, d->draws.as<vkb_draw>(),
                              d->xforms.as<vkb_xform>(), d->strokes.as<vkb_stroke>(), sd, d->sjob_base.as<uint32_t>(), st);
        vkb_exclusive_scan<uint32_t, uint32_t>(d->sjob_base.as<uint32_t>(), d->sjob_base.as<uint32_t>(), d->n_sjobs, (uint32_t *)(totals + 2), d->scan, st);
    }
    }
    const bool nz_split = d->nz_any && d->n_fjobs;
    commit_flatten_k<<<1, 1, 0, st>>>(C, totals, d->n_fjobs ? 1u : 0u, d->n_sjobs ? 1u : 0u, nz_split ? 1u : 0u);
    VKB_LAUNCHED();
    if (d->n_elems)
        vkb_launch_flatten_emit(d->elem_hdr.as<uint32_t>(), d->elem_data.as<float>(), d->n_elems, d->elem_cnt.as<uint32_t>(), d->pts.as<float2>(),
                                d->ptflags.as<uint8_t>(), C, fcache, fcache_n, st);
    if (nz_split) {  // NON_ZERO fills / clips as libtess makes them: classify the draws, then every fill edge in its pieces
        d->nz_mode.ensure((size_t)d->n_draws + 16, st);
        d->edges.ensure(((size_t)cv[VKC_EDGES] + 1) * 16, st);
        d->edge_draw.ensure(((size_t)cv[VKC_EDGES] + 1) * 4, st);
        if (d->failed) return;
        vkb_launch_nz_classify(d->draws.as<vkb_draw>(), d->n_draws, d->sp_first.as<uint32_t>(), d->sp_count.as<uint32_t>(), d->pts.as<float2>(), C,
                               d->paints.as<vkb_paint>(), d->nz_mode.as<uint8_t>(), st);
        vkb_launch_nz_split(d->pts.as<float2>(), d->draws.as<vkb_draw>(), d->xforms.as<vkb_xform>(), d->fjob_draw.as<uint32_t>(), d->fjob_sp.as<uint32_t>(),
                            d->fjob_base.as<uint32_t>(), d->n_fjobs, d->sp_first.as<uint32_t>(), d->sp_count.as<uint32_t>(), d->fcnt.as<uint32_t>(), d->n_draws,
                            d->nz_mode.as<uint8_t>(), cv[VKC_FILL], C, sd, d->edges.as<vkb_edge>(), d->edge_draw.as<uint32_t>(), (uint32_t *)(totals + 11), cv[VKC_FEDGES],
                            d->draw_bbox.as<int32_t>(), st);
    }
    VKB_EVENT_RECORD(d, d->ev_stage[1]);

    // ---- 3. strokes: (dash phase scan) -> count -> scan -> emit ----
    const uint32_t cap_items = d->n_sjobs ? cv[VKC_SITEMS] : 0;
    const bool     legacy_emit = vkb_stroke_emit_mode() == 1;
    if (cap_items) {
        StrokeArgs sa = {d->pts.as<float2>(), d->ptflags.as<uint8_t>(), d->draws.as<vkb_draw>(), d->strokes.as<vkb_stroke>(), d->dashes.as<float>(), d->sjob_draw.as<uint32_t>(),
                         d->sjob_sp.as<uint32_t>(), d->sjob_base.as<uint32_t>(), d->n_sjobs, d->sp_first.as<uint32_t>(), d->sp_count.as<uint32_t>(),
                         d->subpaths.as<vkb_subpath>(), nullptr, cap_items, C};
        if (d->any_dash) {
            d->seglen.ensure((size_t)(cap_items + 1) * 4, st);
            d->cum.ensure((size_t)(cap_items + 1) * 8, st);
            if (d->failed) return;  // an allocation failed: nothing that would use the buffer is launched
            vkb_launch_stroke_seglen(sa, d->seglen.as<float>(), st);
            vkb_exclusive_scan<float, double>(d->seglen.as<float>(), d->cum.as<double>(), 1, nullptr, d->scan, st, C, VKC_SITEMS, (uint64_t)cap_items + 1);
            sa.cum = d->cum.as<double>();
        }
        d->item_counts.ensure((size_t)(cap_items + 1) * 8, st);
        d->job_inverse.ensure((size_t)(d->n_sjobs + 1) * 4, st);
        if (d->failed) return;  // an allocation failed: nothing that would use the buffer is launched
        // (job_inverse needs no clearing: stroke_patch_closed_k reads it for closed undashed jobs of at least two points, and the last item of
        // exactly those writes it)
        unsigned long long *ic = d->item_counts.as<unsigned long long>();
        vkb_launch_stroke_count(sa, ic, st);
        vkb_exclusive_scan<unsigned long long, unsigned long long>(ic, ic, 0, (unsigned long long *)(totals + 3), d->scan, st, C, VKC_SITEMS, cap_items);
        d->sdraw_first_item.ensure((size_t)d->n_sdraws * 4 + 16, st);
        if (d->failed) return;  // an allocation failed: nothing that would use the buffer is launched
        commit_stroke_k<<<vkb_div_up(d->n_sdraws ? d->n_sdraws : 1, 256), 256, 0, st>>>(C, totals, 1u, d->n_extra, d->sdraw_first_job.as<uint32_t>(), d->sjob_base.as<uint32_t>(), d->n_sdraws,
                                                                                      d->sdraw_first_item.as<uint32_t>());
        VKB_LAUNCHED();
        const bool want_verts = cap && cap->verts;   // float vertices: geometry captures only (the rasteriser reads the snapped ones)
        if (legacy_emit || want_verts) d->verts.ensure((size_t)(cv[VKC_VERTS] + 1) * 8, st);
        d->inds.ensure((size_t)(cv[VKC_INDS] + 4) * 4, st);
        d->snapped.ensure((size_t)(cv[VKC_VERTS] + 2) * 8, st);
        if (d->failed) return;  // an allocation failed: nothing that would use the buffer is launched
        if (legacy_emit) vkb_launch_stroke_emit(sa, ic, d->verts.as<float2>(), d->inds.as<uint32_t>(), d->job_inverse.as<uint32_t>(), st);
        else vkb_launch_stroke_emit_snapped(sa, ic, d->xforms.as<vkb_xform>(), sd, want_verts ? d->verts.as<float2>() : nullptr, d->snapped.as<int2>(), d->inds.as<uint32_t>(),
                                            d->job_inverse.as<uint32_t>(), st);
    } else {
        commit_stroke_k<<<1, 32, 0, st>>>(C, totals, 0u, d->n_extra, nullptr, nullptr, 0u, nullptr);
        VKB_LAUNCHED();
    }

    // ---- 4. edges ----
    VKB_EVENT_RECORD(d, d->ev_stage[2]);
    d->edges.ensure(((size_t)cv[VKC_EDGES] + 1) * 16, st);
    d->edge_draw.ensure(((size_t)cv[VKC_EDGES] + 1) * 4, st);
    if (d->failed) return;  // an allocation failed: nothing that would use the buffer is launched
    vkb_edge *edges = d->edges.as<vkb_edge>();
    uint32_t *edraw = d->edge_draw.as<uint32_t>();
    if (!nz_split)  // (else the fill edges are in place since the end of the flatten stage)
        vkb_launch_fill_edges(d->pts.as<float2>(), d->draws.as<vkb_draw>(), d->xforms.as<vkb_xform>(), d->fjob_draw.as<uint32_t>(), d->fjob_sp.as<uint32_t>(), d->fjob_base.as<uint32_t>(),
                              d->n_fjobs, d->sp_first.as<uint32_t>(), d->sp_count.as<uint32_t>(), d->n_fjobs ? cv[VKC_FILL] : 0, C, sd, edges, edraw, d->draw_bbox.as<int32_t>(), st);
    const bool      stops_here = (cap && cap->geometry_only) || d->n_draws == 0;
    const uint32_t *live_edges = nullptr;   // the stored stroke edges, when the binning stage is to commit their number
    if (cap_items && d->n_sdraws) {   // (sdraw_first_item: written by commit_stroke_k)
        vkb_launch_tri_edges(legacy_emit ? d->verts.as<float2>() : nullptr, cv[VKC_VERTS], d->snapped.as<int2>(), d->inds.as<uint32_t>(), cv[VKC_TRIS], C, d->draws.as<vkb_draw>(), d->xforms.as<vkb_xform>(), d->sdraw_id.as<uint32_t>(),
                             d->sdraw_first_item.as<uint32_t>(), d->n_sdraws, d->item_counts.as<unsigned long long>(), sd, edges, edraw, d->n_extra, (uint32_t *)(totals + 9), C, d->draw_bbox.as<int32_t>(), stops_here, st);
        if (!stops_here) live_edges = (const uint32_t *)(totals + 9);
    }
    if (d->n_extra) {
        extra_rect_edges_k<<<vkb_div_up(d->n_extra / 4, 64), 64, 0, st>>>(edges, edraw, d->extra_edge_draw.as<uint32_t>(), d->n_extra / 4, (int32_t)sd.width,
                                                                         (int32_t)sd.height, C, d->draw_bbox.as<int32_t>());
        VKB_LAUNCHED();
    }
    if (stops_here) return;
    enqueue_bin_and_fine(d, surf, sd, d->n_draws, cap, d->draws.as<vkb_draw>(), wbuf, live_edges);
}

// Everything that shapes the launches of a flush besides the data in the buffers.  Two flushes with equal keys issue the
// same kernels with the same arguments, so the second one can be a replay of a CUDA graph captured from the first: a frame
// loop that redraws a scene of the same structure pays one graph launch per frame instead of ~40 kernel launches.
struct FlushKey {
    uint32_t n_elems, n_sp, n_draws, n_fjobs, n_sjobs, n_sdraws, n_extra, n_curves, n_grads;
    uint32_t flags;  // any_dash | has_clip_draws << 1 | has_stencil_ops << 2 | stencil_after << 3
    uint32_t capv[16];
    const void *surf;
    uint32_t w, h, samples, full_h, origin_y;
    uint32_t known_clear, stencil_live, stencil_samples, tile_ms_allocated;
    uint32_t fine_mode, pad0;  // vkb_fine_set_mode: which fine kernel a captured graph holds
    const void *readback;      // the host pointer the band copies of a captured graph write to
    unsigned long long alloc_generation;
};
static_assert(sizeof(FlushKey) <= 256, "FlushKey");
static void enqueue_flush_maybe_graph(vkb_device_impl *d, vkb_surface_impl *surf, SurfaceDesc sd, vkb_capture *cap, DevBuf &wbuf, bool stage_timing) {
    cudaStream_t st = d->stream;
    if (!d->begin_recorded) VKB_CUDA_OK(cudaEventRecord(d->ev_begin, st));
    d->begin_recorded = false;
    if (cap || stage_timing || !d->graphs_enabled) {  // captures download intermediates, stage timing needs events between the kernels
        enqueue_flush(d, surf, sd, cap, wbuf);
        d->have_last_key = false;
        return;
    }
    plan_caps(d, sd);
    uint8_t  kbuf[256];
    memset(kbuf, 0, sizeof kbuf);
    FlushKey k;
    memset(&k, 0, sizeof k);
    k.n_curves = d->n_curves; k.n_grads = d->n_grads;
    k.n_elems = d->n_elems; k.n_sp = d->n_sp; k.n_draws = d->n_draws; k.n_fjobs = d->n_fjobs; k.n_sjobs = d->n_sjobs; k.n_sdraws = d->n_sdraws; k.n_extra = d->n_extra;
    k.flags = (d->any_dash ? 1u : 0u) | (d->has_clip_draws ? 2u : 0u) | (d->has_stencil_ops ? 4u : 0u) | ((uint32_t)d->stencil_after << 3) | (d->nz_any ? 32u : 0u) | (d->max_sp_elems > VKB_SP_LONG ? 64u : 0u);
    memcpy(k.capv, d->capv, sizeof k.capv);
    k.surf = surf; k.w = sd.width; k.h = sd.height; k.samples = sd.samples; k.full_h = sd.full_height; k.origin_y = sd.origin_y;
    k.known_clear = surf->known_clear; k.stencil_live = surf->stencil_live; k.stencil_samples = surf->stencil_samples;
    k.tile_ms_allocated = surf->tile_ms.p != nullptr;
    k.alloc_generation = g_vkb_alloc_generation;
    k.fine_mode = (uint32_t)vkb_fine_get_mode();
    k.readback = surf->readback;

    memcpy(kbuf, &k, sizeof k);
    if (d->graph_exec && !memcmp(kbuf, d->graph_key, sizeof kbuf)) {
        // replay; the host-side effects of enqueue_flush on the surface flags are re-applied by hand
        VKB_CUDA_OK(cudaGraphLaunch(d->graph_exec, st));
        g_vkb_launches += d->graph_launches;
        d->n_graph_replays++;
        surf_restore(surf, d->graph_after);
        return;
    }
    if (d->have_last_key && !memcmp(kbuf, d->last_key, sizeof kbuf)) {
        // second flush of this shape: every buffer already has its size, so nothing allocates while the stream is capturing
        if (d->graph_exec) { cudaGraphExecDestroy(d->graph_exec); d->graph_exec = nullptr; }
        const unsigned long long l0 = g_vkb_launches, gen0 = g_vkb_alloc_generation;
        const SurfFlags          f0 = surf_flags(surf);
        cudaGraph_t g = nullptr;
        if (cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal) == cudaSuccess) {
            d->capturing = true;
            enqueue_flush(d, surf, sd, nullptr, wbuf);
            d->capturing = false;
            cudaError_t e = cudaStreamEndCapture(st, &g);
            if (e == cudaSuccess && g && gen0 == g_vkb_alloc_generation && cudaGraphInstantiate(&d->graph_exec, g, 0) == cudaSuccess) {
                d->graph_launches = g_vkb_launches - l0;
                g_vkb_launches    = l0;
                d->graph_after    = surf_flags(surf);
                memcpy(d->graph_key, kbuf, sizeof kbuf);
                cudaGraphDestroy(g);
                VKB_CUDA_OK(cudaGraphLaunch(d->graph_exec, st));
                g_vkb_launches += d->graph_launches;
                memcpy(d->last_key, kbuf, sizeof kbuf);
                return;
            }
            if (g) cudaGraphDestroy(g);
            cudaGetLastError();
            d->graph_exec = nullptr;
            g_vkb_launches = l0;
            surf_restore(surf, f0);
        } else cudaGetLastError();
    }
    enqueue_flush(d, surf, sd, nullptr, wbuf);
    memcpy(d->last_key, kbuf, sizeof kbuf);
    d->have_last_key = true;
}

static void fill_stats(vkb_device_impl *d, const vkb_counts &h, vkb_stats &S, bool with_stages) {
    S.n_elems = d->n_elems; S.h2d_bytes = d->h2d_bytes; S.ms_host_upload = d->ms_host_upload;
    S.n_points = h.n[VKC_POINTS]; S.n_fill_edges = h.n[VKC_FEDGES]; S.n_stroke_items = h.n[VKC_SITEMS];
    S.n_verts = h.n[VKC_VERTS]; S.n_inds = h.n[VKC_INDS]; S.n_edges = h.n[VKC_EDGES];
    S.n_path_tiles = h.n[VKC_PT]; S.n_nonempty = h.n[VKC_NE]; S.n_tile_edges = h.n[VKC_TE];
    cudaEventElapsedTime(&S.ms_total, d->ev_begin, d->ev_end);
    if (!with_stages && d->graph_exec && d->graph_fine_events) {  // graph replay: only the fine kernel is bracketed (external event nodes)
        if (cudaEventElapsedTime(&S.ms_fine, d->ev_fine0, d->ev_fine1) != cudaSuccess) { cudaGetLastError(); S.ms_fine = 0.f; }
    }
    if (with_stages) {
        cudaEventElapsedTime(&S.ms_fine, d->ev_fine0, d->ev_fine1);
        for (int i = 0; i < VKB_N_STAGES; i++) cudaEventElapsedTime(&S.ms_stage[i], d->ev_stage[i], d->ev_stage[i + 1]);
    }
}

// Runs the resident batch onto surf.  allow_async: return right after the work is queued; the overflow check (and the
// replay it may ask for) then happens in finish_pending, which every entry point that touches the device calls first.
static int run_flush(vkb_device_impl *d, vkb_surface_impl *surf, uint32_t samples, vkb_capture *cap, vkb_stats *stats, bool allow_async) {
    cudaStream_t st = d->stream;
    SurfaceDesc  sd = {surf->w, surf->h, samples, (surf->w + VKB_TILE - 1) / VKB_TILE, (surf->h + VKB_TILE - 1) / VKB_TILE, surf->full_h, surf->origin_y,
                       surf->band_h / VKB_TILE};
    const SurfFlags before = surf_flags(surf);
    const bool      geometry_only = (cap && cap->geometry_only) || d->n_draws == 0;
    for (int attempt = 0; attempt < 16; attempt++) {
        DevBuf wbuf;
        surf_restore(surf, before);
        enqueue_flush_maybe_graph(d, surf, sd, cap, wbuf, stats != nullptr && d->stage_timing);
        VKB_CUDA_OK(cudaEventRecord(d->ev_end, st));
        VKB_CUDA_OK(cudaMemcpyAsync(d->counts_host, d->counts.p, sizeof(vkb_counts), cudaMemcpyDeviceToHost, st));
        if (allow_async && !cap && !stats) {
            d->pending = true; d->pending_surf = surf; d->pending_samples = samples; d->pending_before = before;
            return g_cuda_failed;
        }
        VKB_CUDA_OK(cudaStreamSynchronize(st));
        const vkb_counts &h = *d->counts_host;
        if (h.overflow) {
            grow_caps_from_need(d, h);
            wbuf.release();
            if (g_cuda_failed) return 1;
            continue;
        }
        if (cap) {
            download(d, cap->points, d->pts.p, (size_t)h.n[VKC_POINTS] * 2);
            download(d, cap->ptflags, d->ptflags.p, (size_t)h.n[VKC_POINTS]);
            download(d, cap->sp_first, d->sp_first.p, d->n_sp);
            download(d, cap->sp_count, d->sp_count.p, d->n_sp);
            download(d, cap->verts, d->verts.p, (size_t)h.n[VKC_VERTS] * 2);
            download(d, cap->inds, d->inds.p, (size_t)h.n[VKC_INDS]);
            download(d, cap->edges, d->edges.p, (size_t)h.n[VKC_EDGES] * 4);
            download(d, cap->edge_draw, d->edge_draw.p, (size_t)h.n[VKC_EDGES]);
            if (cap->winding && !geometry_only) {
                VKB_CUDA_OK(cudaMemcpyAsync(cap->winding, wbuf.p, (size_t)sd.width * sd.height * (samples ? samples : 1) * 4, cudaMemcpyDeviceToHost, st));
                VKB_CUDA_OK(cudaStreamSynchronize(st));
            }
        }
        wbuf.release();
        if (stats) {
            vkb_stats S;
            memset(&S, 0, sizeof S);
            fill_stats(d, h, S, !geometry_only && d->stage_timing);
            *stats = S;
        }
        return g_cuda_failed;
    }
    fprintf(stderr, "vkvg_b200: intermediate buffers still overflow after 16 attempts\n");
    return 1;
}
static int finish_pending(vkb_device_impl *d) {
    if (!d->pending) return g_cuda_failed;
    d->pending = false;
    VKB_CUDA_OK(cudaStreamSynchronize(d->stream));
    if (d->counts_host->overflow) {  // rare: first flush of a new kind of scene.  Nothing was written to the surface; replay with room.
        grow_caps_from_need(d, *d->counts_host);
        surf_restore(d->pending_surf, d->pending_before);
        return run_flush(d, d->pending_surf, d->pending_samples, nullptr, nullptr, false);
    }
    return g_cuda_failed;
}

int vkb_render_resident(vkb_device_impl *d, vkb_surface_impl *surf, uint32_t samples, vkb_capture *cap, vkb_stats *stats) {
    dev_enter(d);
    finish_pending(d);
    return run_flush(d, surf, samples, cap, stats, false);
}

int vkb_render(vkb_device_impl *d, vkb_surface_impl *s, uint32_t samples, const vkb_batch &b, vkb_capture *cap, vkb_stats *stats) {
    if (vkb_upload(d, b)) return 1;
    const int r = run_flush(d, s, samples, cap, stats, true);
    // the element arrays were copied straight out of the recorder's pinned vectors, which the caller goes on writing into as soon
    // as the flush returns: wait for that copy (not for the kernels queued behind it)
    VKB_CUDA_OK(cudaEventSynchronize(d->ev_h2d));
    return r | g_cuda_failed;
}


// The packed command stream decoded on the device and rendered: 0 = queued, 1 = device error, 2 = the stream is outside what the device
// decodes (nothing was touched: the caller decodes it on the host).  cmds / args are host memory (pinned for full speed) and are free
// again when this returns.
int vkb_submit_stream(vkb_device_impl *d, vkb_surface_impl *surf, uint32_t samples, const uint32_t *cmds, uint64_t n_cmds, const float *args, uint64_t n_args,
                      const vkb_decode_init &init, vkb_decode_census *census_out, vkb_stats *stats) {
    dev_enter(d);
    cudaStream_t st = d->stream;
    if (n_cmds == 0 || n_cmds > 0x7fffff00ull || n_args > 0xfffffff0ull) return 2;
    finish_pending(d);
    VKB_CUDA_OK(cudaStreamSynchronize(st));  // the previous flush may still be reading the batch these kernels overwrite
    const uint32_t nc = (uint32_t)n_cmds, na = (uint32_t)n_args, NFIELDS = vkd_n_fields();
    d->dc_cmds.ensure((size_t)nc * 4 + 16, st);
    d->dc_args.ensure((size_t)na * 4 + 16, st);
    d->dc_S.ensure(vkd_scan_words(nc) * 4, st);
    d->dc_blocksum.ensure(vkd_blocksum_words(nc) * 4 + 16, st);
    d->dc_small.ensure(1024, st);
    if (!d->dc_host) VKB_CUDA_OK(cudaHostAlloc((void **)&d->dc_host, 1024, cudaHostAllocDefault));
    if (d->failed || !d->dc_host) return 1;
    uint32_t          *irregular = d->dc_small.as<uint32_t>();
    vkb_decode_census *census    = (vkb_decode_census *)(d->dc_small.as<uint8_t>() + 256);
    VKB_CUDA_OK(cudaMemsetAsync(d->dc_small.p, 0, 1024, st));
    VKB_CUDA_OK(cudaMemcpyAsync(d->dc_cmds.p, cmds, (size_t)nc * 4, cudaMemcpyHostToDevice, st));
    if (na) VKB_CUDA_OK(cudaMemcpyAsync(d->dc_args.p, args, (size_t)na * 4, cudaMemcpyHostToDevice, st));
    // ---- phase A: where everything goes ----
    vkb_launch_decode_scan(d->dc_cmds.as<uint32_t>(), nc, d->dc_S.as<uint32_t>(), d->dc_blocksum.as<uint32_t>(), irregular, st);
    VKB_CUDA_OK(cudaMemcpyAsync(d->dc_host, d->dc_S.as<uint32_t>() + (size_t)nc * NFIELDS, NFIELDS * 4, cudaMemcpyDeviceToHost, st));
    VKB_CUDA_OK(cudaMemcpyAsync(d->dc_host + 64, irregular, 4, cudaMemcpyDeviceToHost, st));
    VKB_CUDA_OK(cudaStreamSynchronize(st));
    if (d->failed) return 1;
    if (d->dc_host[64]) { if (getenv("VKVG_B200_DEBUG")) fprintf(stderr, "vkvg_b200_submit: host decoder (scan flags 0x%x)\n", d->dc_host[64]); return 2; }
    const vkd_totals t = vkd_read_totals(d->dc_host);
    if (t.n_draws == 0 || t.n_xforms > 65000u || t.n_strokes > 65000u) return 2;
    // ---- phase B: the batch, written where an upload would have put it ----
    d->elem_hdr.ensure((size_t)(t.n_elems + 1) * 4 + 16, st);
    d->elem_data.ensure((size_t)(t.n_data + 4) * 4 + 16, st);
    d->subpaths.ensure((size_t)(t.n_subpaths + 1) * sizeof(vkb_subpath), st);
    d->draws.ensure((size_t)(t.n_draws + 1) * sizeof(vkb_draw), st);
    d->xforms.ensure((size_t)(t.n_xforms + 1) * sizeof(vkb_xform), st);
    d->dc_xfscale.ensure((size_t)(t.n_xforms + 1) * 8, st);
    d->strokes.ensure((size_t)(t.n_strokes + 1) * sizeof(vkb_stroke), st);
    d->grads.ensure((size_t)(t.n_grads + 2) * sizeof(vkb_gradient), st);
    d->dashes.ensure((size_t)(t.n_dash_floats + 4) * 4, st);
    d->surfpats.ensure(sizeof(vkb_surfpat), st);
    d->dc_lists.ensure((size_t)(t.n_list_entries + 4) * 4, st);
    d->long_sp.ensure((size_t)(t.n_subpaths + 2) * 4, st);   // (scratch until the flush reuses it for the long sub-paths' blocks)
    if (d->failed) return 1;
    vkb_launch_decode_emit(d->dc_cmds.as<uint32_t>(), d->dc_args.as<float>(), nc, t, d->dc_S.as<uint32_t>(), d->dc_lists.as<uint32_t>(), irregular, d->elem_hdr.as<uint32_t>(),
                           d->elem_data.as<float>(), d->subpaths.as<vkb_subpath>(), d->draws.as<vkb_draw>(), d->xforms.as<vkb_xform>(), d->dc_xfscale.as<float>(),
                           d->strokes.as<vkb_stroke>(), d->grads.as<vkb_gradient>(), d->dashes.as<float>(), census, d->long_sp.as<uint32_t>(), init, st);
    VKB_CUDA_OK(cudaMemcpyAsync(d->dc_host + 128, census, sizeof(vkb_decode_census), cudaMemcpyDeviceToHost, st));
    VKB_CUDA_OK(cudaStreamSynchronize(st));
    if (d->failed) return 1;
    const vkb_decode_census c = *(const vkb_decode_census *)(d->dc_host + 128);
    if (c.irregular) {
        if (getenv("VKVG_B200_DEBUG")) fprintf(stderr, "vkvg_b200_submit: host decoder (flags 0x%x)\n", c.irregular);
        d->n_draws = 0; d->n_elems = 0; d->n_sp = 0;
        return 2;
    }  // (the resident batch was overwritten: nothing to replay)
    *census_out = c;
    // ---- the host's view of the batch, then the job tables and the pipeline as after an upload ----
    d->n_elems = c.n_elems; d->n_sp = c.n_subpaths; d->n_draws = c.n_draws; d->n_curves = c.n_curves; d->n_grads = c.n_grads + 1;
    d->n_fjobs = c.n_fjobs; d->n_sjobs = c.n_sjobs; d->n_sdraws = c.n_sdraws; d->n_extra = 0;
    d->any_dash = c.any_dash != 0; d->nz_any = c.nz_any != 0;
    d->max_sp_elems = c.max_sp_elems;
    d->has_clip_draws = d->has_stencil_ops = false;
    d->stencil_after = 0;
    d->h2d_bytes = (uint64_t)nc * 4 + (uint64_t)na * 4;
    build_job_tables(d);
    return run_flush(d, surf, samples, nullptr, stats, true) | d->failed;
}

int vkb_device_ordinal(vkb_device_impl *d) { return d->ordinal; }

// raw directed edges as a single non-zero draw; returns the per-sample winding the fine pass computed
int vkb_winding_raw(vkb_device_impl *d, uint32_t samples, const int32_t *edges_h, uint64_t n, uint32_t w, uint32_t h, int32_t *out) {
    dev_enter(d);
    cudaStream_t st = d->stream;
    finish_pending(d);
    VKB_CUDA_OK(cudaStreamSynchronize(st));
    vkb_surface_impl *surf = vkb_surface_new(d, w, h, h, 0);
    SurfaceDesc sd = {w, h, samples, (w + VKB_TILE - 1) / VKB_TILE, (h + VKB_TILE - 1) / VKB_TILE, h, 0, 0};
    d->totals.ensure(16 * 8, st);
    d->edges.ensure((n + 1) * 16, st);
    d->edge_draw.ensure((n + 1) * 4, st);
    d->paints.ensure(sizeof(vkb_paint), st);
    d->grads.ensure(sizeof(vkb_gradient), st);
    vkb_paint p = {VKB_RULE_NON_ZERO | (VKB_PAT_SOLID << 8), 0xffffffffu, 1.0f, 0};
    VKB_CUDA_OK(cudaMemcpyAsync(d->paints.p, &p, sizeof p, cudaMemcpyHostToDevice, st));
    if (n) VKB_CUDA_OK(cudaMemcpyAsync(d->edges.p, edges_h, n * 16, cudaMemcpyHostToDevice, st));
    VKB_CUDA_OK(cudaMemsetAsync(d->edge_draw.p, 0, (n + 1) * 4, st));
    VKB_CUDA_OK(cudaStreamSynchronize(st));
    vkb_capture cap;
    cap.winding = out; cap.winding_draw = 0;
    const SurfFlags before = surf_flags(surf);
    int r = 1;
    for (int attempt = 0; attempt < 16; attempt++) {
        surf_restore(surf, before);
        // capacities for the binning stages only; the edge count is given
        uint32_t *c = d->capv;
        if (c[VKC_EDGES] < n) c[VKC_EDGES] = (uint32_t)n;
        const uint64_t n_tiles = (uint64_t)sd.tiles_x * sd.tiles_y;
        if (c[VKC_PT] < n_tiles + 1024) c[VKC_PT] = (uint32_t)(n_tiles + 1024);
        if (c[VKC_ROWS] < c[VKC_PT]) c[VKC_ROWS] = c[VKC_PT];
        if (c[VKC_NE] < c[VKC_PT]) c[VKC_NE] = c[VKC_PT];
        if (c[VKC_TE] < 2 * n + 1024) c[VKC_TE] = (uint32_t)(2 * n + 1024);
        enqueue_counts_reset(d, nullptr, 0);
        set_edge_count_k<<<1, 1, 0, st>>>(d->counts.as<vkb_counts>(), (uint32_t)n);
        VKB_LAUNCHED();
        DevBuf wbuf;
        enqueue_bin_and_fine(d, surf, sd, 1, &cap, nullptr, wbuf);
        VKB_CUDA_OK(cudaMemcpyAsync(d->counts_host, d->counts.p, sizeof(vkb_counts), cudaMemcpyDeviceToHost, st));
        VKB_CUDA_OK(cudaStreamSynchronize(st));
        if (d->counts_host->overflow) { grow_caps_from_need(d, *d->counts_host); wbuf.release(); continue; }
        VKB_CUDA_OK(cudaMemcpyAsync(out, wbuf.p, (size_t)w * h * (samples ? samples : 1) * 4, cudaMemcpyDeviceToHost, st));
        VKB_CUDA_OK(cudaStreamSynchronize(st));
        wbuf.release();
        r = g_cuda_failed;
        break;
    }
    vkb_surface_free(surf);
    return r;
}

// bench support: `steps` resident replays, each timed on its own with CUDA events on the pipeline stream; between
// steps (outside the timed events) a 256 MiB scratch buffer is overwritten so that no step starts with its inputs in L2.
int vkb_time_resident(vkb_device_impl *d, vkb_surface_impl *s, uint32_t samples, uint32_t steps, bool clear_first, bool flush_l2, vkb_stats *sum) {
    dev_enter(d);
    finish_pending(d);
    vkb_stats acc;
    memset(&acc, 0, sizeof acc);
    for (uint32_t i = 0; i < steps; i++) {
        if (flush_l2) {
            d->l2_flush.ensure((size_t)256 << 20, d->stream);
            VKB_CUDA_OK(cudaMemsetAsync(d->l2_flush.p, (int)(i & 0xff), (size_t)256 << 20, d->stream));
        }
        if (clear_first) {  // BASELINE C1 counts clear + render + flush as the frame: the clear runs inside the timed events
            VKB_CUDA_OK(cudaEventRecord(d->ev_begin, d->stream));
            vkb_surface_clear(s);
            d->begin_recorded = true;
        }
        vkb_stats st;
        if (vkb_render_resident(d, s, samples, nullptr, &st)) return 1;
        float tot = acc.ms_total + st.ms_total, fine = acc.ms_fine + st.ms_fine, stage[VKB_N_STAGES];
        for (int k = 0; k < VKB_N_STAGES; k++) stage[k] = acc.ms_stage[k] + st.ms_stage[k];
        acc = st;
        acc.ms_total = tot; acc.ms_fine = fine;
        for (int k = 0; k < VKB_N_STAGES; k++) acc.ms_stage[k] = stage[k];
    }
    if (sum) *sum = acc;
    return g_cuda_failed;
}

// stage timing on: every stats-producing flush records events between its stages (plain launches).  Off: such flushes may
// replay the cached CUDA graph and report only the whole-flush time (and the fine kernel's, when the graph carries events).
void vkb_device_set_stage_timing(vkb_device_impl *d, bool on) { d->stage_timing = on; }
void vkb_device_set_graphs(vkb_device_impl *d, bool on) { d->graphs_enabled = on; }
unsigned long long vkb_device_graph_replays(vkb_device_impl *d) { return d->n_graph_replays; }
