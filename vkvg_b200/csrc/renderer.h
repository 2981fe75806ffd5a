// Host-side C++ interface of the CUDA pipeline, used by the C API layer (vkvg_api.cpp).
#pragma once
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include "vkb_types.h"
#include "decode_types.h"

// Growable array of plain data that never value-initialises (std::vector::resize would write every element of a
// 1M-point polyline twice) and keeps its capacity across clear().
void *vkb_host_alloc(size_t bytes);  // pinned when a CUDA device is present (pipeline.cu)
void  vkb_host_free(void *p);
template <class T> struct PodVec {
    T     *p_ = nullptr;
    size_t n_ = 0, cap_ = 0;
    PodVec() {}
    PodVec(const PodVec &o) { assign(o.p_, o.p_ + o.n_); }
    PodVec &operator=(const PodVec &o) { if (this != &o) assign(o.p_, o.p_ + o.n_); return *this; }
    ~PodVec() { vkb_host_free(p_); }
    void reserve(size_t c) {
        if (c <= cap_) return;
        size_t nc = cap_ ? cap_ : 1024;
        while (nc < c) nc *= 2;
        T *np = (T *)vkb_host_alloc(nc * sizeof(T));
        if (n_) memcpy(np, p_, n_ * sizeof(T));
        vkb_host_free(p_);
        p_   = np;
        cap_ = nc;
    }
    void   resize(size_t n) { reserve(n); n_ = n; }  // new elements are uninitialised
    void   clear() { n_ = 0; }
    size_t size() const { return n_; }
    bool   empty() const { return n_ == 0; }
    T     *data() { return p_; }
    const T *data() const { return p_; }
    T       &operator[](size_t i) { return p_[i]; }
    const T &operator[](size_t i) const { return p_[i]; }
    T       *begin() { return p_; }
    T       *end() { return p_ + n_; }
    const T *begin() const { return p_; }
    const T *end() const { return p_ + n_; }
    void     push_back(const T &v) { if (n_ == cap_) reserve(n_ + 1); p_[n_++] = v; }
    void     append(const T *a, const T *b) { size_t k = (size_t)(b - a); reserve(n_ + k); memcpy(p_ + n_, a, k * sizeof(T)); n_ += k; }
    void     assign(const T *a, const T *b) { n_ = 0; append(a, b); }
};

// One flush worth of recorded work (host memory).  See vkb_types.h for the record formats.
struct vkb_batch {
    PodVec<uint32_t>          elem_hdr;
    PodVec<float>             elem_data;
    std::vector<vkb_subpath>  subpaths;
    std::vector<vkb_draw>     draws;
    std::vector<vkb_xform>    xforms;
    std::vector<vkb_stroke>   strokes;
    std::vector<vkb_gradient> grads;
    std::vector<float>        dashes;
    std::vector<vkb_surfpat>  surfpats;
    uint32_t                  n_curves = 0;  // cubic / arc elements among elem_hdr (sizing hint for the flatten stage)
    void clear_draws() { draws.clear(); xforms.clear(); strokes.clear(); grads.clear(); dashes.clear(); surfpats.clear(); }
    void clear() {
        elem_hdr.clear(); elem_data.clear(); subpaths.clear(); clear_draws(); n_curves = 0;
    }
};

// Optional host-side copies of intermediate results (parity tests only; null members are skipped)
struct vkb_capture {
    std::vector<float>    *points    = nullptr;  // x,y per flattened point
    std::vector<uint8_t>  *ptflags   = nullptr;
    std::vector<uint32_t> *sp_first  = nullptr, *sp_count = nullptr;
    std::vector<float>    *verts     = nullptr;  // stroke vertices x,y
    std::vector<uint32_t> *inds      = nullptr;  // stroke indices (absolute into verts)
    std::vector<int32_t>  *edges     = nullptr;  // x0,y0,x1,y1 fixed point
    std::vector<uint32_t> *edge_draw = nullptr;
    int32_t               *winding   = nullptr;  // width*height*samples, winding of draw `winding_draw`
    uint32_t               winding_draw = 0;
    bool                   geometry_only = false;  // stop before binning / fine
};

#define VKB_N_STAGES 5  // flatten, stroke, edges, binning, fine
struct vkb_stats {            // filled by every render
    uint64_t n_elems, n_points, n_fill_edges, n_stroke_items, n_verts, n_inds, n_edges, n_path_tiles, n_nonempty, n_tile_edges;
    float    ms_total, ms_fine;  // CUDA-event durations on the device stream
    uint64_t h2d_bytes;
    float    ms_stage[VKB_N_STAGES];  // flatten | job tables + stroke expansion | edge build | binning + sort | fine pass
    float    ms_host_upload;          // wall clock of the host side of the last upload (job tables, staging copy, H2D enqueue)
};

struct vkb_device_impl;
struct vkb_surface_impl;

vkb_device_impl *vkb_device_open(int ordinal);  // nullptr when no usable CUDA device
void             vkb_device_close(vkb_device_impl *d);
int              vkb_device_failed(vkb_device_impl *d);
void             vkb_device_sync(vkb_device_impl *d);
void             vkb_device_set_stage_timing(vkb_device_impl *d, bool on);
void             vkb_device_set_graphs(vkb_device_impl *d, bool on);
void             vkb_fine_set_mode(int mode);  // raster.cu: which fine kernel serves batches without clip state (process-wide)
int              vkb_fine_get_mode();
unsigned long long vkb_device_graph_replays(vkb_device_impl *d);

vkb_surface_impl *vkb_surface_new(vkb_device_impl *d, uint32_t w, uint32_t h, uint32_t full_h, uint32_t origin_y);
int               vkb_surface_copy_to_device(vkb_surface_impl *s, void *dst);
int               vkb_surface_ipc_export(vkb_surface_impl *s, void *handle64);
void             *vkb_ipc_open(vkb_device_impl *d, const void *handle64);
int               vkb_ipc_close(vkb_device_impl *d, void *p);
void              vkb_surface_free(vkb_surface_impl *s);
void              vkb_surface_clear(vkb_surface_impl *s);
// clip support: forget the stencil plane (a new context clears the stencil attachment with its first render pass,
// src/vkvg_context.c:44-49) and the whole-plane spill used every six nested clip saves
void              vkb_surface_stencil_reset(vkb_surface_impl *s);
void              vkb_surface_set_band_height(vkb_surface_impl *s, uint32_t band_h);  // batch of canvases stacked in one surface
int               vkb_surface_stencil_push(vkb_surface_impl *s, uint32_t samples);
int               vkb_surface_stencil_pop(vkb_surface_impl *s, uint32_t samples);
// premultiplied RGBA8 rows, or un-premultiplied as vkvg_surface_write_to_memory does; synchronous
int vkb_surface_download(vkb_surface_impl *s, uint8_t *out, bool unpremultiply);
// host memory every later flush also delivers the premultiplied image to, band by band while it renders (NULL: off)
void vkb_surface_set_readback(vkb_surface_impl *s, uint8_t *host);
const uint32_t *vkb_surface_device_pixels(vkb_surface_impl *s);
int             vkb_surface_upload(vkb_surface_impl *s, const uint8_t *rgba);

// upload `b` and render it onto `s`.  Returns 0 on success.
int vkb_render(vkb_device_impl *d, vkb_surface_impl *s, uint32_t samples, const vkb_batch &b, vkb_capture *cap, vkb_stats *stats);
// upload only (bench: inputs resident in HBM); then vkb_render_resident re-runs the pipeline on that batch
int vkb_upload(vkb_device_impl *d, const vkb_batch &b);
int vkb_render_resident(vkb_device_impl *d, vkb_surface_impl *s, uint32_t samples, vkb_capture *cap, vkb_stats *stats);
// the packed command stream decoded on the device (decode.cu) and rendered: 0 queued, 1 device error, 2 not decodable on the device (nothing touched)
int vkb_submit_stream(vkb_device_impl *d, vkb_surface_impl *surf, uint32_t samples, const uint32_t *cmds, uint64_t n_cmds, const float *args, uint64_t n_args,
                      const vkb_decode_init &init, vkb_decode_census *census_out, vkb_stats *stats);
int vkb_time_resident(vkb_device_impl *d, vkb_surface_impl *s, uint32_t samples, uint32_t steps, bool clear_first, bool flush_l2, vkb_stats *sum);
