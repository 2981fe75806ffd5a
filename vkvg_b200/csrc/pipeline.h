// Internal interface between the CUDA stages (flatten.cu, stroke.cu, raster.cu) and the orchestrator
// (pipeline.cu).  Nothing here is exported from the shared library.
#pragma once
#include "dev_util.cuh"
#include "vkb_types.h"

// ---- flatten.cu ----
// cache: n * cache_n points kept by the counting pass for the emitting pass (null: every curve is walked twice)
void vkb_launch_flatten_count(const uint32_t *elem_hdr, const float *elem_data, uint32_t n, uint32_t *counts, float2 *cache, uint32_t cache_n, cudaStream_t s);
void vkb_launch_flatten_emit(const uint32_t *elem_hdr, const float *elem_data, uint32_t n, const uint32_t *offsets, float2 *pts, uint8_t *flags,
                             const vkb_counts *C, float2 *cache, uint32_t cache_n, cudaStream_t s);
void vkb_launch_subpath_ranges(const vkb_subpath *sps, uint32_t n_sp, const uint32_t *elem_off, uint32_t n_elems, const uint32_t *total,
                               uint32_t *sp_first, uint32_t *sp_count, cudaStream_t s);

// ---- job tables: one job = one sub-path of one draw; items = its points ----
// sp_bbox: user-space box of every sub-path (vkb_launch_sp_bounds), or null for no culling; job_n = 0 for jobs that cannot touch the surface
#define VKB_SP_LONG 1024  // sub-paths of more elements are reduced by sp_bounds_long_k (one thread per element) instead of by one warp
struct SurfaceDesc;
// any_long: some sub-path holds more than VKB_SP_LONG elements (the host knows from the recorder or the decoder's census)
void vkb_launch_sp_bounds(const vkb_subpath *sps, uint32_t n_sp, const uint32_t *elem_hdr, const float *elem_data, uint32_t n_elems, uint32_t *long_blocks, uint32_t *n_long_blocks,
                          ScanScratch &scan, int4 *sp_bbox, bool any_long, cudaStream_t s);

// ---- stroke.cu ----
struct StrokeArgs {
    const float2      *pts;
    const uint8_t     *ptflags;
    const vkb_draw    *draws;
    const vkb_stroke  *strokes;
    const float       *dash_table;
    const uint32_t    *job_draw, *job_sp, *job_base;
    uint32_t           n_jobs;
    const uint32_t    *sp_first, *sp_count;
    const vkb_subpath *sps;
    const double      *cum;  // exclusive scan of segment lengths over all stroke items (+1), or null if nothing is dashed
    uint32_t           n_items;  // CAPACITY of the item space (grid size); the count is C->n[VKC_SITEMS]
    const vkb_counts  *C;
};
void vkb_launch_stroke_seglen(const StrokeArgs &a, float *seglen, cudaStream_t s);
void vkb_launch_stroke_count(const StrokeArgs &a, unsigned long long *counts, cudaStream_t s);
void vkb_launch_stroke_emit(const StrokeArgs &a, const unsigned long long *offsets, float2 *verts, uint32_t *inds, uint32_t *job_inverse, cudaStream_t s);

// ---- decode.cu: the packed command stream decoded on the device (vkvg_b200_submit) ----
#include "decode_types.h"
size_t   vkd_scan_words(uint32_t n_cmds);
size_t   vkd_blocksum_words(uint32_t n_cmds);
// phase A: the scan over the commands; S row n_cmds = the totals (vkd_read_totals picks out what the host needs to size phase B)
void vkb_launch_decode_scan(const uint32_t *cmds, uint32_t n_cmds, uint32_t *S, uint32_t *blocksum, uint32_t *irregular, cudaStream_t st);
struct vkd_totals { uint32_t n_elems, n_data, n_subpaths, n_draws, n_grads, n_dash_floats, n_xforms, n_strokes, n_list_entries; };
vkd_totals vkd_read_totals(const uint32_t *totals_row_host);
uint32_t   vkd_n_fields();
// phase B: the records
void vkb_launch_decode_emit(const uint32_t *cmds, const float *args, uint32_t n_cmds, const vkd_totals &t, uint32_t *S, uint32_t *lists, uint32_t *irregular,
                            uint32_t *elem_hdr, float *elem_data, vkb_subpath *subpaths, vkb_draw *draws, vkb_xform *xforms, float *xf_scale, vkb_stroke *strokes,
                            vkb_gradient *grads, float *dashes, vkb_decode_census *census, uint32_t *sp_null, const vkb_decode_init &in, cudaStream_t st);

// ---- raster.cu ----
struct SurfaceDesc {
    uint32_t width, height, samples;
    uint32_t tiles_x, tiles_y;
    // a stripe of a larger logical surface (multi-GPU tile-row sharding): the vertex stage and the paint evaluation use
    // the logical height, and snapped y coordinates are shifted by origin_y pixels (a multiple of the tile size), so a
    // stripe holds exactly the pixels the same rows of the whole surface would hold
    uint32_t full_height, origin_y;
    // a batch of independent canvases stacked vertically in one surface (vkvg_b200_surface_create_batch): canvas b owns the
    // band_tiles tile rows from b * band_tiles; every draw carries its canvas (vkb_xform.band), its geometry is snapped in
    // canvas coordinates (so each canvas holds exactly the pixels it would hold alone) and shifted by whole tiles.  0: no bands
    uint32_t band_tiles;
};
// the emitter of stroke.cu that runs the vertex stage itself: snapped = every vertex on the 1/256 grid (what vkb_launch_tri_edges reads), verts =
// the float vertices for geometry captures or null.  vkb_stroke_emit_mode(): 0 this emitter with block-staged 16-byte stores, 1 the first
// emitter (vkb_launch_stroke_emit + snap_verts_k), 2 this emitter writing straight to global memory (VKVG_B200_STROKE=legacy|direct, for A/B runs)
int  vkb_stroke_emit_mode();
void vkb_launch_stroke_emit_snapped(const StrokeArgs &a, const unsigned long long *offsets, const vkb_xform *xforms, const SurfaceDesc &sd, float2 *verts, int2 *snapped,
                                    uint32_t *inds, uint32_t *job_inverse, cudaStream_t s);

// ---- vertex stage: shaders/vkvg_main.vert:74-79 + viewport + 8-bit sub-pixel snap (round half up) ----
__device__ __forceinline__ void vs_snap(const float *m, float W, float H, float x, float y, int32_t &fx, int32_t &fy) {
    float px = m[0] * x + m[2] * y + m[4];
    float py = m[1] * x + m[3] * y + m[5];
    // x / 2^k and x * 2^-k are the same correctly rounded value, and every surface of the benchmark configurations is a power of two wide and
    // high: the two IEEE divisions were half of the vertex stage's instructions (ncu, profiles/r2y_stroke_c3_lines.txt)
    const uint32_t wb = __float_as_uint(W), hb = __float_as_uint(H);
    float nx = (wb & 0x007FFFFFu) == 0u ? px * 2.0f * __uint_as_float(0x7F000000u - wb) - 1.0f : px * 2.0f / W - 1.0f;
    float ny = (hb & 0x007FFFFFu) == 0u ? py * 2.0f * __uint_as_float(0x7F000000u - hb) - 1.0f : py * 2.0f / H - 1.0f;
    float wx = nx * (W * 0.5f) + (W * 0.5f);
    float wy = ny * (H * 0.5f) + (H * 0.5f);
    // clamp far outside the guard band so the int32 conversion is defined; such coordinates are off-surface
    wx = fminf(fmaxf(wx, -1.0e6f), 1.0e6f);
    wy = fminf(fmaxf(wy, -1.0e6f), 1.0e6f);
    fx = (int32_t)floorf(wx * 256.0f + 0.5f);
    fy = (int32_t)floorf(wy * 256.0f + 0.5f);
}

void vkb_launch_job_counts(const uint32_t *job_sp, const uint32_t *job_draw, uint32_t n_jobs, const uint32_t *sp_count, uint32_t min_points, const int4 *sp_bbox,
                           const vkb_draw *draws, const vkb_xform *xforms, const vkb_stroke *strokes, SurfaceDesc sd, uint32_t *job_n, cudaStream_t s);
void vkb_launch_job_counts2(const uint32_t *fjob_sp, const uint32_t *fjob_draw, uint32_t nf, uint32_t *fjob_n, const uint32_t *sjob_sp, const uint32_t *sjob_draw, uint32_t ns,
                            uint32_t *sjob_n, const uint32_t *sp_count, const int4 *sp_bbox, const vkb_draw *draws, const vkb_xform *xforms, const vkb_stroke *strokes,
                            SurfaceDesc sd, cudaStream_t s);
void vkb_launch_fill_edges(const float2 *pts, const vkb_draw *draws, const vkb_xform *xforms, const uint32_t *job_draw, const uint32_t *job_sp, const uint32_t *job_base,
                           uint32_t n_jobs, const uint32_t *sp_first, const uint32_t *sp_count, uint32_t cap_items, const vkb_counts *C, SurfaceDesc sd,
                           vkb_edge *edges, uint32_t *edge_draw, int32_t *draw_bbox, cudaStream_t s);
// NON_ZERO fills / clips as the reference's libtess makes them (raster.cu): nz_mode per draw (0 none, 1 fan fast path -> COUNT rule, 2 edges
// split at their crossings, 3 too large to split), then the pieces in one pass.  draw_first_job: exclusive scan of fill jobs per draw
void vkb_launch_nz_classify(const vkb_draw *draws, uint32_t n_draws, const uint32_t *sp_first, const uint32_t *sp_count, const float2 *pts, const vkb_counts *C,
                            vkb_paint *paints, uint8_t *nz_mode, cudaStream_t s);
// one pass over the fill items: pieces written where the warp reserved room (*n_out: zeroed device counter), then VKC_FEDGES committed from it
void vkb_launch_nz_split(const float2 *pts, const vkb_draw *draws, const vkb_xform *xforms, const uint32_t *job_draw, const uint32_t *job_sp, const uint32_t *job_base,
                         uint32_t n_jobs, const uint32_t *sp_first, const uint32_t *sp_count, const uint32_t *draw_first_job, uint32_t n_draws, const uint8_t *nz_mode,
                         uint32_t cap_items, vkb_counts *C, SurfaceDesc sd, vkb_edge *edges, uint32_t *edge_draw, uint32_t *n_out, uint32_t cap_edges, int32_t *draw_bbox, cudaStream_t s);
// edges / edge_draw: start of the edge arrays (the kernel skips the C->n[VKC_FEDGES] fill edges and the n_extra rectangle edges itself);
// live: zeroed device counter of the stroke edges stored (cancelled ones are dropped); Cw->n[VKC_EDGES] is set to the stored total
// (commit_live; otherwise the caller passes live on to vkb_launch_draw_rects, which sets it)
// snapped: cap_verts int2 (every stroke vertex goes through the vertex stage once, then the triangles read integers): filled here from verts
// by snap_verts_k, or already by the emitter (verts == null)
void vkb_launch_tri_edges(const float2 *verts, uint32_t cap_verts, int2 *snapped, const uint32_t *inds, uint32_t cap_tris, const vkb_counts *C,
                          const vkb_draw *draws, const vkb_xform *xforms, const uint32_t *sdraw_id, const uint32_t *sdraw_first_item, uint32_t n_sdraws,
                          const unsigned long long *item_offsets, SurfaceDesc sd, vkb_edge *edges, uint32_t *edge_draw, uint32_t n_extra, uint32_t *live,
                          vkb_counts *Cw, int32_t *draw_bbox, bool commit_live, cudaStream_t s);

struct BinBuffers {  // all device pointers
    int32_t  *draw_bbox;    // n_draws x 4 (minx, miny, maxx, maxy), fixed point
    int32_t  *draw_rect;    // n_draws x 4 (tx0, ty0, tw, th) in tiles
    uint32_t *draw_ptbase;  // n_draws (+1): first path-tile of the draw
    uint32_t *draw_rowbase; // n_draws (+1): first path-tile row of the draw
};
// every launcher below sizes its grid for a capacity (cap_*) and reads the count from C (dev_util.cuh: vkb_counts)
// every kernel that emits edges of a draw grows draw_bbox[draw] itself (the boxes are emptied by the flush's counts_reset_k); vkb_launch_draw_bbox_init /
// vkb_launch_draw_bbox are for raw edge lists
void vkb_launch_draw_bbox_init(uint32_t n_draws, int32_t *draw_bbox, cudaStream_t s);
void vkb_launch_draw_bbox(const vkb_edge *edges, const uint32_t *edge_draw, uint64_t cap_edges, const vkb_counts *C, uint32_t n_draws, int32_t *draw_bbox,
                          cudaStream_t s);
// gradients whose position-independent terms (FineArgs::gprep, 16 floats each) the same launch evaluates; n == 0: none
struct GradPrep { const vkb_gradient *grads; uint32_t n; float W, H; float *out; };
// draws may be null (raw edge lists); a VKB_DRAW_CLIP draw takes the whole surface as its rectangle
void vkb_launch_draw_rects(const int32_t *draw_bbox, const vkb_draw *draws, const vkb_xform *xforms, uint32_t n_draws, SurfaceDesc sd, int32_t *draw_rect,
                           unsigned long long *tile_row_counts, vkb_counts *Cw, const uint32_t *live, uint32_t n_extra, const GradPrep &gp, cudaStream_t s);
// (live: the counter vkb_launch_tri_edges filled when it was told not to commit it, or null)
// + VKC_PT / VKC_ROWS committed from *total (path-tiles | rows << 32, the scan's total)
void vkb_launch_split_bases(const unsigned long long *packed, uint32_t n, uint32_t *lo, uint32_t *hi, vkb_counts *C, const unsigned long long *total, cudaStream_t s);
void vkb_launch_bin_count(const vkb_edge *edges, const uint32_t *edge_draw, uint64_t cap_edges, const vkb_counts *C, const int32_t *draw_rect,
                          const uint32_t *draw_ptbase, uint32_t *pt_count, int32_t *pt_backdrop, uint32_t *long_list, uint32_t *long_n, cudaStream_t s);
// pt_owner[path-tile] / row_owner[path-tile row] = draw index; pt_count / pt_backdrop of every path-tile zeroed
void vkb_launch_owners(const int32_t *draw_rect, const uint32_t *draw_ptbase, const uint32_t *draw_rowbase, uint32_t n_draws, const vkb_counts *C,
                       uint32_t *pt_owner, uint32_t *row_owner, uint32_t *pt_count, int32_t *pt_backdrop, cudaStream_t s);
void vkb_launch_backdrop_prefix(const int32_t *draw_rect, const uint32_t *draw_ptbase, const uint32_t *draw_rowbase, const uint32_t *row_owner,
                                uint32_t cap_rows, const vkb_counts *C, int32_t *pt_backdrop, cudaStream_t s);
// keep_clip: the batch holds VKB_DRAW_CLIP draws, whose path-tiles are all kept (an empty one means "clipped out")
void vkb_launch_pt_flags(const uint32_t *pt_count, const int32_t *pt_backdrop, uint32_t cap_pt, const vkb_counts *C, const vkb_draw *draws,
                         const uint32_t *pt_owner, bool keep_clip, uint32_t *flags, cudaStream_t s);
void vkb_launch_pt_compact(const uint32_t *flags, const uint32_t *flag_scan, uint32_t cap_pt, const vkb_counts *C, const int32_t *draw_rect,
                           const uint32_t *draw_ptbase, const uint32_t *pt_owner, SurfaceDesc sd, uint32_t *keys, uint32_t *vals, uint32_t *pt_draw, cudaStream_t s);
// also zeroes cursor[0, n_ne), tile_first / tile_end[0, n_tiles) and, if given, the per-tile multisample flags (as words)
void vkb_launch_sorted_counts(const uint32_t *vals, uint32_t cap_ne, const vkb_counts *C, const uint32_t *pt_count, uint32_t *sorted_cnt, uint32_t *pt_slot,
                              uint32_t *cursor, uint32_t *tile_first, uint32_t *tile_end, uint32_t n_tiles, uint32_t *tile_ms_words, cudaStream_t s);
void vkb_launch_headers(const uint32_t *keys, const uint32_t *vals, uint32_t cap_ne, const vkb_counts *C, const uint32_t *pt_draw_by_flagpos,
                        const uint32_t *flag_scan, const int32_t *pt_backdrop, const uint32_t *pt_count, const uint32_t *eoff, const vkb_paint *paints, int4 *hdr,
                        uint32_t *tile_first, uint32_t *tile_end, cudaStream_t s);
void vkb_launch_bin_scatter(const vkb_edge *edges, const uint32_t *edge_draw, uint64_t cap_edges, const vkb_counts *C, const int32_t *draw_rect,
                            const uint32_t *draw_ptbase, const uint32_t *pt_slot, const uint32_t *eoff, uint32_t *cursor, vkb_edge *tile_edges,
                            const uint32_t *long_list, const uint32_t *long_n, cudaStream_t s);

struct FineArgs {
    SurfaceDesc         sd;
    const vkb_counts   *counts;      // the fine pass only runs when no intermediate overflowed
    const uint32_t     *tile_first, *tile_end;
    const int4         *hdr;         // per sorted path-tile, 2 x int4: {draw, backdrop, edge offset, edge count}, {vkb_paint of the draw}
    const vkb_edge     *tile_edges;
    const vkb_paint    *paints;      // per draw
    const vkb_gradient *grads;
    const vkb_surfpat  *surfpats;    // surface paints of the batch
    const float        *gprep;       // per gradient: VKB_GPREP_FLOATS position-independent terms of the paint evaluation (grad_prep_one, run by draw_rects_k)
    uint32_t           *image;       // width*height premultiplied RGBA8 (resolved)
    uint32_t           *ms_image;    // per-sample colours, tile-major [tile][256][S]; valid for tiles whose tile_ms flag is set
    uint8_t            *tile_ms;     // per tile: 1 when the samples of some pixel differ (the resolved image alone would lose them)
    uint32_t           *ms_mask;     // per tile: 8 words (one per warp = two pixel rows), bit = that pixel's samples are in ms_image
    uint32_t            tile_lo, tile_hi;  // tiles this launch renders (all of them, or a band of tile rows: pipeline.cu, read-back overlap)
    uint32_t           *tile_counter; // zeroed before the launch: the warp-per-tile kernel hands out tiles (relative to tile_lo) from it
    int32_t            *wscratch;    // vkb_fine_wscratch_words() int32: one per-sample winding plane per resident warp of fine_warp_k (COUNT rule / huge lists only)
    int                 dst_is_clear;  // destination known to be transparent black: do not read it
    uint32_t           *stencil;     // per-sample stencil bytes (clip bit + save bits), tile-major [tile][256][ceil(S/4)] words; null: no clip in play
    int                 stencil_in;  // the plane holds state from earlier flushes (else it is taken as all zero)
    int32_t            *winding_out; // optional: per-sample winding of the LAST draw touching each sample (parity tests), or null
    uint32_t            winding_draw; // draw index captured into winding_out
};
void vkb_launch_fine(const FineArgs &a, cudaStream_t s);
size_t vkb_fine_wscratch_words(uint32_t samples);
// fine kernel for batches without clip state or winding capture: 0 = chosen by tile count, 1 = block-per-tile fine_k, 2 = warp-per-tile
// fine_warp_k.  Process-wide; VKVG_B200_FINE=block|warp sets it at start-up
void vkb_fine_set_mode(int mode);
int  vkb_fine_get_mode();

void vkb_launch_unpremultiply(const uint32_t *image, uint64_t n_pixels, uint32_t *out, cudaStream_t s);
