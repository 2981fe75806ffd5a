/*
 * vkvg_b200.h — extension entry points of libvkvg_b200.so beyond the vkvg.h drop-in surface.
 *
 * Plain C ABI (pointers and sizes only).  Three groups:
 *   1. stage introspection for parity tests: run the CUDA flatten / stroke / raster stages on the context's
 *      current path or on raw edges and copy the intermediate results to host buffers,
 *   2. measurement: launch counters, device-timed stage durations, a resident-batch replay so that the bench
 *      can time the pipeline with its inputs already in HBM,
 *   3. a packed command stream (`vkvg_b200_replay`) that drives the ordinary vkvg_* calls from arrays, so a
 *      host program in any language can submit a whole scene with one FFI call instead of one per point
 *      (the reference's equivalent is its optional recording facility, src/recording/vkvg_record_internal.h:31-95).
 */
#ifndef VKVG_B200_H
#define VKVG_B200_H
#include "vkvg.h"
#ifdef __cplusplus
extern "C" {
#endif

/* ---- 0. coverage mode -------------------------------------------------------------------------------------- */
/* MSAA (default): per-sample integer winding at the sample positions of the device's sample count, per-sample blend and
 * box resolve — the mode that is bit-exact with the reference's rasterisation.
 * ANALYTIC: coverage of a pixel = exact area of the shape inside it (integral of the winding number over the pixel
 * square, then the fill rule), one colour per pixel, the paint scaled by the coverage blended once.  Not bit-comparable
 * with the reference (it has no such mode); strokes cover the UNION of their triangles, so self-overlapping translucent
 * strokes are not double-blended as the reference does.  Select it once, before the first draw on the device's surfaces
 * (or set VKVG_B200_COVERAGE=analytic in the environment of an unmodified vkvg client). */
enum { VKVG_B200_COVERAGE_MSAA = 0, VKVG_B200_COVERAGE_ANALYTIC = 1 };
vkvg_public vkvg_status_t vkvg_b200_device_set_coverage_mode(VkvgDevice dev, int mode);
vkvg_public int           vkvg_b200_device_get_coverage_mode(VkvgDevice dev);

/* ---- 1. stage introspection ------------------------------------------------------------------------- */

/* Flatten the context's current path on the GPU (replaces _recursive_bezier + arc loops,
 * src/vkvg_context_internal.c:1313-1461, src/vkvg_context.c:394-503).  Writes up to cap_points (x,y) pairs,
 * the per-point "curved segment" flag, and per sub-path first/count.  Returns the number of points (which may
 * exceed cap_points; nothing past the capacity is written).  The path is preserved. */
vkvg_public uint32_t vkvg_b200_flatten_path(VkvgContext ctx, float *xy, uint8_t *curved, uint32_t cap_points, uint32_t *sp_first,
                                            uint32_t *sp_count, uint32_t cap_subpaths, uint32_t *n_subpaths);

/* Stroke the current path with the current line state on the GPU and return the triangle list in the
 * reference's vertex order (replaces _stroke_preserve / _build_vb_step / _draw_stoke_cap / _draw_dashed_segment,
 * src/vkvg_context.c:822-948, src/vkvg_context_internal.c:924-1264).  Indices are relative to the first
 * vertex of the call.  Nothing is drawn; the path is preserved. */
vkvg_public void vkvg_b200_stroke_geometry(VkvgContext ctx, float *xy, uint32_t cap_verts, uint32_t *n_verts, uint32_t *indices,
                                           uint32_t cap_indices, uint32_t *n_indices);

/* Device-space edges (24.8 fixed point x0,y0,x1,y1) that filling or stroking the current path would hand to
 * the rasteriser.  kind: 0 = fill, 1 = stroke.  Returns the edge count.  Fill edges come one per path point, in path order, with
 * an edge that lies wholly above, below or right of the surface stored as 0,0,0,0 (it cannot change any sample); stroke edges
 * are only the ones that survive (an edge shared by two consecutive triangles in opposite directions cancels, off-surface
 * edges are dropped), in no particular order. */
vkvg_public uint64_t vkvg_b200_path_edges(VkvgContext ctx, int kind, int32_t *edges_xyxy, uint64_t cap_edges);

/* Flush the context; additionally copy the per-sample integer winding computed by the tile rasteriser for the
 * LAST draw of the flushed batch into winding (height*width*samples int32, 0 where that draw has no tile).
 * In ANALYTIC coverage mode the buffer receives height*width floats instead: the area integral A of that draw. */
vkvg_public void vkvg_b200_flush_capture_winding(VkvgContext ctx, int32_t *winding);

/* Run binning + fine pass on raw directed edges as one draw and return the per-sample winding
 * (height*width*samples).  This is the integer core that must match oracle ovk_winding_brute bit for bit.
 * In ANALYTIC coverage mode: height*width floats, the area integral A per pixel (oracle: ovk_area_brute). */
vkvg_public vkvg_status_t vkvg_b200_winding(VkvgDevice dev, const int32_t *edges_xyxy, uint64_t n_edges, uint32_t width, uint32_t height,
                                            int32_t *winding);

/* premultiplied RGBA8 pixels exactly as stored (vkvg_surface_write_to_memory un-premultiplies) */
vkvg_public vkvg_status_t vkvg_b200_surface_read_premultiplied(VkvgSurface surf, unsigned char *rgba);
/* Read-back overlapped with rendering: host_rgba (width * height * 4 bytes, pinned memory for the copies to be asynchronous) becomes the
 * place every later flush onto surf ALSO delivers the premultiplied image to - the fine pass then runs in bands of tile rows and each
 * finished band is copied on a second stream while the next ones render.  vkvg_b200_surface_read_premultiplied(surf, host_rgba) then only
 * waits for the last band instead of copying the whole image after the whole frame.  NULL switches it off; the memory must stay valid
 * until then. */
vkvg_public vkvg_status_t vkvg_b200_surface_set_readback(VkvgSurface surf, unsigned char *host_rgba);
/* The same delivery to ANOTHER GPU of the node (tile-row stripes of one picture, one process per GPU; no counterpart in the reference, which
 * renders a surface on one device): the root exports the image of its full-height surface as a 64-byte inter-process handle
 * (cudaIpcMemHandle_t), every other rank opens it - the pointer is the root's premultiplied RGBA8 image, row-major - and names
 * `ptr + y0 * width * 4` as the read-back target of its stripe surface: finished bands of the stripe then travel over NVLink on the copy
 * engines while later bands render, and no gather follows the frame.  The target of vkvg_b200_surface_set_readback may therefore be host
 * memory or device memory of any GPU this process can reach.  A rank must synchronise (vkvg_b200_surface_read_premultiplied onto the same
 * pointer returns once its bands are delivered) before it tells the root, by whatever barrier the ranks share, that its rows are there. */
vkvg_public vkvg_status_t vkvg_b200_surface_ipc_export(VkvgSurface surf, unsigned char *handle64);
vkvg_public void         *vkvg_b200_ipc_open(VkvgDevice dev, const unsigned char *handle64);
vkvg_public vkvg_status_t vkvg_b200_ipc_close(VkvgDevice dev, void *ptr);

/* ---- 2. measurement ------------------------------------------------------------------------------------ */
typedef struct {
    uint64_t n_elems, n_points, n_fill_edges, n_stroke_items, n_verts, n_inds, n_edges, n_path_tiles, n_nonempty, n_tile_edges;
    float    ms_total, ms_fine;
    uint64_t h2d_bytes;
    float    ms_stage[5]; /* flatten | job tables + stroke expansion | edge build | binning + sort | fine pass */
    float    ms_host_upload; /* wall clock of the host side of the last upload */
} vkvg_b200_stats_t;

vkvg_public uint64_t vkvg_b200_launch_count(void);                          /* kernels launched by this library so far */
vkvg_public void     vkvg_b200_set_profiling(VkvgDevice dev, int on);       /* on: every flush synchronises and records stats */
vkvg_public void     vkvg_b200_last_stats(VkvgDevice dev, vkvg_b200_stats_t *out);
vkvg_public void     vkvg_b200_device_synchronize(VkvgDevice dev);
vkvg_public int      vkvg_b200_device_ordinal(VkvgDevice dev);
vkvg_public const void *vkvg_b200_surface_device_pointer(VkvgSurface surf); /* CUDA device pointer of the RGBA8 image */

/* Resident replay: `vkvg_b200_flush_keep` flushes like vkvg_flush but keeps the uploaded batch on the device;
 * `vkvg_b200_replay_resident` re-runs the whole pipeline on it (no host->device traffic) onto surf. */
vkvg_public void vkvg_b200_flush_keep(VkvgContext ctx);
vkvg_public void vkvg_b200_replay_resident(VkvgDevice dev, VkvgSurface surf, int clear_first);
/* bench support: run `steps` resident replays, each timed on its own with CUDA events on the library's stream (which is
 * not torch's current stream); with flush_l2 a 256 MiB scratch buffer is overwritten before every step, outside the
 * timed events.  `sum` receives the counters of the last step and the SUMS of ms_total / ms_fine / ms_stage. */
vkvg_public vkvg_status_t vkvg_b200_time_resident(VkvgDevice dev, VkvgSurface surf, uint32_t steps, int clear_first, int flush_l2,
                                                  vkvg_b200_stats_t *sum);

/* ---- 2b. multi-GPU sharding ---------------------------------------------------------------- */
/* A surface that is the tile-row stripe [origin_y, origin_y + height) of a logical width x full_height surface
 * (origin_y a multiple of 16).  Drawing the SAME calls on it as on the whole surface leaves exactly the pixels those
 * rows of the whole surface would hold (the vertex stage and the paint evaluation use the logical size), so ranks can
 * render disjoint stripes with no exchange and gather the rows afterwards. */
vkvg_public VkvgSurface   vkvg_b200_surface_create_stripe(VkvgDevice dev, uint32_t width, uint32_t full_height, uint32_t origin_y, uint32_t height);
/* device-to-device copy of the premultiplied RGBA8 rows into caller-owned device memory (e.g. a buffer handed to an
 * NCCL gather); synchronous */
vkvg_public vkvg_status_t vkvg_b200_surface_copy_to_device(VkvgSurface surf, void *device_dst);

/* ---- 3. packed command stream -------------------------------------------------------------------------- */
enum {
    VKVG_B200_OP_MOVE_TO = 1,   /* x y */
    VKVG_B200_OP_LINE_TO,       /* x y */
    VKVG_B200_OP_CURVE_TO,      /* x1 y1 x2 y2 x3 y3 */
    VKVG_B200_OP_CLOSE_PATH,
    VKVG_B200_OP_NEW_PATH,
    VKVG_B200_OP_ARC,           /* xc yc r a1 a2 */
    VKVG_B200_OP_ARC_NEGATIVE,  /* xc yc r a1 a2 */
    VKVG_B200_OP_RECTANGLE,     /* x y w h */
    VKVG_B200_OP_FILL,
    VKVG_B200_OP_FILL_PRESERVE,
    VKVG_B200_OP_STROKE,
    VKVG_B200_OP_STROKE_PRESERVE,
    VKVG_B200_OP_PAINT,
    VKVG_B200_OP_SET_SOURCE_RGBA,   /* r g b a */
    VKVG_B200_OP_SET_LINE_WIDTH,    /* w */
    VKVG_B200_OP_SET_LINE_CAP,      /* cap */
    VKVG_B200_OP_SET_LINE_JOIN,     /* join */
    VKVG_B200_OP_SET_MITER_LIMIT,   /* limit */
    VKVG_B200_OP_SET_FILL_RULE,     /* rule */
    VKVG_B200_OP_SET_DASH,          /* n offset d0 .. dn-1 */
    VKVG_B200_OP_SET_SOURCE_LINEAR, /* x0 y0 x1 y1 nstops (offset r g b a)* */
    VKVG_B200_OP_SET_SOURCE_RADIAL, /* cx0 cy0 r0 cx1 cy1 r1 nstops (offset r g b a)* */
    VKVG_B200_OP_TRANSLATE,         /* dx dy */
    VKVG_B200_OP_SCALE,             /* sx sy */
    VKVG_B200_OP_ROTATE,            /* radians */
    VKVG_B200_OP_IDENTITY_MATRIX,
    VKVG_B200_OP_SAVE,
    VKVG_B200_OP_RESTORE,
    VKVG_B200_OP_CLEAR,
    VKVG_B200_OP_SET_OPACITY,       /* opacity */
    VKVG_B200_OP_POLYLINE,          /* n x0 y0 .. : move_to the first point, line_to the others; n is a uint32 stored in the float slot bit for bit */
    VKVG_B200_OP_FLUSH,
    VKVG_B200_OP_SET_CANVAS,        /* index (batch surfaces) */
    VKVG_B200_OP_CLIP,
    VKVG_B200_OP_CLIP_PRESERVE,
    VKVG_B200_OP_RESET_CLIP,
};
/* ops[i] selects the call; its float arguments are consumed from args in order.  Equivalent to issuing the
 * corresponding vkvg_* calls one by one (it does exactly that).  Returns the context status. */
vkvg_public vkvg_status_t vkvg_b200_replay(VkvgContext ctx, const uint8_t *ops, uint64_t n_ops, const float *args, uint64_t n_args);

/* The same commands with EXPLICIT argument counts: cmds[i] = op | n_args << 8 (n_args < 2^24), arguments in `args` in order.  Argument
 * lists as above except that counts are not repeated inside them: POLYLINE x0 y0 x1 y1 ..., SET_DASH offset d0 .. dn-1, SET_SOURCE_LINEAR
 * x0 y0 x1 y1 (offset r g b a)*, SET_SOURCE_RADIAL cx0 cy0 r0 cx1 cy1 r1 (offset r g b a)*.
 * vkvg_b200_submit = the calls one by one, then vkvg_flush - but the stream is uploaded as it is and turned into path elements, sub-paths,
 * draws and side tables BY KERNELS (vkvg_b200/csrc/decode.cu) when it is a regular bulk scene: move_to / line_to / curve_to / polyline /
 * close_path / new_path, fill / stroke (and _preserve), solid and gradient sources, line and dash state, fill rule, opacity,
 * identity_matrix / translate / set_canvas, starting from a context with no path under construction, a solid or gradient source and no clip.
 * Everything else (and any stream in which the reference would drop a point or ignore a close_path) is decoded on the host exactly as
 * vkvg_b200_replay would: same pixels either way.  cmds / args may be reused as soon as the call returns. */
vkvg_public vkvg_status_t vkvg_b200_submit(VkvgContext ctx, const uint32_t *cmds, uint64_t n_cmds, const float *args, uint64_t n_args);
vkvg_public void          vkvg_b200_set_submit_decoder(int mode);                          /* 0: device when possible (default), 1: always the host */
vkvg_public void          vkvg_b200_submit_counts(uint64_t *on_device, uint64_t *on_host); /* streams decoded where, so far (tests, bench) */

/* parity tests: the surface-paint state a draw issued now would use — source x, y, width, height and the inverse matrix
 * (xx yx xy yy x0 y0), i.e. pushConsts.source / pushConsts.matInv of the reference (src/vkvg_context_internal.h:74-81) */
vkvg_public void vkvg_b200_get_source_push(VkvgContext ctx, float out[10]);

/* ---- batches of independent canvases (BASELINE config C5b; SURVEY.md §8e "independent canvases") ----
 * One surface holds `count` canvases of width x height stacked vertically (height a multiple of 16).  Draws recorded after
 * vkvg_b200_set_canvas(ctx, i) go to canvas i: coordinates, gradients and clipping are those of a width x height surface of
 * its own, so every canvas is bit-identical to rendering it alone — but one flush (one set of kernel launches, or one CUDA
 * graph replay) renders all of them.  Rows [i * height, (i + 1) * height) of the surface read-back are canvas i. */
vkvg_public VkvgSurface   vkvg_b200_surface_create_batch(VkvgDevice dev, uint32_t width, uint32_t height, uint32_t count);
vkvg_public vkvg_status_t vkvg_b200_set_canvas(VkvgContext ctx, uint32_t index);

/* ---- execution model knobs ----
 * A flush queues its ~40 kernels without any host round trip (counts that are only known on the device stay there; see
 * vkvg_b200/csrc/dev_util.cuh: vkb_counts) and returns; the next call that needs the result waits for it.  When two
 * consecutive flushes have the same structure (same numbers of path elements, sub-paths and draws, same surface) the
 * second is captured into a CUDA graph and later ones replay it with one launch.
 *   set_graphs(0)        plain launches only
 *   set_stage_timing(1)  (default) flushes that produce statistics record CUDA events between the pipeline stages, which
 *                        needs plain launches; with 0 they may replay the graph and report only ms_total and ms_fine */
vkvg_public void     vkvg_b200_device_set_graphs(VkvgDevice dev, int on);
vkvg_public void     vkvg_b200_device_set_stage_timing(VkvgDevice dev, int on);
vkvg_public uint64_t vkvg_b200_device_graph_replays(VkvgDevice dev);
/* Which fine-pass kernel renders batches that carry no clip state (process-wide; both give identical pixels):
 *   0 (default)  by surface size: fine_warp_k from 16384 tiles (2048 x 2048 pixels) up, fine_k below
 *   1            fine_k: one block of 8 warps per tile (the kernel that also serves clip / save / restore batches)
 *   2            fine_warp_k: one warp per 16x16 tile, covered pixels compacted into a queue
 * Environment: VKVG_B200_FINE=block|warp selects 1 / 2 at start-up. */
vkvg_public void     vkvg_b200_set_fine_kernel(int mode);
vkvg_public int      vkvg_b200_get_fine_kernel(void);

/* ---- SVG parser introspection (parity tests against nanoSVG dumps, tests/test_svg.py) ----
 * Flat dump of a document parsed by vkvg_svg_load (include/vkvg-svg.h), byte-compatible with what oracle/nsvg_dump.c
 * writes for the reference's nanoSVG: "NSVG" f32 width f32 height u32 nshapes, then per shape u32 fillType fillColor
 * strokeType strokeColor, f32 opacity strokeWidth, u32 npaths, then per path u32 npts u32 closed f32 xy[2*npts].
 * Returns the dump size in bytes and writes at most cap of them to out (out may be NULL). */
struct _vkvg_svg_t;
vkvg_public uint64_t vkvg_b200_svg_serialize(struct _vkvg_svg_t *svg, uint8_t *out, uint64_t cap);

#ifdef __cplusplus
}
#endif
#endif
