// Plain-data records shared by the host recorder (vkvg_api.cpp) and the CUDA pipeline (pipeline.cu).
// These are the host->device wire format of one flushed batch.
#pragma once
#include <stdint.h>

// ---- path elements: what the host records instead of flattening on the CPU -------------------------
// (the reference flattens in vkvg_curve_to / vkvg_arc on the calling thread, src/vkvg_context.c:394-566)
// An element is a 32-bit header plus a payload of floats in a separate array:
//   header = type | VKB_EL_CURVED | payload_offset << 3        (payload_offset counts floats)
//   VKB_EL_POINT  payload {x, y}                      -> that point (move_to / line_to target, arc start/end, rectangle corner)
//   VKB_EL_CUBIC  payload {x0,y0, x1,y1, x2,y2, x3,y3, tol} -> the recursion points followed by the end point
//   VKB_EL_ARC    payload {xc, yc, radius, a_first, a_limit, step} -> interior points a_first, a_first+step, ... before a_limit
// A 1M-point polyline is therefore 12 MB on the wire, not 48.
enum : uint32_t {
    VKB_EL_POINT  = 0,
    VKB_EL_CUBIC  = 1,
    VKB_EL_ARC    = 2,
    VKB_EL_NONE   = 3,    // no points (a curve_to the reference skips, written by the device-side stream decoder: its slot stays)
    VKB_EL_TYPE_MASK = 0x3,
    VKB_EL_CURVED = 0x4,  // points of this element belong to a curved segment (PATH_HAS_CURVES_BIT on the segment)
    VKB_EL_PAYLOAD_SHIFT = 3,
};

enum : uint32_t {
    VKB_SP_CLOSED    = 1,  // PATH_CLOSED_BIT
    VKB_SP_CONVEX    = 2,  // PATH_IS_CONVEX_BIT
    VKB_SP_DROP_LAST = 4,  // close_path removed a last point equal to the first (src/vkvg_context.c:363-368)
};
struct vkb_subpath {
    uint32_t first_elem, n_elems;
    uint32_t flags;
    uint32_t pad;
};

// ---- draws ------------------------------------------------------------------------------------------
enum : uint32_t {
    VKB_DRAW_FILL   = 0,  // polygon edges of every sub-path with > 2 points
    VKB_DRAW_STROKE = 1,  // stroke triangles
    VKB_DRAW_PAINT  = 2,  // whole-surface paint
    VKB_DRAW_CLIP   = 3,  // vkvg_clip: polygon edges like a fill, applied to the whole surface (outside the path = clipped out)
    VKB_DRAW_STENCIL = 4, // whole-surface stencil bookkeeping of reset_clip / save / restore (no geometry, like a paint)
};
enum : uint32_t {
    VKB_RULE_EVEN_ODD = 0,  // blend once where winding is odd      (stencil INVERT fan + cover)
    VKB_RULE_NON_ZERO = 1,  // blend once where winding != 0        (libtess triangles, non-overlapping)
    VKB_RULE_COUNT    = 2,  // blend |winding| times                (stroke triangles blended one by one)
    // rules >= VKB_RULE_CLIP_EO write the per-sample stencil byte instead of colour.  Bit 1 of that byte is the
    // reference's STENCIL_CLIP_BIT (set = clipped out), bits 2-7 its six save levels (src/vkvg_device_internal.h:28-30)
    VKB_RULE_CLIP_EO    = 3,  // CLIP |= !(winding odd)
    VKB_RULE_CLIP_NZ    = 4,  // CLIP |= !(winding != 0)
    VKB_RULE_ST_CLEAR   = 5,  // stencil := 0 (vkvg_reset_clip: the reference clears the whole attachment, save bits included)
    VKB_RULE_ST_SAVE    = 6,  // save bit := CLIP     (draw.color = bit | log2(bit) << 8)
    VKB_RULE_ST_RESTORE = 7,  // CLIP := save bit
};
#define VKB_STENCIL_CLIP 0x2u
// blend operator of a colour draw, bits 16-23 of rule_pattern: the three pipelines the reference builds (pipe_OVER, pipe_CLEAR,
// pipe_SUB: src/vkvg_device_internal.c:358-373; _bind_draw_pipeline, src/vkvg_context_internal.c:606-621)
enum : uint32_t { VKB_OP_OVER = 0, VKB_OP_CLEAR = 1, VKB_OP_SUB = 2 };
enum : uint32_t { VKB_PAT_SOLID = 0, VKB_PAT_SURFACE = 1, VKB_PAT_LINEAR = 2, VKB_PAT_RADIAL = 3 };  // vkvg_pattern_type_t values

// A draw is 32 bytes; what rarely changes between draws (CTM, stroke state) lives in side tables that grow only when
// the state differs from the previous entry (100k fills under one CTM upload 3.2 MB of draws, not 11 MB).
struct vkb_draw {
    uint32_t kind;       // VKB_DRAW_*
    uint32_t rule_pattern;  // VKB_RULE_* | VKB_PAT_* << 8 | VKB_OP_* << 16
    uint32_t first_subpath, n_subpaths;
    uint32_t color;      // premultiplied RGBA8, R in byte 0 (CreateRgbaf, src/vkvg_context_internal.h:60-62)
    float    opacity;
    uint32_t gradient;   // index into the batch's gradient table
    uint32_t xform_stroke;  // xform index | stroke-state index << 16  (both tables are capped at 65535 entries per batch)
};
static_assert(sizeof(vkb_draw) == 32, "vkb_draw layout");
struct vkb_xform {       // CTM at draw time: xx yx xy yy x0 y0
    float    mat[6];
    uint32_t band;       // canvas of a batch surface the draw goes to (0 on ordinary surfaces)
    uint32_t pad;
};
struct vkb_stroke {      // stroke parameters (src/vkvg_context.c:830-832, internal.c:245-252)
    float    hw, lhMax, arcStep;
    uint32_t join, cap;
    uint32_t dash_first, dash_count;  // into the batch's dash table
    float    dash_offset;
};

struct vkb_gradient {  // vkvg_gradient_t in scalar block layout, src/vkvg_pattern.h:38-47
    float    colors[16][4];
    float    stops[16];
    float    cp[2][4];
    uint32_t count;
    uint32_t pad[3];
};
static_assert(sizeof(vkb_gradient) == 368, "vkb_gradient layout");

// A surface used as paint (vkvg_set_source_surface / vkvg_pattern_create_for_surface): what the fragment shader gets through
// its push constants and sampler (shaders/vkvg_main.frag:72-82, src/vkvg_context_internal.c:705-773).  draw.gradient indexes
// this table when the pattern is VKB_PAT_SURFACE.
enum : uint32_t { VKB_TEX_NEAREST = 0, VKB_TEX_LINEAR = 1 };
enum : uint32_t { VKB_TEX_BORDER = 0, VKB_TEX_REPEAT = 1, VKB_TEX_MIRROR = 2, VKB_TEX_EDGE = 3 };  // vkvg_extend_t order: NONE, REPEAT, REFLECT, PAD
struct vkb_surfpat {
    uint64_t image;         // device address of the source surface's resolved premultiplied RGBA8 pixels
    uint32_t width, height;
    uint32_t filter_extend; // VKB_TEX_* filter | address mode << 8
    float    sx, sy;        // pushConsts.source.xy: the offset given to vkvg_set_source_surface (device pixels)
    float    minv[6];       // pushConsts.matInv (inverse CTM, times the pattern matrix): xx yx xy yy x0 y0
    uint32_t pad;
};
static_assert(sizeof(vkb_surfpat) == 56, "vkb_surfpat layout");

// what the fine pass reads per draw (16 B)
struct vkb_paint {
    uint32_t rule_pattern;  // rule | pattern << 8
    uint32_t color;
    float    opacity;
    uint32_t gradient;
};

// device-space edge, 24.8 fixed point window coordinates (16 B per edge)
struct vkb_edge {
    int32_t x0, y0, x1, y1;
};

#define VKB_TILE 16            // pixels
#define VKB_TILE_FX 4096       // VKB_TILE << 8
#define VKB_MAX_DASHES 32
