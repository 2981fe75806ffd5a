// Plain-data records shared by the device-side command-stream decoder (decode.cu), the orchestrator (pipeline.cu) and the C API layer.
#pragma once
#include <stdint.h>
#include "vkb_types.h"

struct vkb_decode_init {  // the context's state when the stream starts
    float    mat[6];
    uint32_t band, color, rule, cap, join, bop, dash_count;
    float    lw, miter, opacity, dash_offset;
    float    dashes[VKB_MAX_DASHES];
    uint32_t pattern;        // VKB_PAT_SOLID, or VKB_PAT_LINEAR / VKB_PAT_RADIAL with the gradient below (as _update_cur_pattern left it)
    vkb_gradient grad;
};
struct vkb_decode_census {  // read back once per stream: what the host sizes the pipeline from, and the state the context is left in
    uint32_t n_elems, n_data, n_subpaths, n_draws, n_curves, n_grads, n_dash_floats, n_xforms, n_strokes;
    uint32_t n_fjobs, n_sjobs, n_sdraws, any_dash, nz_any;
    uint32_t irregular;        // non-zero: the stream is outside what the device decodes; nothing was committed
    int32_t  last_setter[9];   // command index of the last setter of: source, fill rule, line width, cap, join, miter limit, dash, opacity, gradient (or -1)
    uint32_t last_setter_arg[9];  // where its arguments start in `args`
    float    final_mat[6];
    uint32_t final_band;
    uint32_t max_sp_elems;     // elements of the longest sub-path when some hold more than 1024 (else 0): the flush then reduces their boxes block by block
    uint32_t pad;
};
