/* A plain C program against include/vkvg.h + libvkvg_b200.so, in the shape of the reference's tests/offscreen.c (device with
 * 4 samples, surface, context, draw, destroy the context - which flushes - write the PNG): the only part of the drop-in boundary that
 * Python cannot exercise is a compiled C client, so this is one.  tests/test_c_client.py builds it with gcc, runs it on the GPU box and
 * compares the PNG it writes with the oracle's rendering of the same calls. */
#include <stdio.h>
#include "vkvg.h"

int main(int argc, char *argv[]) {
    const char *out = argc > 1 ? argv[1] : "offscreen.png";
    vkvg_device_create_info_t info = {VK_SAMPLE_COUNT_4_BIT, false};
    VkvgDevice  dev  = vkvg_device_create(&info);
    if (vkvg_device_status(dev)) { fprintf(stderr, "device: %s\n", vkvg_status_to_string(vkvg_device_status(dev))); return 2; }
    VkvgSurface surf = vkvg_surface_create(dev, 256, 192);
    VkvgContext ctx  = vkvg_create(surf);

    vkvg_clear(ctx);
    vkvg_rectangle(ctx, 10, 10, 120, 90);
    vkvg_set_source_rgb(ctx, 1, 0, 0);
    vkvg_fill(ctx);

    vkvg_set_fill_rule(ctx, VKVG_FILL_RULE_EVEN_ODD);
    vkvg_move_to(ctx, 60.5f, 40.25f);
    vkvg_curve_to(ctx, 200, 10, 240, 180, 100.75f, 150);
    vkvg_line_to(ctx, 180, 60);
    vkvg_close_path(ctx);
    VkvgPattern pat = vkvg_pattern_create_linear(40, 20, 220, 170);
    vkvg_pattern_add_color_stop(pat, 0.0f, 0.1f, 0.3f, 0.9f, 1.0f);
    vkvg_pattern_add_color_stop(pat, 1.0f, 0.9f, 0.8f, 0.1f, 0.5f);
    vkvg_set_source(ctx, pat);
    vkvg_pattern_destroy(pat);
    vkvg_fill_preserve(ctx);

    vkvg_set_source_rgba(ctx, 0.0f, 0.4f, 0.1f, 0.8f);
    vkvg_set_line_width(ctx, 5.0f);
    vkvg_set_line_join(ctx, VKVG_LINE_JOIN_ROUND);
    const float dashes[2] = {9.0f, 4.0f};
    vkvg_set_dash(ctx, dashes, 2, 1.5f);
    vkvg_stroke(ctx);

    vkvg_destroy(ctx);
    vkvg_status_t st = vkvg_surface_write_to_png(surf, out);
    vkvg_surface_destroy(surf);
    vkvg_device_destroy(dev);
    return st == VKVG_STATUS_SUCCESS ? 0 : 1;
}
