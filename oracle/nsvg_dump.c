/* TEST INFRASTRUCTURE ONLY — fixture generator.
 * Parses an SVG with the reference's own nanoSVG (src/nsvg/nanosvg.h, included from /root/reference at
 * build time, never copied) exactly as vkvg_svg_load does (src/nsvg/vkvg_nsvg.c:71: units "px", 96 dpi)
 * and writes the flat shape list that vkvg_svg_render (src/nsvg/vkvg_nsvg.c:79-136) walks:
 *   "NSVG" f32 width f32 height u32 nshapes
 *   per shape: u32 fillType u32 fillColor u32 strokeType u32 strokeColor f32 opacity f32 strokeWidth u32 npaths
 *   per path : u32 npts u32 closed f32 pts[2*npts]
 * For gradient paints the colour written is stops[0].color (what the reference driver uses, :113-116).
 * Usage: nsvg_dump in.svg out.bin */
#include <stdio.h>
#include <string.h>
#include <math.h>
#include <stdint.h>
#define NANOSVG_IMPLEMENTATION
#include "nanosvg.h"

static void w32(FILE *f, uint32_t v) { fwrite(&v, 4, 1, f); }
static void wf(FILE *f, float v) { fwrite(&v, 4, 1, f); }
static uint32_t paint_color(NSVGpaint *p) {
    if (p->type == NSVG_PAINT_COLOR) return p->color;
    if (p->type == NSVG_PAINT_LINEAR_GRADIENT || p->type == NSVG_PAINT_RADIAL_GRADIENT) return p->gradient->stops[0].color;
    return 0;
}
int main(int argc, char **argv) {
    if (argc < 3) return 2;
    NSVGimage *img = nsvgParseFromFile(argv[1], "px", 96.0f);
    if (!img) return 1;
    FILE *f = fopen(argv[2], "wb");
    uint32_t n = 0, ncub = 0;
    for (NSVGshape *s = img->shapes; s; s = s->next) n++;
    fwrite("NSVG", 4, 1, f); wf(f, img->width); wf(f, img->height); w32(f, n);
    for (NSVGshape *s = img->shapes; s; s = s->next) {
        uint32_t np = 0;
        for (NSVGpath *p = s->paths; p; p = p->next) np++;
        w32(f, s->fill.type); w32(f, paint_color(&s->fill)); w32(f, s->stroke.type); w32(f, paint_color(&s->stroke));
        wf(f, s->opacity); wf(f, s->strokeWidth); w32(f, np);
        for (NSVGpath *p = s->paths; p; p = p->next) {
            w32(f, p->npts); w32(f, p->closed); fwrite(p->pts, sizeof(float), 2 * p->npts, f);
            ncub += (p->npts - 1) / 3;
        }
    }
    fclose(f);
    fprintf(stderr, "%s: %gx%g shapes=%u cubics=%u\n", argv[1], img->width, img->height, n, ncub);
    nsvgDelete(img);
    return 0;
}
