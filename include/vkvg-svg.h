/* vkvg-svg.h — SVG loading and rendering on top of vkvg.h.
 *
 * Drop-in for the reference's include/vkvg-svg.h:30-96 (implemented there by src/nsvg/vkvg_nsvg.c on top of the
 * vendored nanoSVG).  Same entry points, argument meaning and behaviour: documents are parsed at 96 dpi in "px" units
 * into a list of shapes whose paths are cubic Béziers; vkvg_svg_render replays them through the public vkvg_* calls
 * (even-odd fills, plain strokes, gradients reduced to the colour of their first stop, exactly as the reference driver
 * does).  The parser is this repository's own (vkvg_b200/csrc/svg.cpp); it needs no device. */
#ifndef VKVG_SVG_H
#define VKVG_SVG_H

#include "vkvg.h"

#ifdef __cplusplus
extern "C" {
#endif

/* opaque parsed document (the reference typedefs this to nanoSVG's NSVGimage, include/vkvg-svg.h:27-31) */
typedef struct _vkvg_svg_t *VkvgSvg;

/* reference include/vkvg-svg.h:44: render the file into a new surface of the document's size (width/height are ignored there too) */
vkvg_public VkvgSurface vkvg_surface_create_from_svg(VkvgDevice dev, uint32_t width, uint32_t height, const char *svgFilePath);
/* :56 */
vkvg_public VkvgSurface vkvg_surface_create_from_svg_fragment(VkvgDevice dev, uint32_t width, uint32_t height, char *svgFragment);
/* :66 */
vkvg_public void vkvg_svg_get_dimensions(VkvgSvg svg, uint32_t *width, uint32_t *height);
/* :75 — NULL when the file cannot be read */
vkvg_public VkvgSvg vkvg_svg_load(const char *svgFilePath);
/* :82 */
vkvg_public VkvgSvg vkvg_svg_load_fragment(char *svgFragment);
/* :90 — id == NULL renders every shape, otherwise only shapes whose id matches */
vkvg_public void vkvg_svg_render(VkvgSvg svg, VkvgContext ctx, const char *id);
/* :96 */
vkvg_public void vkvg_svg_destroy(VkvgSvg svg);

#ifdef __cplusplus
}
#endif
#endif
