// SVG front end of the drop-in: vkvg-svg.h (reference include/vkvg-svg.h, src/nsvg/vkvg_nsvg.c).
//
// The reference delegates parsing to nanoSVG (src/nsvg/nanosvg.h, third party, vendored there) and then walks the
// resulting shape list with vkvg calls (vkvg_svg_render, src/nsvg/vkvg_nsvg.c:79-136).  This file is an independent
// C++ parser that produces the SAME shape list for the same document: every path is reduced to cubic Béziers whose
// control points go through the same float operations in the same order (number scanning as integer part +
// fraction / 10^digits in double, lines as cubics with handles at one third, quadratic -> cubic by the 2/3 rule,
// elliptical arcs split into <= 90 degree cubic pieces, group transforms pre-multiplied and applied per point, then the
// viewBox mapping), because the tiger frame (BASELINE config C1) has to match the reference bit for bit downstream.
// tests/test_svg.py compares the serialised shape list against dumps made by nanoSVG itself (oracle/nsvg_dump.c).
// Host-only code: parsing needs no device; rendering goes through the public vkvg_* entry points.
#include "../../include/vkvg.h"
#include "../../include/vkvg-svg.h"
#include "../../include/vkvg_b200.h"
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <vector>

namespace {

const float kPi      = 3.14159265358979323846264338327f;
const float kKappa90 = 0.5522847493f;  // handle length of a 90 degree arc, relative to the radius

enum PaintType { PAINT_NONE = 0, PAINT_COLOR = 1, PAINT_LINEAR = 2, PAINT_RADIAL = 3 };
enum Units { U_USER, U_PX, U_PT, U_PC, U_MM, U_CM, U_IN, U_PERCENT, U_EM, U_EX };
enum Align { ALIGN_MIN = 0, ALIGN_MID = 1, ALIGN_MAX = 2, ALIGN_NONE = 0, ALIGN_MEET = 1, ALIGN_SLICE = 2 };

struct Coord { float value; int units; };

struct SvgPath {
    std::vector<float> pts;  // x,y pairs: 1 + 3k points
    bool               closed;
    float              bounds[4];
};
struct SvgShape {
    char                 id[64];
    int                  fillType, strokeType;
    uint32_t             fillColor, strokeColor;  // for gradients: colour of the first stop (all the reference driver uses)
    float                opacity, strokeWidth;
    float                bounds[4];
    std::vector<SvgPath> paths;  // in the order the reference walks them (last parsed sub-path first)
};

struct GradientData {
    char     id[64], ref[64];
    int      type;
    std::vector<std::pair<float, uint32_t>> stops;  // (offset, colour), kept sorted by offset
};

struct Style {  // inherited presentation state of the element being parsed
    char     id[64];
    float    xform[6];
    uint32_t fillColor, strokeColor;
    float    opacity, fillOpacity, strokeOpacity;
    char     fillGradient[64], strokeGradient[64];
    float    strokeWidth, fontSize;
    uint32_t stopColor;
    float    stopOpacity, stopOffset;
    char     hasFill, hasStroke, visible;
};

inline bool is_space(char c) { return c && strchr(" \t\n\v\f\r", c) != nullptr; }
inline bool is_digit(char c) { return c >= '0' && c <= '9'; }
inline bool is_numch(char c) { return c && strchr("0123456789+-.eE", c) != nullptr; }
inline uint32_t rgb(unsigned r, unsigned g, unsigned b) { return r | (g << 8) | (b << 16); }

// ---- 2x3 affine helpers (column layout a b c d e f: x' = a x + c y + e, y' = b x + d y + f) ----
void xf_identity(float *t) { t[0] = 1; t[1] = 0; t[2] = 0; t[3] = 1; t[4] = 0; t[5] = 0; }
void xf_translate(float *t, float x, float y) { xf_identity(t); t[4] = x; t[5] = y; }
void xf_scale(float *t, float x, float y) { xf_identity(t); t[0] = x; t[3] = y; }
void xf_rotate(float *t, float a) {
    float cs = cosf(a), sn = sinf(a);
    t[0] = cs; t[1] = sn; t[2] = -sn; t[3] = cs; t[4] = 0; t[5] = 0;
}
void xf_mul(float *t, const float *s) {  // t = t * s (t applied first)
    float t0 = t[0] * s[0] + t[1] * s[2];
    float t2 = t[2] * s[0] + t[3] * s[2];
    float t4 = t[4] * s[0] + t[5] * s[2] + s[4];
    t[1]     = t[0] * s[1] + t[1] * s[3];
    t[3]     = t[2] * s[1] + t[3] * s[3];
    t[5]     = t[4] * s[1] + t[5] * s[3] + s[5];
    t[0] = t0; t[2] = t2; t[4] = t4;
}
void xf_premul(float *t, const float *s) {  // t = s * t
    float s2[6];
    memcpy(s2, s, sizeof s2);
    xf_mul(s2, t);
    memcpy(t, s2, sizeof s2);
}
inline void xf_point(float &ox, float &oy, float x, float y, const float *t) {
    ox = x * t[0] + y * t[2] + t[4];
    oy = x * t[1] + y * t[3] + t[5];
}
inline void xf_vec(float &ox, float &oy, float x, float y, const float *t) {
    ox = x * t[0] + y * t[2];
    oy = x * t[1] + y * t[3];
}

// ---- numbers ----
// decimal text -> double the way nanoSVG does it (locale free): integer part, then fraction digits / 10^n, then 10^exp
double scan_double(const char *s) {
    const char *cur = s;
    char       *end = nullptr;
    double      res = 0.0, sign = 1.0;
    bool        has_int = false, has_frac = false;
    if (*cur == '+') cur++;
    else if (*cur == '-') { sign = -1; cur++; }
    if (is_digit(*cur)) {
        long long ip = (long long)(double)strtoll(cur, &end, 10);
        if (cur != end) { res = (double)ip; has_int = true; cur = end; }
    }
    if (*cur == '.') {
        cur++;
        if (is_digit(*cur)) {
            long long fp = strtoll(cur, &end, 10);
            if (cur != end) { res += (double)fp / pow(10.0, (double)(end - cur)); has_frac = true; cur = end; }
        }
    }
    if (!has_int && !has_frac) return 0.0;
    if (*cur == 'e' || *cur == 'E') {
        cur++;
        long ep = strtol(cur, &end, 10);
        if (cur != end) res *= pow(10.0, (double)ep);
    }
    return res * sign;
}
// copies one numeric token ([sign] digits [. digits] [e [sign] digits]; a leading 0 is a token digit of its own; an 'e'
// that starts "em"/"ex" is a unit, not an exponent) into it[size]; returns the first character after it
const char *scan_number_token(const char *s, char *it, int size) {
    const int last = size - 1;
    int       i    = 0;
    auto put = [&](char c) { if (i < last) it[i++] = c; };
    if (*s == '-' || *s == '+') put(*s++);
    if (*s == '0') put(*s++);
    else while (is_digit(*s)) put(*s++);
    if (*s == '.') {
        put(*s++);
        while (is_digit(*s)) put(*s++);
    }
    if ((*s == 'e' || *s == 'E') && (s[1] != 'm' && s[1] != 'x')) {
        put(*s++);
        if (*s == '-' || *s == '+') put(*s++);
        while (is_digit(*s)) put(*s++);
    }
    it[i] = '\0';
    return s;
}
const char *next_path_item(const char *s, char *it) {
    it[0] = '\0';
    while (is_space(*s) || *s == ',') s++;
    if (!*s) return s;
    if (*s == '-' || *s == '+' || *s == '.' || is_digit(*s)) return scan_number_token(s, it, 64);
    it[0] = *s++;
    it[1] = '\0';
    return s;
}
const char *next_arc_flag(const char *s, char *it) {
    it[0] = '\0';
    while (is_space(*s) || *s == ',') s++;
    if (*s == '0' || *s == '1') { it[0] = *s++; it[1] = '\0'; }
    return s;
}
int units_of(const char *u) {
    if (u[0] == 'p' && u[1] == 'x') return U_PX;
    if (u[0] == 'p' && u[1] == 't') return U_PT;
    if (u[0] == 'p' && u[1] == 'c') return U_PC;
    if (u[0] == 'm' && u[1] == 'm') return U_MM;
    if (u[0] == 'c' && u[1] == 'm') return U_CM;
    if (u[0] == 'i' && u[1] == 'n') return U_IN;
    if (u[0] == '%') return U_PERCENT;
    if (u[0] == 'e' && u[1] == 'm') return U_EM;
    if (u[0] == 'e' && u[1] == 'x') return U_EX;
    return U_USER;
}
Coord scan_coord(const char *str) {
    char  buf[64];
    Coord c;
    c.units = units_of(scan_number_token(str, buf, 64));
    c.value = (float)scan_double(buf);
    return c;
}

// ---- colours ----
uint32_t color_hex(const char *str) {
    unsigned c = 0;
    int      n = 0;
    str++;
    while (str[n] && !is_space(str[n])) n++;
    if (n == 6) sscanf(str, "%x", &c);
    else if (n == 3) {
        sscanf(str, "%x", &c);
        c = (c & 0xf) | ((c & 0xf0) << 4) | ((c & 0xf00) << 8);
        c |= c << 4;
    }
    return rgb((c >> 16) & 0xff, (c >> 8) & 0xff, c & 0xff);
}
uint32_t color_rgb_fn(const char *str) {
    int  r = -1, g = -1, b = -1;
    char s1[32] = "", s2[32] = "";
    sscanf(str + 4, "%d%[%%, \t]%d%[%%, \t]%d", &r, s1, &g, s2, &b);
    if (strchr(s1, '%')) return rgb((unsigned)((r * 255) / 100), (unsigned)((g * 255) / 100), (unsigned)((b * 255) / 100));
    return rgb((unsigned)r, (unsigned)g, (unsigned)b);
}
uint32_t color_named(const char *str) {  // the ten keywords a default nanoSVG build knows; anything else is mid grey
    static const struct { const char *name; uint32_t c; } table[] = {
        {"red", 0x0000ffu},  {"green", 0x008000u}, {"blue", 0xff0000u},    {"yellow", 0x00ffffu}, {"cyan", 0xffff00u},
        {"magenta", 0xff00ffu}, {"black", 0u},     {"grey", 0x808080u},    {"gray", 0x808080u},   {"white", 0xffffffu}};
    for (auto &e : table)
        if (!strcmp(e.name, str)) return e.c;
    return 0x808080u;
}
uint32_t parse_color(const char *str) {
    while (*str == ' ') ++str;
    size_t len = strlen(str);
    if (len >= 1 && *str == '#') return color_hex(str);
    if (len >= 4 && !strncmp(str, "rgb(", 4)) return color_rgb_fn(str);
    return color_named(str);
}
float parse_opacity(const char *str) {
    float v = (float)scan_double(str);
    if (v < 0.0f) v = 0.0f;
    if (v > 1.0f) v = 1.0f;
    return v;
}

// ---- cubic bounds (only used to size a document that declares neither size nor viewBox) ----
double bez(double t, double p0, double p1, double p2, double p3) {
    double it = 1.0 - t;
    return it * it * it * p0 + 3.0 * it * it * t * p1 + 3.0 * it * t * t * p2 + t * t * t * p3;
}
inline float minf(float a, float b) { return a < b ? a : b; }
inline float maxf(float a, float b) { return a > b ? a : b; }
void curve_bounds(float *bounds, const float *cv) {
    const float *v0 = cv, *v1 = cv + 2, *v2 = cv + 4, *v3 = cv + 6;
    const double EPS = 1e-12;
    bounds[0] = minf(v0[0], v3[0]); bounds[1] = minf(v0[1], v3[1]);
    bounds[2] = maxf(v0[0], v3[0]); bounds[3] = maxf(v0[1], v3[1]);
    auto in = [&](const float *p) { return p[0] >= bounds[0] && p[0] <= bounds[2] && p[1] >= bounds[1] && p[1] <= bounds[3]; };
    if (in(v1) && in(v2)) return;
    for (int i = 0; i < 2; i++) {
        double a = -3.0 * v0[i] + 9.0 * v1[i] - 9.0 * v2[i] + 3.0 * v3[i];
        double b = 6.0 * v0[i] - 12.0 * v1[i] + 6.0 * v2[i];
        double c = 3.0 * v1[i] - 3.0 * v0[i];
        double roots[2];
        int    count = 0;
        if (fabs(a) < EPS) {
            if (fabs(b) > EPS) {
                double t = -c / b;
                if (t > EPS && t < 1.0 - EPS) roots[count++] = t;
            }
        } else {
            double disc = b * b - 4.0 * c * a;
            if (disc > EPS) {
                double t = (-b + sqrt(disc)) / (2.0 * a);
                if (t > EPS && t < 1.0 - EPS) roots[count++] = t;
                t = (-b - sqrt(disc)) / (2.0 * a);
                if (t > EPS && t < 1.0 - EPS) roots[count++] = t;
            }
        }
        for (int j = 0; j < count; j++) {
            double v = bez(roots[j], v0[i], v1[i], v2[i], v3[i]);
            bounds[i]     = minf(bounds[i], (float)v);
            bounds[2 + i] = maxf(bounds[2 + i], (float)v);
        }
    }
}

// ====================================================================================================
struct Parser {
    float dpi = 96.0f;
    float width = 0, height = 0;
    float viewMinx = 0, viewMiny = 0, viewWidth = 0, viewHeight = 0;
    int   alignX = 0, alignY = 0, alignType = 0;
    bool  inDefs = false;
    std::vector<Style>        stack;  // at most 128 levels, deeper pushes are ignored (and their pops still pop)
    int                       head = 0;
    std::vector<float>        pts;    // path under construction (user space of the element)
    std::vector<SvgPath>      plist;  // finished sub-paths of the current element, most recent first
    std::vector<SvgShape>     shapes;
    std::vector<GradientData> grads;  // most recent last; lookups scan from the back

    Parser() {
        stack.resize(128);
        Style &a = stack[0];
        memset(&a, 0, sizeof a);
        xf_identity(a.xform);
        a.opacity = a.fillOpacity = a.strokeOpacity = a.stopOpacity = 1;
        a.strokeWidth = 1;
        a.hasFill = 1; a.visible = 1;
    }
    Style &cur() { return stack[head]; }
    void   push() { if (head < 127) { head++; stack[head] = stack[head - 1]; } }
    void   pop() { if (head > 0) head--; }

    float actual_length() const { return sqrtf(viewWidth * viewWidth + viewHeight * viewHeight) / sqrtf(2.0f); }
    float to_pixels(Coord c, float orig, float length) {
        switch (c.units) {
        case U_PT: return c.value / 72.0f * dpi;
        case U_PC: return c.value / 6.0f * dpi;
        case U_MM: return c.value / 25.4f * dpi;
        case U_CM: return c.value / 2.54f * dpi;
        case U_IN: return c.value * dpi;
        case U_EM: return c.value * cur().fontSize;
        case U_EX: return c.value * cur().fontSize * 0.52f;
        case U_PERCENT: return orig + c.value / 100.0f * length;
        default: return c.value;
        }
    }
    float coord(const char *str, float orig, float length) { return to_pixels(scan_coord(str), orig, length); }

    // ---- path construction ----
    void add_point(float x, float y) { pts.push_back(x); pts.push_back(y); }
    size_t npts() const { return pts.size() / 2; }
    void move_to(float x, float y) {
        if (npts() > 0) { pts[pts.size() - 2] = x; pts[pts.size() - 1] = y; }
        else add_point(x, y);
    }
    void line_to(float x, float y) {
        if (!npts()) return;
        float px = pts[pts.size() - 2], py = pts[pts.size() - 1];
        float dx = x - px, dy = y - py;
        add_point(px + dx / 3.0f, py + dy / 3.0f);
        add_point(x - dx / 3.0f, y - dy / 3.0f);
        add_point(x, y);
    }
    void cubic_to(float x1, float y1, float x2, float y2, float x, float y) { add_point(x1, y1); add_point(x2, y2); add_point(x, y); }

    void commit_path(bool closed) {
        if (npts() < 4) return;
        if (closed) line_to(pts[0], pts[1]);
        SvgPath path;
        path.closed = closed;
        path.pts.resize(pts.size());
        const float *t = cur().xform;
        for (size_t i = 0; i < npts(); i++) xf_point(path.pts[2 * i], path.pts[2 * i + 1], pts[2 * i], pts[2 * i + 1], t);
        for (size_t i = 0; i + 1 < npts(); i += 3) {
            float b[4];
            curve_bounds(b, &path.pts[2 * i]);
            if (i == 0) memcpy(path.bounds, b, sizeof b);
            else {
                path.bounds[0] = minf(path.bounds[0], b[0]); path.bounds[1] = minf(path.bounds[1], b[1]);
                path.bounds[2] = maxf(path.bounds[2], b[2]); path.bounds[3] = maxf(path.bounds[3], b[3]);
            }
        }
        plist.insert(plist.begin(), std::move(path));
    }

    const GradientData *find_gradient(const char *id) const {
        for (size_t i = grads.size(); i-- > 0;)
            if (!strcmp(grads[i].id, id)) return &grads[i];
        return nullptr;
    }
    // paint type + first stop colour of url(#id); stops may come from an xlink:href chain
    int resolve_gradient(const char *id, uint32_t *color) const {
        const GradientData *data = find_gradient(id);
        if (!data) return PAINT_NONE;
        const GradientData *ref = data;
        for (int hops = 0; ref && hops < 256; hops++) {
            if (!ref->stops.empty()) { *color = ref->stops[0].second; return data->type; }
            ref = find_gradient(ref->ref);
        }
        return PAINT_NONE;
    }
    void commit_shape() {
        if (plist.empty()) return;
        Style   &a = cur();
        SvgShape sh;
        memcpy(sh.id, a.id, sizeof sh.id);
        float scale    = (sqrtf(a.xform[0] * a.xform[0] + a.xform[2] * a.xform[2]) + sqrtf(a.xform[1] * a.xform[1] + a.xform[3] * a.xform[3])) * 0.5f;
        sh.strokeWidth = a.strokeWidth * scale;
        sh.opacity     = a.opacity;
        sh.paths.swap(plist);
        memcpy(sh.bounds, sh.paths[0].bounds, sizeof sh.bounds);
        for (size_t i = 1; i < sh.paths.size(); i++) {
            sh.bounds[0] = minf(sh.bounds[0], sh.paths[i].bounds[0]); sh.bounds[1] = minf(sh.bounds[1], sh.paths[i].bounds[1]);
            sh.bounds[2] = maxf(sh.bounds[2], sh.paths[i].bounds[2]); sh.bounds[3] = maxf(sh.bounds[3], sh.paths[i].bounds[3]);
        }
        sh.fillType = PAINT_NONE; sh.fillColor = 0;
        if (a.hasFill == 1) {
            sh.fillType  = PAINT_COLOR;
            sh.fillColor = a.fillColor | ((uint32_t)(a.fillOpacity * 255) << 24);
        } else if (a.hasFill == 2)
            sh.fillType = resolve_gradient(a.fillGradient, &sh.fillColor);
        sh.strokeType = PAINT_NONE; sh.strokeColor = 0;
        if (a.hasStroke == 1) {
            sh.strokeType  = PAINT_COLOR;
            sh.strokeColor = a.strokeColor | ((uint32_t)(a.strokeOpacity * 255) << 24);
        } else if (a.hasStroke == 2)
            sh.strokeType = resolve_gradient(a.strokeGradient, &sh.strokeColor);
        shapes.push_back(std::move(sh));
    }

    // ---- transforms ----
    static int transform_args(const char *str, float *args, int max_na, int *na) {
        *na = 0;
        const char *ptr = str;
        while (*ptr && *ptr != '(') ++ptr;
        if (!*ptr) return 1;
        const char *end = ptr;
        while (*end && *end != ')') ++end;
        if (!*end) return 1;
        char it[64];
        while (ptr < end) {
            if (*ptr == '-' || *ptr == '+' || *ptr == '.' || is_digit(*ptr)) {
                if (*na >= max_na) return 0;
                ptr           = scan_number_token(ptr, it, 64);
                args[(*na)++] = (float)scan_double(it);
            } else ++ptr;
        }
        return (int)(end - str);
    }
    static void parse_transform(float *xform, const char *str) {
        float t[6];
        xf_identity(xform);
        xf_identity(t);
        while (*str) {
            float a[6] = {0, 0, 0, 0, 0, 0};
            int   na = 0, len;
            if (!strncmp(str, "matrix", 6)) {
                len = transform_args(str, a, 6, &na);
                if (na == 6) memcpy(t, a, sizeof t);
            } else if (!strncmp(str, "translate", 9)) {
                len = transform_args(str, a, 2, &na);
                if (na == 1) a[1] = 0.0f;
                xf_translate(t, a[0], a[1]);
            } else if (!strncmp(str, "scale", 5)) {
                len = transform_args(str, a, 2, &na);
                if (na == 1) a[1] = a[0];
                xf_scale(t, a[0], a[1]);
            } else if (!strncmp(str, "rotate", 6)) {
                len = transform_args(str, a, 3, &na);
                if (na == 1) a[1] = a[2] = 0.0f;
                float m[6], u[6];
                xf_identity(m);
                if (na > 1) { xf_translate(u, -a[1], -a[2]); xf_mul(m, u); }
                xf_rotate(u, a[0] / 180.0f * kPi);
                xf_mul(m, u);
                if (na > 1) { xf_translate(u, a[1], a[2]); xf_mul(m, u); }
                memcpy(t, m, sizeof t);
            } else if (!strncmp(str, "skewX", 5)) {
                len = transform_args(str, a, 1, &na);
                xf_identity(t);
                t[2] = tanf(a[0] / 180.0f * kPi);
            } else if (!strncmp(str, "skewY", 5)) {
                len = transform_args(str, a, 1, &na);
                xf_identity(t);
                t[1] = tanf(a[0] / 180.0f * kPi);
            } else { ++str; continue; }
            str += len > 0 ? len : 1;  // (a transform with too many arguments makes the reference parser spin forever)
            xf_premul(xform, t);
        }
    }

    // ---- presentation attributes; returns whether the name was one of them ----
    static void parse_url(char *id, const char *str) {
        int i = 0;
        str += 4;
        if (*str == '#') str++;
        while (i < 63 && *str && *str != ')') id[i++] = *str++;
        id[i] = '\0';
    }
    bool attr(const char *name, const char *value) {
        Style &a = cur();
        if (!strcmp(name, "style")) style(value);
        else if (!strcmp(name, "display")) { if (!strcmp(value, "none")) a.visible = 0; }
        else if (!strcmp(name, "fill")) {
            if (!strcmp(value, "none")) a.hasFill = 0;
            else if (!strncmp(value, "url(", 4)) { a.hasFill = 2; parse_url(a.fillGradient, value); }
            else { a.hasFill = 1; a.fillColor = parse_color(value); }
        } else if (!strcmp(name, "opacity")) a.opacity = parse_opacity(value);
        else if (!strcmp(name, "fill-opacity")) a.fillOpacity = parse_opacity(value);
        else if (!strcmp(name, "stroke")) {
            if (!strcmp(value, "none")) a.hasStroke = 0;
            else if (!strncmp(value, "url(", 4)) { a.hasStroke = 2; parse_url(a.strokeGradient, value); }
            else { a.hasStroke = 1; a.strokeColor = parse_color(value); }
        } else if (!strcmp(name, "stroke-width")) a.strokeWidth = coord(value, 0.0f, actual_length());
        else if (!strcmp(name, "stroke-opacity")) a.strokeOpacity = parse_opacity(value);
        else if (!strcmp(name, "font-size")) a.fontSize = coord(value, 0.0f, actual_length());
        else if (!strcmp(name, "transform")) {
            float x[6];
            parse_transform(x, value);
            xf_premul(a.xform, x);
        } else if (!strcmp(name, "stop-color")) a.stopColor = parse_color(value);
        else if (!strcmp(name, "stop-opacity")) a.stopOpacity = parse_opacity(value);
        else if (!strcmp(name, "offset")) a.stopOffset = coord(value, 0.0f, 1.0f);
        else if (!strcmp(name, "id")) { strncpy(a.id, value, 63); a.id[63] = '\0'; }
        else if (!strcmp(name, "stroke-dasharray") || !strcmp(name, "stroke-dashoffset") || !strcmp(name, "stroke-linecap") ||
                 !strcmp(name, "stroke-linejoin") || !strcmp(name, "stroke-miterlimit") || !strcmp(name, "fill-rule")) {
            // recognised, but the reference driver never reads them (src/nsvg/vkvg_nsvg.c:84-131)
        } else return false;
        return true;
    }
    void style(const char *str) {  // "name: value; name: value"
        while (*str) {
            while (is_space(*str)) ++str;
            const char *start = str;
            while (*str && *str != ';') ++str;
            const char *end = str;
            while (end > start && (*end == ';' || is_space(*end) || !*end)) --end;
            ++end;
            // split at the first ':'
            const char *p = start;
            while (p < end && *p != ':') ++p;
            const char *val = p;
            while (p > start && (*p == ':' || is_space(*p))) --p;
            ++p;
            std::string name(start, (size_t)(p - start > 511 ? 511 : p - start));
            while (val < end && (*val == ':' || is_space(*val))) ++val;
            std::string value(val, (size_t)(end - val > 511 ? 511 : (end > val ? end - val : 0)));
            attr(name.c_str(), value.c_str());
            if (*str) ++str;
        }
    }
    typedef std::vector<std::pair<const char *, const char *>> AttrList;
    void attribs(const AttrList &al) {
        for (auto &kv : al) {
            if (!strcmp(kv.first, "style")) style(kv.second);
            else attr(kv.first, kv.second);
        }
    }

    // ---- elements ----
    static int args_per_cmd(char c) {
        switch (c) {
        case 'v': case 'V': case 'h': case 'H': return 1;
        case 'm': case 'M': case 'l': case 'L': case 't': case 'T': return 2;
        case 'q': case 'Q': case 's': case 'S': return 4;
        case 'c': case 'C': return 6;
        case 'a': case 'A': return 7;
        }
        return 0;
    }
    static float vecang(float ux, float uy, float vx, float vy) {
        float r = (ux * vx + uy * vy) / (sqrtf(ux * ux + uy * uy) * sqrtf(vx * vx + vy * vy));
        if (r < -1.0f) r = -1.0f;
        if (r > 1.0f) r = 1.0f;
        return ((ux * vy < uy * vx) ? -1.0f : 1.0f) * acosf(r);
    }
    void arc_to(float &cpx, float &cpy, const float *args, bool rel) {  // SVG implementation notes F.6, pieces of <= 90 degrees
        float rx = fabsf(args[0]), ry = fabsf(args[1]);
        float rotx = args[2] / 180.0f * kPi;
        int   fa = fabsf(args[3]) > 1e-6 ? 1 : 0, fs = fabsf(args[4]) > 1e-6 ? 1 : 0;
        float x1 = cpx, y1 = cpy, x2, y2;
        if (rel) { x2 = cpx + args[5]; y2 = cpy + args[6]; }
        else { x2 = args[5]; y2 = args[6]; }
        float dx = x1 - x2, dy = y1 - y2;
        float d  = sqrtf(dx * dx + dy * dy);
        if (d < 1e-6f || rx < 1e-6f || ry < 1e-6f) {
            line_to(x2, y2);
            cpx = x2; cpy = y2;
            return;
        }
        float sinrx = sinf(rotx), cosrx = cosf(rotx);
        float x1p = cosrx * dx / 2.0f + sinrx * dy / 2.0f;
        float y1p = -sinrx * dx / 2.0f + cosrx * dy / 2.0f;
        d = (x1p * x1p) / (rx * rx) + (y1p * y1p) / (ry * ry);
        if (d > 1) { d = sqrtf(d); rx *= d; ry *= d; }
        float s  = 0.0f;
        float sa = (rx * rx) * (ry * ry) - (rx * rx) * (y1p * y1p) - (ry * ry) * (x1p * x1p);
        float sb = (rx * rx) * (y1p * y1p) + (ry * ry) * (x1p * x1p);
        if (sa < 0.0f) sa = 0.0f;
        if (sb > 0.0f) s = sqrtf(sa / sb);
        if (fa == fs) s = -s;
        float cxp = s * rx * y1p / ry;
        float cyp = s * -ry * x1p / rx;
        float cx = (x1 + x2) / 2.0f + cosrx * cxp - sinrx * cyp;
        float cy = (y1 + y2) / 2.0f + sinrx * cxp + cosrx * cyp;
        float ux = (x1p - cxp) / rx, uy = (y1p - cyp) / ry;
        float vx = (-x1p - cxp) / rx, vy = (-y1p - cyp) / ry;
        float a1 = vecang(1.0f, 0.0f, ux, uy);
        float da = vecang(ux, uy, vx, vy);
        if (fs == 0 && da > 0) da -= 2 * kPi;
        else if (fs == 1 && da < 0) da += 2 * kPi;
        float t[6] = {cosrx, sinrx, -sinrx, cosrx, cx, cy};
        int   ndivs = (int)(fabsf(da) / (kPi * 0.5f) + 1.0f);
        float hda   = (da / (float)ndivs) / 2.0f;
        float kappa = fabsf(4.0f / 3.0f * (1.0f - cosf(hda)) / sinf(hda));
        if (da < 0.0f) kappa = -kappa;
        float px = 0, py = 0, ptanx = 0, ptany = 0;
        for (int i = 0; i <= ndivs; i++) {
            float a = a1 + da * ((float)i / (float)ndivs);
            dx = cosf(a); dy = sinf(a);
            float x, y, tanx, tany;
            xf_point(x, y, dx * rx, dy * ry, t);
            xf_vec(tanx, tany, -dy * rx * kappa, dx * ry * kappa, t);
            if (i > 0) cubic_to(px + ptanx, py + ptany, x - tanx, y - tany, x, y);
            px = x; py = y; ptanx = tanx; ptany = tany;
        }
        cpx = x2; cpy = y2;
    }
    void el_path(const AttrList &al) {
        const char *s = nullptr;
        for (auto &kv : al) {
            if (!strcmp(kv.first, "d")) s = kv.second;
            else { AttrList one(1, kv); attribs(one); }
        }
        if (s) {
            pts.clear();
            float cpx = 0, cpy = 0, cpx2 = 0, cpy2 = 0, args[10];
            bool  closed = false;
            int   nargs = 0, rargs = 0;
            char  cmd = '\0', item[64];
            while (*s) {
                item[0] = '\0';
                if ((cmd == 'A' || cmd == 'a') && (nargs == 3 || nargs == 4)) s = next_arc_flag(s, item);
                if (!*item) s = next_path_item(s, item);
                if (!*item) break;
                if (is_numch(item[0])) {
                    if (nargs < 10) args[nargs++] = (float)scan_double(item);
                    if (nargs < rargs) continue;
                    const bool rel = cmd >= 'a' && cmd <= 'z';
                    switch (cmd) {
                    case 'm': case 'M':
                        if (rel) { cpx += args[0]; cpy += args[1]; } else { cpx = args[0]; cpy = args[1]; }
                        move_to(cpx, cpy);
                        cmd   = rel ? 'l' : 'L';  // further pairs are implicit line-tos
                        rargs = args_per_cmd(cmd);
                        cpx2 = cpx; cpy2 = cpy;
                        break;
                    case 'l': case 'L':
                        if (rel) { cpx += args[0]; cpy += args[1]; } else { cpx = args[0]; cpy = args[1]; }
                        line_to(cpx, cpy);
                        cpx2 = cpx; cpy2 = cpy;
                        break;
                    case 'h': case 'H':
                        if (rel) cpx += args[0]; else cpx = args[0];
                        line_to(cpx, cpy);
                        cpx2 = cpx; cpy2 = cpy;
                        break;
                    case 'v': case 'V':
                        if (rel) cpy += args[0]; else cpy = args[0];
                        line_to(cpx, cpy);
                        cpx2 = cpx; cpy2 = cpy;
                        break;
                    case 'c': case 'C': {
                        float ox = rel ? cpx : 0.0f, oy = rel ? cpy : 0.0f;
                        float c1x = rel ? ox + args[0] : args[0], c1y = rel ? oy + args[1] : args[1];
                        float c2x = rel ? ox + args[2] : args[2], c2y = rel ? oy + args[3] : args[3];
                        float ex = rel ? ox + args[4] : args[4], ey = rel ? oy + args[5] : args[5];
                        cubic_to(c1x, c1y, c2x, c2y, ex, ey);
                        cpx2 = c2x; cpy2 = c2y; cpx = ex; cpy = ey;
                        break;
                    }
                    case 's': case 'S': {
                        float x1 = cpx, y1 = cpy;
                        float c2x = rel ? cpx + args[0] : args[0], c2y = rel ? cpy + args[1] : args[1];
                        float ex = rel ? cpx + args[2] : args[2], ey = rel ? cpy + args[3] : args[3];
                        float c1x = 2 * x1 - cpx2, c1y = 2 * y1 - cpy2;
                        cubic_to(c1x, c1y, c2x, c2y, ex, ey);
                        cpx2 = c2x; cpy2 = c2y; cpx = ex; cpy = ey;
                        break;
                    }
                    case 'q': case 'Q': case 't': case 'T': {
                        float x1 = cpx, y1 = cpy, qx, qy, ex, ey;
                        if (cmd == 'q' || cmd == 'Q') {
                            qx = rel ? cpx + args[0] : args[0]; qy = rel ? cpy + args[1] : args[1];
                            ex = rel ? cpx + args[2] : args[2]; ey = rel ? cpy + args[3] : args[3];
                        } else {
                            ex = rel ? cpx + args[0] : args[0]; ey = rel ? cpy + args[1] : args[1];
                            qx = 2 * x1 - cpx2; qy = 2 * y1 - cpy2;
                        }
                        float c1x = x1 + 2.0f / 3.0f * (qx - x1), c1y = y1 + 2.0f / 3.0f * (qy - y1);
                        float c2x = ex + 2.0f / 3.0f * (qx - ex), c2y = ey + 2.0f / 3.0f * (qy - ey);
                        cubic_to(c1x, c1y, c2x, c2y, ex, ey);
                        cpx2 = qx; cpy2 = qy; cpx = ex; cpy = ey;
                        break;
                    }
                    case 'a': case 'A':
                        arc_to(cpx, cpy, args, rel);
                        cpx2 = cpx; cpy2 = cpy;
                        break;
                    default:
                        if (nargs >= 2) { cpx = args[nargs - 2]; cpy = args[nargs - 1]; cpx2 = cpx; cpy2 = cpy; }
                        break;
                    }
                    nargs = 0;
                } else {
                    cmd   = item[0];
                    rargs = args_per_cmd(cmd);
                    if (cmd == 'M' || cmd == 'm') {
                        if (npts() > 0) commit_path(closed);
                        pts.clear();
                        closed = false; nargs = 0;
                    } else if (cmd == 'Z' || cmd == 'z') {
                        closed = true;
                        if (npts() > 0) {
                            cpx = pts[0]; cpy = pts[1]; cpx2 = cpx; cpy2 = cpy;
                            commit_path(closed);
                        }
                        pts.clear();
                        move_to(cpx, cpy);
                        closed = false; nargs = 0;
                    }
                }
            }
            if (npts()) commit_path(closed);
        }
        commit_shape();
    }
    void el_rect(const AttrList &al) {
        float x = 0, y = 0, w = 0, h = 0, rx = -1.0f, ry = -1.0f;
        for (auto &kv : al) {
            if (attr(kv.first, kv.second)) continue;
            if (!strcmp(kv.first, "x")) x = coord(kv.second, viewMinx, viewWidth);
            if (!strcmp(kv.first, "y")) y = coord(kv.second, viewMiny, viewHeight);
            if (!strcmp(kv.first, "width")) w = coord(kv.second, 0.0f, viewWidth);
            if (!strcmp(kv.first, "height")) h = coord(kv.second, 0.0f, viewHeight);
            if (!strcmp(kv.first, "rx")) rx = fabsf(coord(kv.second, 0.0f, viewWidth));
            if (!strcmp(kv.first, "ry")) ry = fabsf(coord(kv.second, 0.0f, viewHeight));
        }
        if (rx < 0.0f && ry > 0.0f) rx = ry;
        if (ry < 0.0f && rx > 0.0f) ry = rx;
        if (rx < 0.0f) rx = 0.0f;
        if (ry < 0.0f) ry = 0.0f;
        if (rx > w / 2.0f) rx = w / 2.0f;
        if (ry > h / 2.0f) ry = h / 2.0f;
        if (w == 0.0f || h == 0.0f) return;
        pts.clear();
        if (rx < 0.00001f || ry < 0.0001f) {
            move_to(x, y);
            line_to(x + w, y);
            line_to(x + w, y + h);
            line_to(x, y + h);
        } else {
            const float k = 1 - kKappa90;
            move_to(x + rx, y);
            line_to(x + w - rx, y);
            cubic_to(x + w - rx * k, y, x + w, y + ry * k, x + w, y + ry);
            line_to(x + w, y + h - ry);
            cubic_to(x + w, y + h - ry * k, x + w - rx * k, y + h, x + w - rx, y + h);
            line_to(x + rx, y + h);
            cubic_to(x + rx * k, y + h, x, y + h - ry * k, x, y + h - ry);
            line_to(x, y + ry);
            cubic_to(x, y + ry * k, x + rx * k, y, x + rx, y);
        }
        commit_path(true);
        commit_shape();
    }
    void ellipse_path(float cx, float cy, float rx, float ry) {
        pts.clear();
        move_to(cx + rx, cy);
        cubic_to(cx + rx, cy + ry * kKappa90, cx + rx * kKappa90, cy + ry, cx, cy + ry);
        cubic_to(cx - rx * kKappa90, cy + ry, cx - rx, cy + ry * kKappa90, cx - rx, cy);
        cubic_to(cx - rx, cy - ry * kKappa90, cx - rx * kKappa90, cy - ry, cx, cy - ry);
        cubic_to(cx + rx * kKappa90, cy - ry, cx + rx, cy - ry * kKappa90, cx + rx, cy);
        commit_path(true);
        commit_shape();
    }
    void el_circle(const AttrList &al) {
        float cx = 0, cy = 0, r = 0;
        for (auto &kv : al) {
            if (attr(kv.first, kv.second)) continue;
            if (!strcmp(kv.first, "cx")) cx = coord(kv.second, viewMinx, viewWidth);
            if (!strcmp(kv.first, "cy")) cy = coord(kv.second, viewMiny, viewHeight);
            if (!strcmp(kv.first, "r")) r = fabsf(coord(kv.second, 0.0f, actual_length()));
        }
        if (r > 0.0f) ellipse_path(cx, cy, r, r);
    }
    void el_ellipse(const AttrList &al) {
        float cx = 0, cy = 0, rx = 0, ry = 0;
        for (auto &kv : al) {
            if (attr(kv.first, kv.second)) continue;
            if (!strcmp(kv.first, "cx")) cx = coord(kv.second, viewMinx, viewWidth);
            if (!strcmp(kv.first, "cy")) cy = coord(kv.second, viewMiny, viewHeight);
            if (!strcmp(kv.first, "rx")) rx = fabsf(coord(kv.second, 0.0f, viewWidth));
            if (!strcmp(kv.first, "ry")) ry = fabsf(coord(kv.second, 0.0f, viewHeight));
        }
        if (rx > 0.0f && ry > 0.0f) ellipse_path(cx, cy, rx, ry);
    }
    void el_line(const AttrList &al) {
        float x1 = 0, y1 = 0, x2 = 0, y2 = 0;
        for (auto &kv : al) {
            if (attr(kv.first, kv.second)) continue;
            if (!strcmp(kv.first, "x1")) x1 = coord(kv.second, viewMinx, viewWidth);
            if (!strcmp(kv.first, "y1")) y1 = coord(kv.second, viewMiny, viewHeight);
            if (!strcmp(kv.first, "x2")) x2 = coord(kv.second, viewMinx, viewWidth);
            if (!strcmp(kv.first, "y2")) y2 = coord(kv.second, viewMiny, viewHeight);
        }
        pts.clear();
        move_to(x1, y1);
        line_to(x2, y2);
        commit_path(false);
        commit_shape();
    }
    void el_poly(const AttrList &al, bool close) {
        pts.clear();
        for (auto &kv : al) {
            if (attr(kv.first, kv.second)) continue;
            if (strcmp(kv.first, "points")) continue;
            const char *s = kv.second;
            float       args[2];
            int         nargs = 0, n = 0;
            char        item[64];
            while (*s) {
                s             = next_path_item(s, item);
                args[nargs++] = (float)scan_double(item);
                if (nargs >= 2) {
                    if (n == 0) move_to(args[0], args[1]);
                    else line_to(args[0], args[1]);
                    nargs = 0;
                    n++;
                }
            }
        }
        commit_path(close);
        commit_shape();
    }
    void el_svg(const AttrList &al) {
        for (auto &kv : al) {
            if (attr(kv.first, kv.second)) continue;
            if (!strcmp(kv.first, "width")) width = coord(kv.second, 0.0f, 0.0f);
            else if (!strcmp(kv.first, "height")) height = coord(kv.second, 0.0f, 0.0f);
            else if (!strcmp(kv.first, "viewBox")) {
                const char *s = kv.second;
                char        buf[64];
                float      *dst[4] = {&viewMinx, &viewMiny, &viewWidth, &viewHeight};
                for (int k = 0; k < 4; k++) {
                    s       = scan_number_token(s, buf, 64);
                    *dst[k] = (float)scan_double(buf);
                    if (k == 3) break;
                    while (is_space(*s) || *s == '%' || *s == ',') s++;
                    if (!*s) return;
                }
            } else if (!strcmp(kv.first, "preserveAspectRatio")) {
                const char *v = kv.second;
                if (strstr(v, "none")) alignType = ALIGN_NONE;
                else {
                    if (strstr(v, "xMin")) alignX = ALIGN_MIN; else if (strstr(v, "xMid")) alignX = ALIGN_MID; else if (strstr(v, "xMax")) alignX = ALIGN_MAX;
                    if (strstr(v, "yMin")) alignY = ALIGN_MIN; else if (strstr(v, "yMid")) alignY = ALIGN_MID; else if (strstr(v, "yMax")) alignY = ALIGN_MAX;
                    alignType = strstr(v, "slice") ? ALIGN_SLICE : ALIGN_MEET;
                }
            }
        }
    }
    void el_gradient(const AttrList &al, int type) {
        GradientData g;
        memset(g.id, 0, sizeof g.id);
        memset(g.ref, 0, sizeof g.ref);
        g.type = type;
        for (auto &kv : al) {
            if (!strcmp(kv.first, "id")) { strncpy(g.id, kv.second, 63); g.id[63] = '\0'; }
            else if (!attr(kv.first, kv.second)) {
                if (!strcmp(kv.first, "xlink:href")) { strncpy(g.ref, kv.second + (kv.second[0] ? 1 : 0), 62); g.ref[62] = '\0'; }
                // geometry (x1.., cx.., gradientTransform, spreadMethod): the reference driver only reads stops[0].color
            }
        }
        grads.push_back(g);
    }
    void el_stop(const AttrList &al) {
        Style &a = cur();
        a.stopOffset = 0; a.stopColor = 0; a.stopOpacity = 1.0f;
        for (auto &kv : al) attr(kv.first, kv.second);
        if (grads.empty()) return;
        auto  &st  = grads.back().stops;
        size_t idx = st.size();
        for (size_t i = 0; i < st.size(); i++)
            if (cur().stopOffset < st[i].first) { idx = i; break; }
        st.insert(st.begin() + idx, std::make_pair(cur().stopOffset, cur().stopColor | ((uint32_t)(cur().stopOpacity * 255) << 24)));
    }
    void start_element(const char *el, const AttrList &al) {
        if (inDefs) {  // only gradients are read inside <defs>
            if (!strcmp(el, "linearGradient")) el_gradient(al, PAINT_LINEAR);
            else if (!strcmp(el, "radialGradient")) el_gradient(al, PAINT_RADIAL);
            else if (!strcmp(el, "stop")) el_stop(al);
            return;
        }
        if (!strcmp(el, "g")) { push(); attribs(al); }
        else if (!strcmp(el, "path")) { push(); el_path(al); pop(); }
        else if (!strcmp(el, "rect")) { push(); el_rect(al); pop(); }
        else if (!strcmp(el, "circle")) { push(); el_circle(al); pop(); }
        else if (!strcmp(el, "ellipse")) { push(); el_ellipse(al); pop(); }
        else if (!strcmp(el, "line")) { push(); el_line(al); pop(); }
        else if (!strcmp(el, "polyline")) { push(); el_poly(al, false); pop(); }
        else if (!strcmp(el, "polygon")) { push(); el_poly(al, true); pop(); }
        else if (!strcmp(el, "linearGradient")) el_gradient(al, PAINT_LINEAR);
        else if (!strcmp(el, "radialGradient")) el_gradient(al, PAINT_RADIAL);
        else if (!strcmp(el, "stop")) el_stop(al);
        else if (!strcmp(el, "defs")) inDefs = true;
        else if (!strcmp(el, "svg")) el_svg(al);
    }
    void end_element(const char *el) {
        if (!strcmp(el, "g")) pop();
        else if (!strcmp(el, "defs")) inDefs = false;
    }

    // one "<...>" body (NUL terminated, modified in place)
    void tag(char *s) {
        AttrList al;
        bool     start = false, end = false;
        while (is_space(*s)) s++;
        if (*s == '/') { s++; end = true; } else start = true;
        if (!*s || *s == '?' || *s == '!') return;
        char *name = s;
        while (*s && !is_space(*s)) s++;
        if (*s) *s++ = '\0';
        while (!end && *s && al.size() < (256 - 3) / 2 + 1) {
            while (is_space(*s)) s++;
            if (!*s) break;
            if (*s == '/') { end = true; break; }
            char *an = s;
            while (*s && !is_space(*s) && *s != '=') s++;
            if (*s) *s++ = '\0';
            while (*s && *s != '\"' && *s != '\'') s++;
            if (!*s) break;
            char quote = *s++;
            char *av   = s;
            while (*s && *s != quote) s++;
            if (*s) *s++ = '\0';
            al.push_back(std::make_pair((const char *)an, (const char *)av));
        }
        if (start) start_element(name, al);
        if (end) end_element(name);
    }
    void document(char *input) {
        char *s = input, *mark = input;
        bool  in_tag = false;
        while (*s) {
            if (*s == '<' && !in_tag) { *s++ = '\0'; mark = s; in_tag = true; }
            else if (*s == '>' && in_tag) { *s++ = '\0'; tag(mark); mark = s; in_tag = false; }
            else s++;
        }
    }

    // map the user space of the root element onto width x height pixels
    void scale_to_viewbox() {
        float bounds[4] = {0, 0, 0, 0};
        if (!shapes.empty()) {
            memcpy(bounds, shapes[0].bounds, sizeof bounds);
            for (size_t i = 1; i < shapes.size(); i++) {
                bounds[0] = minf(bounds[0], shapes[i].bounds[0]); bounds[1] = minf(bounds[1], shapes[i].bounds[1]);
                bounds[2] = maxf(bounds[2], shapes[i].bounds[2]); bounds[3] = maxf(bounds[3], shapes[i].bounds[3]);
            }
        }
        if (viewWidth == 0) {
            if (width > 0) viewWidth = width;
            else { viewMinx = bounds[0]; viewWidth = bounds[2] - bounds[0]; }
        }
        if (viewHeight == 0) {
            if (height > 0) viewHeight = height;
            else { viewMiny = bounds[1]; viewHeight = bounds[3] - bounds[1]; }
        }
        if (width == 0) width = viewWidth;
        if (height == 0) height = viewHeight;
        float tx = -viewMinx, ty = -viewMiny;
        float sx = viewWidth > 0 ? width / viewWidth : 0, sy = viewHeight > 0 ? height / viewHeight : 0;
        Coord one = {1.0f, U_PX};  // vkvg_svg_load always asks for "px"
        float us  = 1.0f / to_pixels(one, 0.0f, 1.0f);
        auto view_align = [](float content, float container, int type) {
            if (type == ALIGN_MIN) return 0.0f;
            if (type == ALIGN_MAX) return container - content;
            return (container - content) * 0.5f;
        };
        if (alignType == ALIGN_MEET) {
            sx = sy = minf(sx, sy);
            tx += view_align(viewWidth * sx, width, alignX) / sx;
            ty += view_align(viewHeight * sy, height, alignY) / sy;
        } else if (alignType == ALIGN_SLICE) {
            sx = sy = maxf(sx, sy);
            tx += view_align(viewWidth * sx, width, alignX) / sx;
            ty += view_align(viewHeight * sy, height, alignY) / sy;
        }
        sx *= us; sy *= us;
        float avgs = (sx + sy) / 2.0f;
        for (SvgShape &sh : shapes) {
            for (SvgPath &p : sh.paths)
                for (size_t i = 0; i < p.pts.size(); i += 2) {
                    p.pts[i]     = (p.pts[i] + tx) * sx;
                    p.pts[i + 1] = (p.pts[i + 1] + ty) * sy;
                }
            sh.strokeWidth *= avgs;
        }
    }
};

}  // namespace

struct _vkvg_svg_t {
    float                 width, height;
    std::vector<SvgShape> shapes;
};

static VkvgSvg parse_svg_text(char *text, float dpi) {
    Parser p;
    p.dpi = dpi;
    p.document(text);
    p.scale_to_viewbox();
    VkvgSvg svg = new _vkvg_svg_t();
    svg->width  = p.width;
    svg->height = p.height;
    svg->shapes.swap(p.shapes);
    return svg;
}
static char *read_file(const char *path) {
    FILE *fp = path ? fopen(path, "rb") : nullptr;
    if (!fp) return nullptr;
    fseek(fp, 0, SEEK_END);
    long size = ftell(fp);
    fseek(fp, 0, SEEK_SET);
    char *data = size >= 0 ? (char *)malloc((size_t)size + 1) : nullptr;
    if (data && fread(data, 1, (size_t)size, fp) != (size_t)size) { free(data); data = nullptr; }
    if (data) data[size] = '\0';
    fclose(fp);
    return data;
}

extern "C" {

VkvgSvg vkvg_svg_load(const char *svgFilePath) {  // src/nsvg/vkvg_nsvg.c:69 ("px", 96 dpi); NULL when unreadable
    char *data = read_file(svgFilePath);
    if (!data) return NULL;
    VkvgSvg svg = parse_svg_text(data, 96.0f);
    free(data);
    return svg;
}
VkvgSvg vkvg_svg_load_fragment(char *svgFragment) {  // :70 — the reference parser scribbles on its input; this one works on a copy
    if (!svgFragment) return NULL;
    std::string copy(svgFragment);
    return parse_svg_text(&copy[0], 96.0f);
}
void vkvg_svg_destroy(VkvgSvg svg) { delete svg; }
void vkvg_svg_get_dimensions(VkvgSvg svg, uint32_t *width, uint32_t *height) {  // :72-75
    if (!svg) return;
    *width  = (uint32_t)svg->width;
    *height = (uint32_t)svg->height;
}
static void svg_set_color(VkvgContext ctx, uint32_t c, float alpha) {  // _svg_set_color :29-35
    float a = (c >> 24 & 255) / 255.f;
    float b = (c >> 16 & 255) / 255.f;
    float g = (c >> 8 & 255) / 255.f;
    float r = (c & 255) / 255.f;
    vkvg_set_source_rgba(ctx, r, g, b, a * alpha);
}
void vkvg_svg_render(VkvgSvg svg, VkvgContext ctx, const char *subId) {  // :79-136
    if (!svg) return;
    vkvg_save(ctx);
    vkvg_set_fill_rule(ctx, VKVG_FILL_RULE_EVEN_ODD);
    vkvg_set_source_rgba(ctx, 0.0, 0.0, 0.0, 1);
    for (const SvgShape &shape : svg->shapes) {
        if (subId != NULL && strcmp(shape.id, subId) != 0) continue;
        vkvg_new_path(ctx);
        const float o = shape.opacity;
        vkvg_set_line_width(ctx, shape.strokeWidth);
        for (const SvgPath &path : shape.paths) {
            const float *p = path.pts.data();
            const int    n = (int)(path.pts.size() / 2);
            vkvg_move_to(ctx, p[0], p[1]);
            for (int i = 1; i < n; i += 3) {
                const float *q = p + 2 * i;
                vkvg_curve_to(ctx, q[0], q[1], q[2], q[3], q[4], q[5]);
            }
            if (path.closed) vkvg_close_path(ctx);
        }
        // a radial gradient leaves the previous source in place (the driver only handles COLOR and LINEAR)
        if (shape.fillType == PAINT_COLOR || shape.fillType == PAINT_LINEAR) svg_set_color(ctx, shape.fillColor, o);
        if (shape.fillType != PAINT_NONE) {
            if (shape.strokeType == PAINT_NONE) { vkvg_fill(ctx); continue; }
            vkvg_fill_preserve(ctx);
        }
        if (shape.strokeType == PAINT_COLOR || shape.strokeType == PAINT_LINEAR) svg_set_color(ctx, shape.strokeColor, o);
        vkvg_stroke(ctx);
    }
    vkvg_restore(ctx);
}
static VkvgSurface surface_from_svg(VkvgDevice dev, VkvgSvg svg) {  // _svg_load :37-62: the surface takes the document's size
    if (!svg) return NULL;
    VkvgSurface surf = vkvg_surface_create(dev, (uint32_t)svg->width, (uint32_t)svg->height);
    if (!vkvg_surface_status(surf)) {
        VkvgContext ctx = vkvg_create(surf);
        vkvg_svg_render(svg, ctx, NULL);
        vkvg_destroy(ctx);
    }
    vkvg_svg_destroy(svg);
    return surf;
}
VkvgSurface vkvg_surface_create_from_svg(VkvgDevice dev, uint32_t, uint32_t, const char *svgFilePath) {  // :64-66 (device dpi)
    if (vkvg_device_status(dev)) return NULL;
    int hdpi = 96, vdpi = 96;
    vkvg_device_get_dpy(dev, &hdpi, &vdpi);
    char *data = read_file(svgFilePath);
    if (!data) return NULL;
    VkvgSvg svg = parse_svg_text(data, (float)hdpi);
    free(data);
    return surface_from_svg(dev, svg);
}
VkvgSurface vkvg_surface_create_from_svg_fragment(VkvgDevice dev, uint32_t, uint32_t, char *svgFragment) {  // :67-69
    if (vkvg_device_status(dev) || !svgFragment) return NULL;
    int hdpi = 96, vdpi = 96;
    vkvg_device_get_dpy(dev, &hdpi, &vdpi);
    std::string copy(svgFragment);
    return surface_from_svg(dev, parse_svg_text(&copy[0], (float)hdpi));
}

// flat dump of the shape list, same layout as oracle/nsvg_dump.c writes for nanoSVG (parity tests):
//   "NSVG" f32 width f32 height u32 nshapes | per shape: u32 fillType fillColor strokeType strokeColor, f32 opacity strokeWidth, u32 npaths
//   | per path: u32 npts u32 closed f32 pts[2*npts].  Returns the size in bytes; writes at most cap bytes.
uint64_t vkvg_b200_svg_serialize(VkvgSvg svg, uint8_t *out, uint64_t cap) {
    if (!svg) return 0;
    std::vector<uint8_t> buf;
    auto w32 = [&](uint32_t v) { const uint8_t *p = (const uint8_t *)&v; buf.insert(buf.end(), p, p + 4); };
    auto wf  = [&](float v) { const uint8_t *p = (const uint8_t *)&v; buf.insert(buf.end(), p, p + 4); };
    buf.insert(buf.end(), {'N', 'S', 'V', 'G'});
    wf(svg->width); wf(svg->height); w32((uint32_t)svg->shapes.size());
    for (const SvgShape &s : svg->shapes) {
        w32((uint32_t)s.fillType); w32(s.fillType == PAINT_NONE ? 0u : s.fillColor);
        w32((uint32_t)s.strokeType); w32(s.strokeType == PAINT_NONE ? 0u : s.strokeColor);
        wf(s.opacity); wf(s.strokeWidth); w32((uint32_t)s.paths.size());
        for (const SvgPath &p : s.paths) {
            w32((uint32_t)(p.pts.size() / 2)); w32(p.closed ? 1u : 0u);
            for (float v : p.pts) wf(v);
        }
    }
    if (out) memcpy(out, buf.data(), buf.size() < cap ? buf.size() : (size_t)cap);
    return buf.size();
}

}  // extern "C"
