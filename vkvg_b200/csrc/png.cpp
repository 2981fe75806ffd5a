// Minimal RGBA8 PNG encoder on top of zlib (the reference uses stb_image_write, src/vkvg_surface.c:384).
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <vector>
#include <zlib.h>

static void put32(std::vector<unsigned char> &v, uint32_t x) {
    v.push_back(x >> 24); v.push_back(x >> 16); v.push_back(x >> 8); v.push_back(x);
}
static void chunk(FILE *f, const char *type, const unsigned char *data, uint32_t len) {
    std::vector<unsigned char> b;
    put32(b, len);
    b.insert(b.end(), type, type + 4);
    if (len) b.insert(b.end(), data, data + len);
    uint32_t crc = crc32(0, b.data() + 4, len + 4);
    put32(b, crc);
    fwrite(b.data(), 1, b.size(), f);
}
int vkb_write_png(const char *path, const unsigned char *rgba, uint32_t w, uint32_t h) {
    FILE *f = fopen(path, "wb");
    if (!f) return 1;
    static const unsigned char sig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
    fwrite(sig, 1, 8, f);
    std::vector<unsigned char> ihdr;
    put32(ihdr, w); put32(ihdr, h);
    ihdr.push_back(8); ihdr.push_back(6); ihdr.push_back(0); ihdr.push_back(0); ihdr.push_back(0);
    chunk(f, "IHDR", ihdr.data(), (uint32_t)ihdr.size());
    std::vector<unsigned char> raw((size_t)h * (w * 4 + 1));
    for (uint32_t y = 0; y < h; y++) {
        raw[(size_t)y * (w * 4 + 1)] = 0;  // filter type none
        memcpy(&raw[(size_t)y * (w * 4 + 1) + 1], rgba + (size_t)y * w * 4, (size_t)w * 4);
    }
    uLongf clen = compressBound(raw.size());
    std::vector<unsigned char> comp(clen);
    if (compress2(comp.data(), &clen, raw.data(), raw.size(), 6) != Z_OK) { fclose(f); return 1; }
    chunk(f, "IDAT", comp.data(), (uint32_t)clen);
    chunk(f, "IEND", nullptr, 0);
    fclose(f);
    return 0;
}

// Minimal PNG decoder to RGBA8 on top of zlib's inflate (the reference loads images with stb_image, forcing four channels:
// src/vkvg_surface.c:151-170).  Handles 8-bit grey, grey+alpha, RGB, RGBA and palette (with tRNS) images, 16-bit ones by
// keeping the high byte, no interlacing.  Returns 0 on success.
static uint32_t get32(const unsigned char *p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }
static int paeth(int a, int b, int c) {
    int p = a + b - c, pa = p > a ? p - a : a - p, pb = p > b ? p - b : b - p, pc = p > c ? p - c : c - p;
    return (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
}
int vkb_read_png(const char *path, std::vector<unsigned char> &rgba, uint32_t &w, uint32_t &h) {
    FILE *f = path ? fopen(path, "rb") : nullptr;
    if (!f) return 1;
    std::vector<unsigned char> file;
    unsigned char buf[65536];
    size_t n;
    while ((n = fread(buf, 1, sizeof buf, f)) > 0) file.insert(file.end(), buf, buf + n);
    fclose(f);
    static const unsigned char sig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
    if (file.size() < 33 || memcmp(file.data(), sig, 8)) return 2;
    uint32_t depth = 0, ctype = 0, interlace = 0;
    std::vector<unsigned char> idat, plte, trns;
    w = h = 0;
    for (size_t pos = 8; pos + 12 <= file.size();) {
        const uint32_t       len = get32(&file[pos]);
        const unsigned char *typ = &file[pos + 4], *data = &file[pos + 8];
        if (pos + 12 + (size_t)len > file.size()) return 2;
        if (!memcmp(typ, "IHDR", 4) && len >= 13) { w = get32(data); h = get32(data + 4); depth = data[8]; ctype = data[9]; interlace = data[12]; }
        else if (!memcmp(typ, "PLTE", 4)) plte.assign(data, data + len);
        else if (!memcmp(typ, "tRNS", 4)) trns.assign(data, data + len);
        else if (!memcmp(typ, "IDAT", 4)) idat.insert(idat.end(), data, data + len);
        else if (!memcmp(typ, "IEND", 4)) break;
        pos += 12 + (size_t)len;
    }
    static const int chans[7] = {1, 0, 3, 1, 2, 0, 4};
    if (!w || !h || w > 65535 || h > 65535 || interlace || ctype > 6 || !chans[ctype] || (depth != 8 && depth != 16) || (ctype == 3 && depth != 8)) return 3;
    const size_t bpp = (size_t)chans[ctype] * (depth / 8), stride = (size_t)w * bpp;
    std::vector<unsigned char> raw((stride + 1) * h);
    uLongf rawlen = (uLongf)raw.size();
    if (uncompress(raw.data(), &rawlen, idat.data(), (uLong)idat.size()) != Z_OK || rawlen != raw.size()) return 4;
    std::vector<unsigned char> prev(stride, 0), cur(stride);
    rgba.assign((size_t)w * h * 4, 255);
    for (uint32_t y = 0; y < h; y++) {
        const unsigned char *row = &raw[(stride + 1) * y];
        const int            ft  = row[0];
        for (size_t x = 0; x < stride; x++) {
            const int a = x >= bpp ? cur[x - bpp] : 0, b = prev[x], c = x >= bpp ? prev[x - bpp] : 0;
            int       v = row[1 + x];
            switch (ft) {
            case 1: v += a; break;
            case 2: v += b; break;
            case 3: v += (a + b) / 2; break;
            case 4: v += paeth(a, b, c); break;
            default: break;
            }
            cur[x] = (unsigned char)v;
        }
        unsigned char *o = &rgba[(size_t)y * w * 4];
        const size_t   s = depth / 8;  // bytes per sample: the high byte comes first
        for (uint32_t x = 0; x < w; x++, o += 4) {
            const unsigned char *p = &cur[x * bpp];
            switch (ctype) {
            case 0: o[0] = o[1] = o[2] = p[0]; break;
            case 2: o[0] = p[0]; o[1] = p[s]; o[2] = p[2 * s]; break;
            case 3: {
                const unsigned i = p[0];
                if (3 * i + 2 < plte.size()) { o[0] = plte[3 * i]; o[1] = plte[3 * i + 1]; o[2] = plte[3 * i + 2]; }
                if (i < trns.size()) o[3] = trns[i];
                break;
            }
            case 4: o[0] = o[1] = o[2] = p[0]; o[3] = p[s]; break;
            case 6: o[0] = p[0]; o[1] = p[s]; o[2] = p[2 * s]; o[3] = p[3 * s]; break;
            }
        }
        prev.swap(cur);
    }
    return 0;
}
