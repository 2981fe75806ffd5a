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
