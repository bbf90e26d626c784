// image_io.cpp — image files for a frame (the reference only shows its frame in a GL window,
// gpu_anim.h; `rgba8` below is that frame buffer, bottom-up rows, camera.cuh:70-78,194-207).
//   .ppm  binary P6, top-down             .png  8-bit RGB, zlib "stored" blocks (no compressor dependency; any reader accepts it)
//   .pfm  linear radiance, bottom-up rows (the frame's own order), little-endian floats
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "mort_b200.h"

namespace {

uint32_t crc_table[256]; bool crc_ready = false;
uint32_t crc32(uint32_t crc, const uint8_t* p, size_t n) {
    if (!crc_ready) { for (uint32_t i = 0; i < 256; i++) { uint32_t c = i; for (int k = 0; k < 8; k++) c = (c & 1) ? 0xEDB88320u ^ (c >> 1) : c >> 1; crc_table[i] = c; } crc_ready = true; }
    crc = ~crc;
    for (size_t i = 0; i < n; i++) crc = crc_table[(crc ^ p[i]) & 0xFF] ^ (crc >> 8);
    return ~crc;
}
void be32(uint8_t* p, uint32_t v) { p[0] = (uint8_t)(v >> 24); p[1] = (uint8_t)(v >> 16); p[2] = (uint8_t)(v >> 8); p[3] = (uint8_t)v; }
bool chunk(FILE* f, const char* type, const std::vector<uint8_t>& data) {
    uint8_t hdr[8]; be32(hdr, (uint32_t)data.size()); memcpy(hdr + 4, type, 4);
    uint32_t c = crc32(0, hdr + 4, 4); if (!data.empty()) c = crc32(c, data.data(), data.size());
    uint8_t tail[4]; be32(tail, c);
    return fwrite(hdr, 1, 8, f) == 8 && (data.empty() || fwrite(data.data(), 1, data.size(), f) == data.size()) && fwrite(tail, 1, 4, f) == 4;
}
bool ends_with(const std::string& s, const char* e) { const size_t n = strlen(e); return s.size() >= n && s.compare(s.size() - n, n, e) == 0; }

}  // namespace

extern "C" int mort_write_image(const char* path, const uint8_t* rgba8, int width, int height) {
    if (!path || !rgba8 || width <= 0 || height <= 0) return MORT_ERR_ARG;
    const std::string p = path;
    const bool png = ends_with(p, ".png"), ppm = ends_with(p, ".ppm");
    if (!png && !ppm) return MORT_ERR_ARG;
    FILE* f = fopen(path, "wb");
    if (!f) return MORT_ERR_IO;
    bool ok = true;
    if (ppm) {
        ok = fprintf(f, "P6\n%d %d\n255\n", width, height) > 0;
        std::vector<uint8_t> row((size_t)width * 3);
        for (int y = height - 1; y >= 0 && ok; y--) {
            for (int x = 0; x < width; x++) memcpy(&row[3 * (size_t)x], rgba8 + 4 * ((size_t)y * width + x), 3);
            ok = fwrite(row.data(), 1, row.size(), f) == row.size();
        }
    } else {
        static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};
        ok = fwrite(sig, 1, 8, f) == 8;
        std::vector<uint8_t> ihdr(13); be32(&ihdr[0], (uint32_t)width); be32(&ihdr[4], (uint32_t)height); ihdr[8] = 8; ihdr[9] = 2; ihdr[10] = ihdr[11] = ihdr[12] = 0;
        ok = ok && chunk(f, "IHDR", ihdr);
        // scanlines top-down, filter type 0, wrapped in stored deflate blocks of at most 65535 bytes
        const size_t stride = (size_t)width * 3 + 1, raw_n = stride * (size_t)height;
        std::vector<uint8_t> raw(raw_n);
        for (int y = 0; y < height; y++) {
            uint8_t* dst = &raw[(size_t)y * stride]; dst[0] = 0;
            const uint8_t* src = rgba8 + 4 * (size_t)(height - 1 - y) * width;
            for (int x = 0; x < width; x++) memcpy(dst + 1 + 3 * (size_t)x, src + 4 * (size_t)x, 3);
        }
        std::vector<uint8_t> z; z.reserve(raw_n + raw_n / 65535 * 5 + 16);
        z.push_back(0x78); z.push_back(0x01);
        uint32_t a = 1, b = 0;
        for (size_t off = 0; off < raw_n || off == 0; ) {
            const size_t n = std::min<size_t>(65535, raw_n - off);
            z.push_back(off + n >= raw_n ? 1 : 0);
            z.push_back((uint8_t)(n & 0xFF)); z.push_back((uint8_t)(n >> 8)); z.push_back((uint8_t)(~n & 0xFF)); z.push_back((uint8_t)((~n >> 8) & 0xFF));
            z.insert(z.end(), raw.begin() + off, raw.begin() + off + n);
            for (size_t i = 0; i < n; i++) { a += raw[off + i]; if (a >= 65521) a -= 65521; b += a; if (b >= 65521) b -= 65521; }
            off += n;
            if (n == 0) break;
        }
        uint8_t ad[4]; be32(ad, (b << 16) | a); z.insert(z.end(), ad, ad + 4);
        ok = ok && chunk(f, "IDAT", z) && chunk(f, "IEND", std::vector<uint8_t>());
    }
    ok = (fclose(f) == 0) && ok;
    return ok ? MORT_OK : MORT_ERR_IO;
}

extern "C" int mort_write_pfm(const char* path, const float* accum4, int width, int height, float scale) {
    if (!path || !accum4 || width <= 0 || height <= 0) return MORT_ERR_ARG;
    FILE* f = fopen(path, "wb");
    if (!f) return MORT_ERR_IO;
    bool ok = fprintf(f, "PF\n%d %d\n-1.0\n", width, height) > 0;
    std::vector<float> row((size_t)width * 3);
    for (int y = 0; y < height && ok; y++) {
        for (int x = 0; x < width; x++) for (int c = 0; c < 3; c++) row[3 * (size_t)x + c] = accum4[4 * ((size_t)y * width + x) + c] * scale;
        ok = fwrite(row.data(), 4, row.size(), f) == row.size();
    }
    ok = (fclose(f) == 0) && ok;
    return ok ? MORT_OK : MORT_ERR_IO;
}
