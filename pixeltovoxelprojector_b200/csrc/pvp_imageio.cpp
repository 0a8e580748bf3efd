// pvp_imageio.cpp -- host image decode feeding K1: the stbi_load(path,&w,&h,&c,1) call of
// load_image_gray (ray_voxel.cpp:195) WITHOUT the unseeded noise of :206-215.
//
// stb_image is an un-vendored, un-pinned dependency of the reference.  This is a
// from-scratch decoder for the lossless formats, following stb_image's PUBLISHED conversion
// rules so the gray bytes are the ones the reference would see:
//   - colour -> gray:   y = (77 r + 150 g + 29 b) >> 8  (integer, alpha ignored)
//   - 16-bit -> 8-bit:  v >> 8, applied after the colour conversion
//   - gray bit depths 1/2/4: scaled by 255/85/17
// Formats: binary PNM (P5, P6; maxval <= 65535), PNG (all colour types, 8/16 bit and sub-byte
// gray/palette, plain or Adam7-interlaced) and uncompressed BMP (1/4/8-bit palettised, 24 and
// 32 bit).  Anything else (JPEG, TGA, GIF, RLE or 16-bit BMP ...) fails to load, which the
// pipeline treats like the reference treats a failed stbi_load: the frame is skipped
// (ray_voxel.cpp:470-473).  JPEG parity is unpinned by design (stb's IDCT != libjpeg).
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <zlib.h>

#include <vector>

#include "pvp_internal.h"

namespace {

struct Decoded {
    int w = 0, h = 0;
    std::vector<uint8_t> gray;
};

bool read_file(const char *path, std::vector<uint8_t> &out)
{
    FILE *f = fopen(path, "rb");
    if (!f) return false;
    fseek(f, 0, SEEK_END);
    long n = ftell(f);
    fseek(f, 0, SEEK_SET);
    if (n <= 0) { fclose(f); return false; }
    out.resize((size_t)n);
    bool ok = fread(out.data(), 1, (size_t)n, f) == (size_t)n;
    fclose(f);
    return ok;
}

inline uint8_t luma8(unsigned r, unsigned g, unsigned b) { return (uint8_t)((r * 77 + g * 150 + b * 29) >> 8); }
inline uint16_t luma16(unsigned r, unsigned g, unsigned b) { return (uint16_t)((r * 77 + g * 150 + b * 29) >> 8); }

// ---- PNM -------------------------------------------------------------------
bool pnm_token(const std::vector<uint8_t> &d, size_t &pos, int &val)
{
    for (;;) {
        while (pos < d.size() && (d[pos] == ' ' || d[pos] == '\t' || d[pos] == '\n' || d[pos] == '\r')) ++pos;
        if (pos < d.size() && d[pos] == '#') { while (pos < d.size() && d[pos] != '\n') ++pos; continue; }
        break;
    }
    if (pos >= d.size() || d[pos] < '0' || d[pos] > '9') return false;
    long v = 0;
    while (pos < d.size() && d[pos] >= '0' && d[pos] <= '9') { v = v * 10 + (d[pos] - '0'); if (v > (1 << 30)) return false; ++pos; }
    val = (int)v;
    return true;
}

bool decode_pnm(const std::vector<uint8_t> &d, Decoded &out)
{
    if (d.size() < 7 || d[0] != 'P' || (d[1] != '5' && d[1] != '6')) return false;
    const int comp = d[1] == '5' ? 1 : 3;
    size_t pos = 2;
    int w, h, maxv;
    if (!pnm_token(d, pos, w) || !pnm_token(d, pos, h) || !pnm_token(d, pos, maxv)) return false;
    if (w <= 0 || h <= 0 || maxv <= 0 || maxv > 65535) return false;
    ++pos;  // the single whitespace byte after maxval
    const int bps = maxv > 255 ? 2 : 1;
    const size_t npx = (size_t)w * h;
    if (pos + npx * comp * bps > d.size()) return false;
    const uint8_t *p = d.data() + pos;
    out.w = w; out.h = h; out.gray.resize(npx);
    if (bps == 1) {
        if (comp == 1) memcpy(out.gray.data(), p, npx);
        else for (size_t i = 0; i < npx; ++i) out.gray[i] = luma8(p[3 * i], p[3 * i + 1], p[3 * i + 2]);
    } else {  // big-endian 16 bit
        auto rd = [&](size_t k) { return (unsigned)((p[2 * k] << 8) | p[2 * k + 1]); };
        if (comp == 1) for (size_t i = 0; i < npx; ++i) out.gray[i] = (uint8_t)(rd(i) >> 8);
        else for (size_t i = 0; i < npx; ++i) out.gray[i] = (uint8_t)(luma16(rd(3 * i), rd(3 * i + 1), rd(3 * i + 2)) >> 8);
    }
    return true;
}

// ---- PNG (non-interlaced) ----------------------------------------------------
inline uint32_t be32(const uint8_t *p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }

inline int paeth(int a, int b, int c)
{
    int p = a + b - c, pa = abs(p - a), pb = abs(p - b), pc = abs(p - c);
    if (pa <= pb && pa <= pc) return a;
    return pb <= pc ? b : c;
}

bool decode_png(const std::vector<uint8_t> &d, Decoded &out)
{
    static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
    if (d.size() < 33 || memcmp(d.data(), sig, 8) != 0) return false;
    size_t pos = 8;
    uint32_t w = 0, h = 0;
    int depth = 0, ctype = 0, interlace = 0;
    std::vector<uint8_t> idat, pal;
    bool have_ihdr = false;
    while (pos + 12 <= d.size()) {
        const uint32_t len = be32(&d[pos]);
        const uint8_t *type = &d[pos + 4];
        if (pos + 12 + (size_t)len > d.size()) return false;
        const uint8_t *body = &d[pos + 8];
        if (!memcmp(type, "IHDR", 4)) {
            if (len != 13) return false;
            w = be32(body); h = be32(body + 4); depth = body[8]; ctype = body[9]; interlace = body[12];
            if (body[10] != 0 || body[11] != 0) return false;
            have_ihdr = true;
        } else if (!memcmp(type, "PLTE", 4)) {
            pal.assign(body, body + len);
        } else if (!memcmp(type, "IDAT", 4)) {
            idat.insert(idat.end(), body, body + len);
        } else if (!memcmp(type, "IEND", 4)) {
            break;
        }
        pos += 12 + (size_t)len;
    }
    if (!have_ihdr || w == 0 || h == 0 || w > (1u << 24) || h > (1u << 24) || interlace > 1) return false;
    int chans;
    switch (ctype) {
        case 0: chans = 1; break;
        case 2: chans = 3; break;
        case 3: chans = 1; break;
        case 4: chans = 2; break;
        case 6: chans = 4; break;
        default: return false;
    }
    if (!(depth == 8 || depth == 16 || ((ctype == 0 || ctype == 3) && (depth == 1 || depth == 2 || depth == 4)))) return false;
    if (ctype == 3 && (depth == 16 || pal.size() < 3)) return false;
    const size_t bpp_bits = (size_t)chans * depth;
    const size_t fbpp = bpp_bits >= 8 ? bpp_bits / 8 : 1;  // filter byte distance

    // The image is one pass, or the seven Adam7 passes of an interlaced file: each pass is a sub-image of the pixels
    // (x0 + i*dx, y0 + j*dy) with its own scanlines and filters; empty passes take no bytes.
    struct Pass { size_t x0, y0, dx, dy, pw, ph, stride, offset; };
    static const unsigned adam7[7][4] = {{0, 0, 8, 8}, {4, 0, 8, 8}, {0, 4, 4, 8}, {2, 0, 4, 4}, {0, 2, 2, 4}, {1, 0, 2, 2}, {0, 1, 1, 2}};
    Pass passes[7];
    int n_pass = 0;
    size_t raw_bytes = 0, max_stride = 0;
    for (int k = 0; k < (interlace ? 7 : 1); ++k) {
        Pass ps;
        ps.x0 = interlace ? adam7[k][0] : 0; ps.y0 = interlace ? adam7[k][1] : 0;
        ps.dx = interlace ? adam7[k][2] : 1; ps.dy = interlace ? adam7[k][3] : 1;
        if (ps.x0 >= w || ps.y0 >= h) continue;
        ps.pw = (w - ps.x0 + ps.dx - 1) / ps.dx; ps.ph = (h - ps.y0 + ps.dy - 1) / ps.dy;
        ps.stride = (ps.pw * bpp_bits + 7) / 8;
        ps.offset = raw_bytes;
        raw_bytes += (ps.stride + 1) * ps.ph;
        if (ps.stride > max_stride) max_stride = ps.stride;
        passes[n_pass++] = ps;
    }
    if (idat.empty() || raw_bytes / 1032 > idat.size() + 64) return false;  // deflate expands at most ~1032x: a forged size
    std::vector<uint8_t> raw(raw_bytes);
    uLongf raw_len = (uLongf)raw.size();
    if (uncompress(raw.data(), &raw_len, idat.data(), (uLong)idat.size()) != Z_OK || raw_len != raw.size()) return false;

    out.w = (int)w; out.h = (int)h; out.gray.resize((size_t)w * h);
    const int npal = (int)(pal.size() / 3);
    static const int scale_tab[9] = {0, 0xff, 0x55, 0, 0x11, 0, 0, 0, 0x01};
    std::vector<uint8_t> zero(max_stride, 0);
    for (int k = 0; k < n_pass; ++k) {
        const Pass &ps = passes[k];
        const size_t stride = ps.stride;
        // un-filter in place (row r occupies raw[offset + r*(stride+1) + 1 ...]); one tight loop per filter type
        for (size_t r = 0; r < ps.ph; ++r) {
            uint8_t *cur = &raw[ps.offset + r * (stride + 1) + 1];
            const uint8_t *up = r ? cur - (stride + 1) : zero.data();
            const size_t head = fbpp < stride ? fbpp : stride;  // bytes without a left neighbour
            switch (cur[-1]) {
                case 0: break;
                case 1:
                    for (size_t i = head; i < stride; ++i) cur[i] = (uint8_t)(cur[i] + cur[i - fbpp]);
                    break;
                case 2:
                    for (size_t i = 0; i < stride; ++i) cur[i] = (uint8_t)(cur[i] + up[i]);
                    break;
                case 3:
                    for (size_t i = 0; i < head; ++i) cur[i] = (uint8_t)(cur[i] + (up[i] >> 1));
                    for (size_t i = head; i < stride; ++i) cur[i] = (uint8_t)(cur[i] + ((cur[i - fbpp] + up[i]) >> 1));
                    break;
                case 4:
                    for (size_t i = 0; i < head; ++i) cur[i] = (uint8_t)(cur[i] + up[i]);  // paeth(0, b, 0) = b
                    for (size_t i = head; i < stride; ++i) cur[i] = (uint8_t)(cur[i] + paeth(cur[i - fbpp], up[i], up[i - fbpp]));
                    break;
                default: return false;
            }
        }
        // pixels of the pass -> gray, scattered to (x0 + x*dx, y0 + r*dy)
        for (size_t r = 0; r < ps.ph; ++r) {
            const uint8_t *row = &raw[ps.offset + r * (stride + 1) + 1];
            uint8_t *o = &out.gray[(ps.y0 + r * ps.dy) * (size_t)w + ps.x0];
            const size_t dx = ps.dx;
            if (depth == 8 && ctype == 0 && dx == 1) { memcpy(o, row, ps.pw); continue; }
            for (size_t x = 0; x < ps.pw; ++x) {
                uint8_t g;
                if (depth == 16) {
                    auto rd = [&](size_t i) { return (unsigned)((row[2 * i] << 8) | row[2 * i + 1]); };
                    unsigned y;
                    if (chans <= 2) y = rd(x * chans);
                    else y = luma16(rd(x * chans), rd(x * chans + 1), rd(x * chans + 2));
                    g = (uint8_t)(y >> 8);
                } else if (depth == 8) {
                    if (ctype == 3) {
                        int idx = row[x];
                        if (idx >= npal) idx = 0;
                        g = luma8(pal[3 * idx], pal[3 * idx + 1], pal[3 * idx + 2]);
                    } else if (chans <= 2) {
                        g = row[x * chans];
                    } else {
                        g = luma8(row[x * chans], row[x * chans + 1], row[x * chans + 2]);
                    }
                } else {  // 1/2/4 bit gray or palette
                    const size_t bit = x * depth;
                    int v = (row[bit >> 3] >> (8 - depth - (bit & 7))) & ((1 << depth) - 1);
                    if (ctype == 3) {
                        if (v >= npal) v = 0;
                        g = luma8(pal[3 * v], pal[3 * v + 1], pal[3 * v + 2]);
                    } else {
                        g = (uint8_t)(v * scale_tab[depth]);
                    }
                }
                o[x * dx] = g;
            }
        }
    }
    return true;
}

// ---- BMP (uncompressed) -----------------------------------------------------------
// stb_image's BMP reader for req_comp = 1 decodes to RGB(A) and applies the integer luma; palette entries likewise.
// Handled: 1/4/8-bit palettised, 24-bit, 32-bit (BI_RGB, or BI_BITFIELDS with the plain 8-bit masks); bottom-up and top-down.
// Not handled (load fails, frame skipped): RLE, 16-bit and exotic bit-field layouts -- their expansion to 8 bits changed between
// stb_image versions and the reference pins none.
inline uint32_t le32(const uint8_t *p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }
inline uint32_t le16(const uint8_t *p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8); }

bool decode_bmp(const std::vector<uint8_t> &d, Decoded &out)
{
    if (d.size() < 14 + 12 || d[0] != 'B' || d[1] != 'M') return false;
    const size_t data_off = le32(&d[10]);
    const uint32_t hsz = le32(&d[14]);
    if (hsz != 12 && hsz != 40 && hsz != 56 && hsz != 108 && hsz != 124) return false;
    if (d.size() < 14 + (size_t)hsz) return false;
    long w, hh;
    int bpp;
    uint32_t compress = 0;
    if (hsz == 12) {
        w = (long)le16(&d[18]); hh = (long)le16(&d[20]);
        if (le16(&d[22]) != 1) return false;
        bpp = (int)le16(&d[24]);
    } else {
        w = (long)(int32_t)le32(&d[18]); hh = (long)(int32_t)le32(&d[22]);
        if (le16(&d[26]) != 1) return false;
        bpp = (int)le16(&d[28]);
        compress = le32(&d[30]);
    }
    const bool flip = hh > 0;  // positive height: rows stored bottom-up
    if (hh < 0) hh = -hh;
    if (w <= 0 || hh <= 0 || w > (1 << 24) || hh > (1 << 24)) return false;
    if (!(bpp == 1 || bpp == 4 || bpp == 8 || bpp == 24 || bpp == 32)) return false;
    if (compress == 3) {  // BI_BITFIELDS: only the layout BI_RGB means anyway
        if (bpp != 32) return false;
        const size_t m = 14 + 40;  // the masks follow the 40-byte core header (inside the larger headers)
        if (d.size() < m + 12) return false;
        if (le32(&d[m]) != 0x00ff0000u || le32(&d[m + 4]) != 0x0000ff00u || le32(&d[m + 8]) != 0x000000ffu) return false;
    } else if (compress != 0) {
        return false;
    }
    // palette: RGBQUAD (b, g, r, x), or RGBTRIPLE for the 12-byte header; like stb_image, every entry between the header and
    // the pixel data counts (biClrUsed is not consulted)
    uint8_t pal_gray[256];
    if (bpp <= 8) {
        const size_t esz = hsz == 12 ? 3 : 4;
        const size_t pal_off = 14 + (size_t)hsz;
        if (data_off < pal_off) return false;
        const size_t ncol = (data_off - pal_off) / esz;
        if (ncol == 0 || ncol > 256 || pal_off + ncol * esz > d.size()) return false;
        memset(pal_gray, 0, sizeof(pal_gray));
        for (size_t i = 0; i < ncol; ++i) { const uint8_t *e = &d[pal_off + i * esz]; pal_gray[i] = luma8(e[2], e[1], e[0]); }
    }
    const size_t stride = (((size_t)w * bpp + 31) / 32) * 4;
    if (data_off < 14 + (size_t)hsz || data_off > d.size() || stride * (size_t)hh > d.size() - data_off) return false;
    out.w = (int)w; out.h = (int)hh; out.gray.resize((size_t)w * hh);
    for (size_t r = 0; r < (size_t)hh; ++r) {
        const uint8_t *row = &d[data_off + r * stride];
        uint8_t *o = &out.gray[(flip ? (size_t)hh - 1 - r : r) * (size_t)w];
        switch (bpp) {
            case 1: for (size_t x = 0; x < (size_t)w; ++x) o[x] = pal_gray[(row[x >> 3] >> (7 - (x & 7))) & 1]; break;
            case 4: for (size_t x = 0; x < (size_t)w; ++x) o[x] = pal_gray[(x & 1) ? (row[x >> 1] & 15) : (row[x >> 1] >> 4)]; break;
            case 8: for (size_t x = 0; x < (size_t)w; ++x) o[x] = pal_gray[row[x]]; break;
            case 24: for (size_t x = 0; x < (size_t)w; ++x) o[x] = luma8(row[3 * x + 2], row[3 * x + 1], row[3 * x]); break;
            default: for (size_t x = 0; x < (size_t)w; ++x) o[x] = luma8(row[4 * x + 2], row[4 * x + 1], row[4 * x]); break;
        }
    }
    return true;
}

bool decode_any(const char *path, Decoded &out)
{
    try {  // a damaged header must not take the process down with std::bad_alloc (also runs on the decode pool's threads)
        std::vector<uint8_t> data;
        if (!read_file(path, data)) return false;
        if (data.size() >= 2 && data[0] == 'P') return decode_pnm(data, out);
        if (data.size() >= 2 && data[0] == 'B' && data[1] == 'M') return decode_bmp(data, out);
        return decode_png(data, out);
    } catch (...) {
        return false;
    }
}

}  // namespace

// One decode per file for the callers inside the library (the C entry below decodes once per call, so its
// size query + read pattern would decode every file twice).
bool pvp::decode_gray_file(const char *path, std::vector<uint8_t> &gray, int *width, int *height)
{
    Decoded d;
    if (!path || !decode_any(path, d)) return false;
    gray.swap(d.gray);
    *width = d.w; *height = d.h;
    return true;
}

extern "C" {

// Decode `path` to 8-bit gray.  out may be NULL (size query).  Returns PVP_ERR_IO when the file
// cannot be decoded (the reference's `stbi_load` == NULL case).
int pvp_load_image_gray(const char *path, uint8_t *out, int64_t cap, int *width, int *height)
{
    if (!path || !width || !height) { pvp::set_error("pvp_load_image_gray: NULL argument"); return PVP_ERR_INVALID; }
    Decoded d;
    if (!decode_any(path, d)) { pvp::set_error("Failed to load image: %s", path); return PVP_ERR_IO; }
    *width = d.w; *height = d.h;
    if (out) {
        if (cap < (int64_t)d.gray.size()) { pvp::set_error("pvp_load_image_gray: buffer too small"); return PVP_ERR_INVALID; }
        memcpy(out, d.gray.data(), d.gray.size());
    }
    return PVP_OK;
}

}  // extern "C"
