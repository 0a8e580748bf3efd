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
// gray/palette, plain or Adam7-interlaced), uncompressed BMP (1/4/8-bit palettised, 24 and 32 bit)
// and JPEG (luma plane; see the JPEG section: this one format is parity-UNPINNED).
// Anything else (TGA, GIF, RLE or 16-bit BMP ...) fails to load, which the
// pipeline treats like the reference treats a failed stbi_load: the frame is skipped
// (ray_voxel.cpp:470-473).
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <fcntl.h>
#include <unistd.h>
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

// ---- JPEG (sequential and progressive Huffman, 8 bit) ----------------------------------------
// What stbi_load(..., 1) returns for a YCbCr or gray JPEG is the LUMA PLANE as its integer IDCT produces it (chroma is not
// touched when one channel is requested), so this reader entropy-decodes every component but reconstructs only component 0,
// with stb_image's published IDCT arithmetic: the LL&M butterfly in 12-bit fixed point, columns kept with 2 extra bits
// ((x + 512) >> 10), rows rounded and biased in one step ((x + 65536 + (128 << 17)) >> 17), clamped to 0..255.  The reference
// pins no stb_image version and none exists in the build image, so this format's parity is UNPINNED; the
// test holds it to +-1 grey level of libjpeg and to exact values on flat blocks.  Sequential and progressive Huffman files are
// read.  Not handled (load fails, frame skipped): arithmetic-coded and lossless files, 12-bit samples, CMYK / RGB-coded
// components, a subsampled luma plane.
static const uint8_t jpeg_zigzag[64] = {0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6, 7, 14, 21, 28,
                                        35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};

struct JHuff {
    bool present = false;
    uint8_t fast_len[512], fast_sym[512];  // codes of up to 9 bits, indexed by the next 9 bits of the stream
    int16_t fast_ac[512];                  // AC tables: code + magnitude bits within 9 bits -> (value << 8) | (run << 4) | bits used; else 0
    int mincode[17], maxcode[17], valptr[17];
    uint8_t vals[256];
    bool build(const uint8_t counts[16], const uint8_t *syms, int n_syms)
    {
        memset(fast_len, 0, sizeof(fast_len));
        int code = 0, k = 0;
        for (int len = 1; len <= 16; ++len) {
            mincode[len] = code; valptr[len] = k; maxcode[len] = -1;
            for (int i = 0; i < counts[len - 1]; ++i, ++code, ++k) {
                if (k >= n_syms || k >= 256 || code >= (1 << len)) return false;  // more codes than the length has room for
                vals[k] = syms[k];
                if (len <= 9)
                    for (int f = 0; f < (1 << (9 - len)); ++f) { fast_len[(code << (9 - len)) | f] = (uint8_t)len; fast_sym[(code << (9 - len)) | f] = syms[k]; }
            }
            if (counts[len - 1]) maxcode[len] = code - 1;
            code <<= 1;
        }
        for (int i = 0; i < 512; ++i) {  // a short code followed by a short magnitude decodes in one lookup (same value as the two-step path)
            fast_ac[i] = 0;
            const int len = fast_len[i], run = fast_sym[i] >> 4, mag = fast_sym[i] & 15;
            if (len && mag && len + mag <= 9) {
                int v = ((i << len) & 511) >> (9 - mag);
                if (v < (1 << (mag - 1))) v += -(1 << mag) + 1;
                if (v >= -128 && v <= 127) fast_ac[i] = (int16_t)(v * 256 + run * 16 + len + mag);
            }
        }
        present = true;
        return true;
    }
};

struct JBits {
    const uint8_t *p, *end;
    uint64_t buf = 0;  // the next bits of the entropy-coded segment, left aligned at bit n-1
    int n = 0;
    int marker = 0;    // the marker that ended the segment (then zero bits are fed, as stb_image does)
    void fill()
    {
        if (n > 32) return;  // every caller needs at most 16 bits
        while (n <= 56) {
            unsigned b = 0;
            if (!marker) {
                if (p >= end) { marker = 0xD9; }
                else {
                    b = *p++;
                    if (b == 0xFF) {
                        while (p < end && *p == 0xFF) ++p;  // fill bytes
                        const unsigned c = p < end ? *p++ : 0xD9;
                        if (c != 0) { marker = (int)c; b = 0; }
                    }
                }
            }
            buf = (buf << 8) | b;
            n += 8;
        }
    }
    unsigned peek(int k) { return (unsigned)((buf >> (n - k)) & ((1u << k) - 1)); }
    void skip(int k) { n -= k; }
    int decode(const JHuff &h)
    {
        fill();
        const unsigned f = peek(9);
        if (h.fast_len[f]) { skip(h.fast_len[f]); return h.fast_sym[f]; }
        for (int len = 10; len <= 16; ++len) {
            const int c = (int)peek(len);
            if (h.maxcode[len] >= 0 && c <= h.maxcode[len] && c >= h.mincode[len]) { skip(len); return h.vals[h.valptr[len] + c - h.mincode[len]]; }
        }
        return -1;
    }
    int receive_extend(int s)  // s magnitude bits -> signed value (ITU T.81 F.2.2.1)
    {
        if (s == 0) return 0;
        fill();
        const int v = (int)peek(s);
        skip(s);
        return v < (1 << (s - 1)) ? v - (1 << s) + 1 : v;
    }
    void reset() { buf = 0; n = 0; marker = 0; }
};

// The butterfly runs in unsigned 32-bit arithmetic (wraps; the same low 32 bits as the signed form) and is read back as
// two's-complement int at the shifts, so damaged files with absurd coefficients cannot overflow a signed int.
#define PVP_F2F(x) ((unsigned)(int)(((x) * 4096 + 0.5)))
// one 8-point pass; leaves the even part in x0..x3 and the odd part in t0..t3
#define PVP_IDCT_1D(s0, s1, s2, s3, s4, s5, s6, s7)                                  \
    unsigned t0, t1, t2, t3, p1, p2, p3, p4, p5, x0, x1, x2, x3;                       \
    p2 = (unsigned)(s2); p3 = (unsigned)(s6);                                          \
    p1 = (p2 + p3) * PVP_F2F(0.5411961f);                                              \
    t2 = p1 + p3 * PVP_F2F(-1.847759065f);                                             \
    t3 = p1 + p2 * PVP_F2F(0.765366865f);                                              \
    p2 = (unsigned)(s0); p3 = (unsigned)(s4);                                          \
    t0 = (p2 + p3) * 4096u; t1 = (p2 - p3) * 4096u;                                    \
    x0 = t0 + t3; x3 = t0 - t3; x1 = t1 + t2; x2 = t1 - t2;                            \
    t0 = (unsigned)(s7); t1 = (unsigned)(s5); t2 = (unsigned)(s3); t3 = (unsigned)(s1); \
    p3 = t0 + t2; p4 = t1 + t3; p1 = t0 + t3; p2 = t1 + t2;                            \
    p5 = (p3 + p4) * PVP_F2F(1.175875602f);                                            \
    t0 = t0 * PVP_F2F(0.298631336f); t1 = t1 * PVP_F2F(2.053119869f);                  \
    t2 = t2 * PVP_F2F(3.072711026f); t3 = t3 * PVP_F2F(1.501321110f);                  \
    p1 = p5 + p1 * PVP_F2F(-0.899976223f); p2 = p5 + p2 * PVP_F2F(-2.562915447f);      \
    p3 = p3 * PVP_F2F(-1.961570560f); p4 = p4 * PVP_F2F(-0.390180644f);                \
    t3 += p1 + p4; t2 += p2 + p3; t1 += p2 + p4; t0 += p1 + p3;

inline uint8_t jclamp(int x) { return (uint8_t)(x < 0 ? 0 : x > 255 ? 255 : x); }
inline int jsar(unsigned x, int k) { return (int)x >> k; }  // arithmetic shift of the two's-complement value

void jpeg_idct_block(uint8_t *out, size_t out_stride, const short d[64])
{
    int val[64];
    for (int i = 0; i < 8; ++i) {  // columns
        int *v = val + i;
        if (!(d[8 + i] | d[16 + i] | d[24 + i] | d[32 + i] | d[40 + i] | d[48 + i] | d[56 + i])) {
            // the butterfly with only s0 set: every output is (s0 * 4096 + 512) >> 10 = s0 * 4, exactly
            v[0] = v[8] = v[16] = v[24] = v[32] = v[40] = v[48] = v[56] = d[i] * 4;
            continue;
        }
        PVP_IDCT_1D(d[i], d[8 + i], d[16 + i], d[24 + i], d[32 + i], d[40 + i], d[48 + i], d[56 + i])
        x0 += 512; x1 += 512; x2 += 512; x3 += 512;
        v[0] = jsar(x0 + t3, 10); v[56] = jsar(x0 - t3, 10);
        v[8] = jsar(x1 + t2, 10); v[48] = jsar(x1 - t2, 10);
        v[16] = jsar(x2 + t1, 10); v[40] = jsar(x2 - t1, 10);
        v[24] = jsar(x3 + t0, 10); v[32] = jsar(x3 - t0, 10);
    }
    for (int i = 0; i < 8; ++i) {  // rows: 1 << 17 to remove, rounded, and the level shift of 128 added before the shift
        const int *v = val + 8 * i;
        uint8_t *o = out + (size_t)i * out_stride;
        const unsigned bias = 65536u + (128u << 17);
        if (!(v[1] | v[2] | v[3] | v[4] | v[5] | v[6] | v[7])) {  // the same formulas with s1..s7 = 0: one value for the row
            memset(o, jclamp(jsar((unsigned)v[0] * 4096u + bias, 17)), 8);
            continue;
        }
        PVP_IDCT_1D(v[0], v[1], v[2], v[3], v[4], v[5], v[6], v[7])
        x0 += bias; x1 += bias; x2 += bias; x3 += bias;
        o[0] = jclamp(jsar(x0 + t3, 17)); o[7] = jclamp(jsar(x0 - t3, 17));
        o[1] = jclamp(jsar(x1 + t2, 17)); o[6] = jclamp(jsar(x1 - t2, 17));
        o[2] = jclamp(jsar(x2 + t1, 17)); o[5] = jclamp(jsar(x2 - t1, 17));
        o[3] = jclamp(jsar(x3 + t0, 17)); o[4] = jclamp(jsar(x3 - t0, 17));
    }
}

struct JComp { int id = 0, h = 1, v = 1, tq = 0, td = 0, ta = 0, dc_pred = 0; };

bool decode_jpeg(const std::vector<uint8_t> &d, Decoded &out)
{
    if (d.size() < 4 || d[0] != 0xFF || d[1] != 0xD8) return false;
    uint16_t quant[4][64];
    bool have_quant[4] = {false, false, false, false};
    JHuff hdc[4], hac[4];
    JComp comp[3];
    int n_comp = 0, width = 0, height = 0, hmax = 1, vmax = 1, restart_interval = 0;
    int mcu_x = 0, mcu_y = 0;
    bool have_sof = false, rgb_coded = false, decoded_luma = false, progressive = false;
    std::vector<short> coef;  // progressive files: the luma coefficients (64 per block, natural order, not yet dequantised)
    size_t blocks_w = 0, blocks_h = 0;
    size_t y_w2 = 0;
    std::vector<uint8_t> luma;  // component 0, padded to whole MCUs
    size_t pos = 2;
    auto be16 = [&](size_t at) { return (unsigned)((d[at] << 8) | d[at + 1]); };
    int pending = 0;  // a marker the entropy decoder ran into
    for (;;) {
        int m;
        if (pending) { m = pending; pending = 0; }
        else {
            while (pos < d.size() && d[pos] != 0xFF) ++pos;  // stb_image also hunts for the next 0xFF
            while (pos < d.size() && d[pos] == 0xFF) ++pos;
            if (pos >= d.size()) break;
            m = d[pos++];
            if (m == 0) continue;
        }
        if (m == 0xD9) break;                                     // EOI
        if (m == 0xD8 || (m >= 0xD0 && m <= 0xD7) || m == 0x01) continue;  // no payload
        if (pos + 2 > d.size()) return false;
        const size_t len = be16(pos);
        if (len < 2 || pos + len > d.size()) return false;
        const size_t seg = pos + 2, seg_end = pos + len;
        pos = seg_end;
        if (m == 0xDB) {                                          // DQT
            size_t q = seg;
            while (q < seg_end) {
                const int pq = d[q] >> 4, tq = d[q] & 15;
                ++q;
                if (pq > 1 || tq > 3 || q + (pq ? 128 : 64) > seg_end) return false;
                for (int i = 0; i < 64; ++i) { quant[tq][i] = (uint16_t)(pq ? be16(q + 2 * i) : d[q + i]); }  // kept in zigzag order
                q += pq ? 128 : 64;
                have_quant[tq] = true;
            }
        } else if (m == 0xC4) {                                   // DHT
            size_t q = seg;
            while (q < seg_end) {
                if (q + 17 > seg_end) return false;
                const int tc = d[q] >> 4, th = d[q] & 15;
                if (tc > 1 || th > 3) return false;
                int total = 0;
                for (int i = 0; i < 16; ++i) total += d[q + 1 + i];
                if (total > 256 || q + 17 + total > seg_end) return false;
                if (!(tc ? hac[th] : hdc[th]).build(&d[q + 1], &d[q + 17], total)) return false;
                q += 17 + total;
            }
        } else if (m == 0xDD) {                                   // DRI
            if (len != 4) return false;
            restart_interval = (int)be16(seg);
        } else if (m == 0xEE) {                                   // APP14 "Adobe": colour transform 0 with three components = RGB
            if (len >= 14 && !memcmp(&d[seg], "Adobe", 5) && d[seg + 11] == 0) rgb_coded = true;
        } else if (m == 0xC0 || m == 0xC1 || m == 0xC2) {          // SOF0 / SOF1 / SOF2 (progressive)
            progressive = m == 0xC2;
            if (have_sof || len < 8 || d[seg] != 8) return false;
            height = (int)be16(seg + 1); width = (int)be16(seg + 3); n_comp = d[seg + 5];
            if (height == 0 || width == 0 || (n_comp != 1 && n_comp != 3) || len != (size_t)(8 + 3 * n_comp)) return false;
            for (int i = 0; i < n_comp; ++i) {
                comp[i].id = d[seg + 6 + 3 * i];
                comp[i].h = d[seg + 7 + 3 * i] >> 4; comp[i].v = d[seg + 7 + 3 * i] & 15; comp[i].tq = d[seg + 8 + 3 * i];
                if (comp[i].h < 1 || comp[i].h > 4 || comp[i].v < 1 || comp[i].v > 4 || comp[i].tq > 3) return false;
                if (comp[i].h > hmax) hmax = comp[i].h;
                if (comp[i].v > vmax) vmax = comp[i].v;
            }
            if (n_comp == 3 && comp[0].id == 'R' && comp[1].id == 'G' && comp[2].id == 'B') rgb_coded = true;
            if (comp[0].h != hmax || comp[0].v != vmax) return false;  // a subsampled luma plane would need stb's resampler
            mcu_x = (width + 8 * hmax - 1) / (8 * hmax); mcu_y = (height + 8 * vmax - 1) / (8 * vmax);
            y_w2 = (size_t)mcu_x * hmax * 8;
            luma.assign(y_w2 * ((size_t)mcu_y * vmax * 8), 0);
            blocks_w = y_w2 / 8; blocks_h = (size_t)mcu_y * vmax;
            if (progressive) coef.assign(blocks_w * blocks_h * 64, 0);
            have_sof = true;
        } else if (m >= 0xC3 && m <= 0xCF && m != 0xC4 && m != 0xC8 && m != 0xCC) {
            return false;                                         // lossless, hierarchical, arithmetic: not handled
        } else if (m == 0xDA) {                                   // SOS + the entropy-coded segment
            if (!have_sof || (n_comp == 3 && rgb_coded)) return false;
            const int ns = d[seg];
            if (ns < 1 || ns > n_comp || len != (size_t)(6 + 2 * ns)) return false;
            int order[3];
            for (int i = 0; i < ns; ++i) {
                int which = -1;
                for (int c = 0; c < n_comp; ++c) if (comp[c].id == d[seg + 1 + 2 * i]) which = c;
                if (which < 0) return false;
                comp[which].td = d[seg + 2 + 2 * i] >> 4; comp[which].ta = d[seg + 2 + 2 * i] & 15;
                if (comp[which].td > 3 || comp[which].ta > 3) return false;
                if (!progressive && (!hdc[comp[which].td].present || !hac[comp[which].ta].present || !have_quant[comp[which].tq])) return false;
                order[i] = which;
            }
            const int ss = d[seg + 1 + 2 * ns], se = d[seg + 2 + 2 * ns], ah = d[seg + 3 + 2 * ns] >> 4, al = d[seg + 3 + 2 * ns] & 15;
            if (!progressive && (ss != 0 || se != 63 || ah != 0 || al != 0)) return false;  // spectral selection of a sequential scan
            JBits br{&d[0] + seg_end, &d[0] + d.size()};
            for (int c = 0; c < n_comp; ++c) comp[c].dc_pred = 0;
            if (progressive) {
                // ITU T.81 G.1.2: a scan carries either the DC terms (first pass: value >> al; later passes: one more bit) of its
                // components, or a band ss..se of the AC terms of ONE component (first pass with end-of-band runs, refinement
                // passes that add one bit to the non-zero history and place new +-1 << al values).  Only luma is kept; a chroma
                // AC scan is stepped over without decoding.  Dequantisation and the IDCT follow when all scans are in.
                if (ss > 63 || se > 63 || ss > se || ah > 13 || al > 13 || (ss == 0 && se != 0) || (ss != 0 && ns != 1)) return false;
                if (ss != 0 && order[0] != 0) {  // chroma AC band: find the marker that ends its entropy-coded segment
                    size_t q = seg_end;
                    while (q + 1 < d.size() && !(d[q] == 0xFF && d[q + 1] != 0 && d[q + 1] != 0xFF && !(d[q + 1] >= 0xD0 && d[q + 1] <= 0xD7))) ++q;
                    pos = q;
                    continue;
                }
                auto bit1 = [&]() -> int { br.fill(); const int b = (int)br.peek(1); br.skip(1); return b; };
                auto bits = [&](int k) -> int { if (k == 0) return 0; br.fill(); const int b = (int)br.peek(k); br.skip(k); return b; };
                int eob_run = 0;
                int todo = restart_interval ? restart_interval : 0x7fffffff;
                bool stop = false;
                auto after_mcu = [&]() {
                    if (--todo > 0) return;
                    br.fill();
                    if (!(br.marker >= 0xD0 && br.marker <= 0xD7)) { stop = true; return; }
                    br.reset();
                    for (int c = 0; c < n_comp; ++c) comp[c].dc_pred = 0;
                    eob_run = 0;
                    todo = restart_interval;
                };
                auto dc_block = [&](JComp &c, short *blk) -> bool {  // blk == nullptr: a chroma block, decoded only to stay in step
                    if (ah == 0) {
                        if (!hdc[c.td].present) return false;
                        const int t = br.decode(hdc[c.td]);
                        if (t < 0 || t > 16) return false;
                        const long long dcv = (long long)c.dc_pred + br.receive_extend(t);
                        if (dcv < -32768 || dcv > 32767 || dcv * (1 << al) < -32768 || dcv * (1 << al) > 32767) return false;
                        c.dc_pred = (int)dcv;
                        if (blk) blk[0] = (short)(c.dc_pred * (1 << al));
                    } else if (bit1() && blk) {
                        blk[0] = (short)(blk[0] + (1 << al));
                    }
                    return true;
                };
                auto ac_block = [&](JComp &c, short *blk) -> bool {
                    const JHuff &ac = hac[c.ta];
                    if (!ac.present) return false;
                    if (ah == 0) {  // first pass over the band
                        if (eob_run) { --eob_run; return true; }
                        int k = ss;
                        do {
                            const int rs = br.decode(ac);
                            if (rs < 0) return false;
                            const int r = rs >> 4, sz = rs & 15;
                            if (sz == 0) {
                                if (r < 15) { eob_run = (1 << r) + bits(r) - 1; break; }
                                k += 16;
                            } else {
                                k += r;
                                if (k > 63) return false;
                                const long long v = (long long)br.receive_extend(sz) * (1 << al);
                                if (v < -32768 || v > 32767) return false;
                                blk[jpeg_zigzag[k++]] = (short)v;
                            }
                        } while (k <= se);
                        return true;
                    }
                    const short bit = (short)(1 << al);  // refinement pass
                    auto refine = [&](short *p) { if (bit1() && (*p & bit) == 0) *p = (short)(*p > 0 ? *p + bit : *p - bit); };
                    if (eob_run) {
                        --eob_run;
                        for (int k = ss; k <= se; ++k) { short *p = &blk[jpeg_zigzag[k]]; if (*p != 0) refine(p); }
                        return true;
                    }
                    int k = ss;
                    do {
                        const int rs = br.decode(ac);
                        if (rs < 0) return false;
                        int r = rs >> 4, sz = rs & 15, val = 0;
                        if (sz == 0) {
                            if (r < 15) { eob_run = (1 << r) - 1 + bits(r); r = 64; }  // the rest of the band only gets refined
                        } else {
                            if (sz != 1) return false;
                            val = bit1() ? bit : -bit;
                        }
                        while (k <= se) {  // walk over r still-zero coefficients, refining the non-zero ones passed on the way
                            short *p = &blk[jpeg_zigzag[k++]];
                            if (*p != 0) refine(p);
                            else if (r == 0) { *p = (short)val; break; }
                            else --r;
                        }
                    } while (k <= se);
                    return true;
                };
                if (ns == 1) {
                    JComp &c = comp[order[0]];
                    const int cw = (width * c.h + hmax - 1) / hmax, chh = (height * c.v + vmax - 1) / vmax;
                    const int bw = (cw + 7) >> 3, bh = (chh + 7) >> 3;
                    const bool keep = order[0] == 0;
                    for (int by = 0; by < bh && !stop; ++by)
                        for (int bx = 0; bx < bw && !stop; ++bx) {
                            short *blk = keep ? &coef[((size_t)by * blocks_w + bx) * 64] : nullptr;
                            if (!(ss == 0 ? dc_block(c, blk) : ac_block(c, blk))) return false;
                            after_mcu();
                        }
                    if (keep) decoded_luma = true;
                } else {
                    for (int my = 0; my < mcu_y && !stop; ++my)
                        for (int mx = 0; mx < mcu_x && !stop; ++mx) {
                            for (int i = 0; i < ns; ++i) {
                                JComp &c = comp[order[i]];
                                for (int y = 0; y < c.v; ++y)
                                    for (int x = 0; x < c.h; ++x) {
                                        short *blk = order[i] == 0 ? &coef[((size_t)(my * c.v + y) * blocks_w + (size_t)(mx * c.h + x)) * 64] : nullptr;
                                        if (!dc_block(c, blk)) return false;
                                    }
                                if (order[i] == 0) decoded_luma = true;
                            }
                            after_mcu();
                        }
                }
                br.fill();
                pending = br.marker == 0xD9 && br.p >= br.end ? 0 : br.marker;
                pos = (size_t)(br.p - &d[0]);
                if (br.marker == 0xD9) break;
                continue;
            }
            short blk[64];
            auto decode_block = [&](JComp &c, bool keep) -> bool {
                const JHuff &dc = hdc[c.td], &ac = hac[c.ta];
                const uint16_t *q = quant[c.tq];
                if (keep) memset(blk, 0, sizeof(blk));
                const int t = br.decode(dc);
                if (t < 0 || t > 16) return false;
                const long long dcv = (long long)c.dc_pred + br.receive_extend(t);
                if (dcv < -32768 || dcv > 32767 || dcv * q[0] < -32768 || dcv * q[0] > 32767) return false;  // no valid file gets here
                c.dc_pred = (int)dcv;
                if (keep) blk[0] = (short)(c.dc_pred * q[0]);
                int k = 1;
                do {
                    br.fill();
                    const int fa = ac.fast_ac[br.peek(9)];
                    if (fa) {  // run, size and value from one lookup
                        k += (fa >> 4) & 15;
                        if (k > 63) return false;
                        br.skip(fa & 15);
                        const long long v = (long long)(fa >> 8) * q[k];
                        if (v < -32768 || v > 32767) return false;
                        if (keep) blk[jpeg_zigzag[k]] = (short)v;
                        ++k;
                        continue;
                    }
                    const int rs = br.decode(ac);
                    if (rs < 0) return false;
                    const int r = rs >> 4, sz = rs & 15;
                    if (sz == 0) {
                        if (rs != 0xF0) break;  // end of block
                        k += 16;
                    } else {
                        k += r;
                        if (k > 63) return false;
                        const long long v = (long long)br.receive_extend(sz) * q[k];
                        if (v < -32768 || v > 32767) return false;
                        if (keep) blk[jpeg_zigzag[k]] = (short)v;
                        ++k;
                    }
                } while (k < 64);
                return true;
            };
            int todo = restart_interval ? restart_interval : 0x7fffffff;
            bool stop = false;
            auto after_mcu = [&]() {  // restart intervals: the segment must continue with RSTn, predictors start over
                if (--todo > 0) return;
                br.fill();
                if (!(br.marker >= 0xD0 && br.marker <= 0xD7)) { stop = true; return; }
                br.reset();
                for (int c = 0; c < n_comp; ++c) comp[c].dc_pred = 0;
                todo = restart_interval;
            };
            if (ns == 1) {  // non-interleaved: the component's own blocks in raster order
                JComp &c = comp[order[0]];
                const int cw = (width * c.h + hmax - 1) / hmax, chh = (height * c.v + vmax - 1) / vmax;
                const int bw = (cw + 7) >> 3, bh = (chh + 7) >> 3;
                const bool keep = order[0] == 0;
                for (int by = 0; by < bh && !stop; ++by)
                    for (int bx = 0; bx < bw && !stop; ++bx) {
                        if (!decode_block(c, keep)) return false;
                        if (keep) jpeg_idct_block(&luma[(size_t)by * 8 * y_w2 + (size_t)bx * 8], y_w2, blk);
                        after_mcu();
                    }
                if (keep) decoded_luma = true;
            } else {
                for (int my = 0; my < mcu_y && !stop; ++my)
                    for (int mx = 0; mx < mcu_x && !stop; ++mx) {
                        for (int i = 0; i < ns; ++i) {
                            JComp &c = comp[order[i]];
                            const bool keep = order[i] == 0;
                            for (int y = 0; y < c.v; ++y)
                                for (int x = 0; x < c.h; ++x) {
                                    if (!decode_block(c, keep)) return false;
                                    if (keep) jpeg_idct_block(&luma[((size_t)(my * c.v + y) * 8) * y_w2 + (size_t)(mx * c.h + x) * 8], y_w2, blk);
                                }
                            if (keep) decoded_luma = true;
                        }
                        after_mcu();
                    }
            }
            // continue with the marker the bit reader met (or hunt for the next one behind the consumed bytes)
            br.fill();
            pending = br.marker == 0xD9 && br.p >= br.end ? 0 : br.marker;
            pos = (size_t)(br.p - &d[0]);
            if (br.marker == 0xD9) break;
        }
        // APPn, COM and everything else: skipped
    }
    if (!have_sof || !decoded_luma) return false;
    if (progressive) {  // all scans are in: dequantise (16-bit wrap like the sequential path) and transform the luma blocks
        if (!have_quant[comp[0].tq]) return false;
        const uint16_t *q = quant[comp[0].tq];
        short blk[64];
        for (size_t by = 0; by < blocks_h; ++by)
            for (size_t bx = 0; bx < blocks_w; ++bx) {
                const short *c = &coef[(by * blocks_w + bx) * 64];
                for (int k = 0; k < 64; ++k) blk[jpeg_zigzag[k]] = (short)(c[jpeg_zigzag[k]] * q[k]);
                jpeg_idct_block(&luma[by * 8 * y_w2 + bx * 8], y_w2, blk);
            }
    }
    out.w = width; out.h = height; out.gray.resize((size_t)width * height);
    for (int y = 0; y < height; ++y) memcpy(&out.gray[(size_t)y * width], &luma[(size_t)y * y_w2], (size_t)width);
    return true;
}

bool decode_any(const char *path, Decoded &out)
{
    try {  // a damaged header must not take the process down with std::bad_alloc (also runs on the decode pool's threads)
        std::vector<uint8_t> data;
        if (!read_file(path, data)) return false;
        if (data.size() >= 2 && data[0] == 'P') return decode_pnm(data, out);
        if (data.size() >= 2 && data[0] == 'B' && data[1] == 'M') return decode_bmp(data, out);
        if (data.size() >= 2 && data[0] == 0xFF && data[1] == 0xD8) return decode_jpeg(data, out);
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

// Streaming variant for the decode pool: the pixels go straight into the caller's buffer -- a slot of the pinned
// staging ring -- so a raw 8-bit PGM frame is ONE copy, page cache -> pinned memory (pread), with no heap traffic: a pool that
// allocates two 2 MB vectors per frame spends its time in mmap/munmap under the process-wide mm lock.  Other formats decode into
// the thread's scratch vectors (kept across calls) and are copied in.  Returns 1 = decoded into dst, 0 = the reference's failed
// stbi_load, -1 = decoded but larger than cap: the pixels are in `spill`.  Same acceptance as decode_gray_file (a truncated
// PNM payload fails, ray_voxel.cpp:196-199).
int pvp::decode_gray_file_into(const char *path, uint8_t *dst, size_t cap, int *width, int *height, std::vector<uint8_t> &spill)
{
    if (!path) return 0;
    int fd = open(path, O_RDONLY | O_CLOEXEC);
    if (fd < 0) return 0;
    uint8_t head[64];
    const ssize_t got = pread(fd, head, sizeof(head), 0);
    if (got >= 7 && head[0] == 'P' && head[1] == '5') {   // binary PGM: parse the header from the first bytes
        size_t pos = 2;
        int v[3];
        bool ok = true;
        for (int k = 0; k < 3 && ok; ++k) {
            for (;;) {
                while (pos < (size_t)got && (head[pos] == ' ' || head[pos] == '\t' || head[pos] == '\n' || head[pos] == '\r')) ++pos;
                if (pos < (size_t)got && head[pos] == '#') { ok = false; break; }   // comments: the general decoder handles them
                break;
            }
            if (!ok || pos >= (size_t)got || head[pos] < '0' || head[pos] > '9') { ok = false; break; }
            long x = 0;
            while (pos < (size_t)got && head[pos] >= '0' && head[pos] <= '9') { x = x * 10 + (head[pos] - '0'); if (x > (1 << 30)) { ok = false; break; } ++pos; }
            v[k] = (int)x;
        }
        // the token must have ended inside the 64 bytes read (otherwise the number may continue) and maxval must be one byte per sample
        if (ok && pos < (size_t)got && v[0] > 0 && v[1] > 0 && v[2] > 0 && v[2] <= 255) {
            ++pos;  // the single whitespace byte after maxval
            const size_t npx = (size_t)v[0] * (size_t)v[1];
            *width = v[0]; *height = v[1];
            uint8_t *out = dst;
            int rc = 1;
            if (npx > cap) {
                try { spill.resize(npx); } catch (...) { close(fd); return 0; }
                out = spill.data(); rc = -1;
            }
            size_t done = 0;
            while (done < npx) {
                const ssize_t r = pread(fd, out + done, npx - done, (off_t)(pos + done));
                if (r <= 0) break;
                done += (size_t)r;
            }
            close(fd);
            return done == npx ? rc : 0;   // short payload: decode_pnm refuses it too
        }
    }
    close(fd);
    thread_local Decoded d;   // scratch kept per pool thread
    if (!decode_any(path, d)) return 0;
    *width = d.w; *height = d.h;
    if (d.gray.size() > cap) { spill.swap(d.gray); return -1; }
    memcpy(dst, d.gray.data(), d.gray.size());
    return 1;
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
