// jpeg_oracle.cpp — CPU ORACLE for the image-decode row.  TEST INFRASTRUCTURE ONLY.
//
// The reference decodes pages with `image::load_from_memory(..).to_rgb8()` (retto-core/src/image_helper.rs:34-44; wire format =
// file bytes, retto-cli/src/main.rs:83-84).  `image 0.25.6` delegates JPEG to the zune-jpeg crate, which is NOT under
// /root/reference (Cargo.lock dependency) — so, like every third-party routine of this path, the reference's exact pixels are
// unpinned.  What IS on this box is libjpeg-turbo (inside Pillow and OpenCV), the de-facto reference decoder, and this file restates
// its default decode path for baseline JPEG so that the CUDA decoder can be checked bit for bit:
//   jdhuff.c   Huffman entropy decoding (restart intervals, byte stuffing, EXTEND)
//   jidctint.c jpeg_idct_islow: 13-bit fixed point, two passes, range-limit table semantics
//   jdsample.c fancy (triangle) up-sampling h2v1 / h2v2 / h1v2 with libjpeg's edge rules, box replication otherwise
//   jdcolor.c  YCbCr -> RGB with the 16-bit fixed-point tables
// PINNED: tests/test_cpu_jpeg.py compares this restatement with Pillow's AND OpenCV's decode (both libjpeg-turbo) on every
// sub-sampling, odd sizes, restart intervals, optimised Huffman tables and quality levels: identical bytes.
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>

namespace jorc {

static const uint8_t ZIGZAG[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48,
                                   41, 34, 27, 20, 13, 6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,
                                   30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};

struct HuffTable {
    bool present = false;
    uint8_t bits[17] = {0};
    uint8_t vals[256] = {0};
    int mincode[17], maxcode[18], valptr[17];
    void build() {   // jdhuff.c jpeg_make_d_derived_tbl (canonical codes)
        int code = 0, k = 0;
        for (int l = 1; l <= 16; ++l) {
            valptr[l] = k;
            mincode[l] = code;
            code += bits[l];
            k += bits[l];
            maxcode[l] = bits[l] ? code - 1 : -1;
            code <<= 1;
        }
        maxcode[17] = 0x7fffffff;
    }
};

struct Component {
    int id = 0, h = 1, v = 1, tq = 0, td = 0, ta = 0;
    int width_in_blocks = 0, height_in_blocks = 0;   // padded to whole MCUs
    int ds_w = 0, ds_h = 0;                          // downsampled_width / height (real samples)
    std::vector<uint8_t> plane;                      // [height_in_blocks*8][width_in_blocks*8]
};

struct BitReader {
    const uint8_t* p;
    const uint8_t* end;
    uint32_t buf = 0;
    int cnt = 0;
    bool hit_marker = false;
    int next_bit() {
        if (cnt == 0) {
            uint8_t b = 0;
            if (!hit_marker && p < end) {
                b = *p;
                if (b == 0xFF) {
                    if (p + 1 < end && p[1] == 0x00) p += 2;
                    else { hit_marker = true; b = 0; }   // marker: feed zeros (jdhuff.c "insufficient data" rule)
                } else ++p;
            }
            buf = b;
            cnt = 8;
        }
        --cnt;
        return (buf >> cnt) & 1;
    }
    int bits(int n) { int v = 0; for (int i = 0; i < n; ++i) v = (v << 1) | next_bit(); return v; }
    void restart() {   // byte-align, skip to just after the next RSTn marker
        cnt = 0; buf = 0; hit_marker = false;
        while (p + 1 < end) {
            if (p[0] == 0xFF && p[1] >= 0xD0 && p[1] <= 0xD7) { p += 2; return; }
            ++p;
        }
        p = end;
    }
};

static int decode_symbol(BitReader& br, const HuffTable& t) {
    int code = 0;
    for (int l = 1; l <= 16; ++l) {
        code = (code << 1) | br.next_bit();
        if (t.maxcode[l] >= 0 && code <= t.maxcode[l] && code >= t.mincode[l]) return t.vals[t.valptr[l] + code - t.mincode[l]];
    }
    return 0;   // corrupt code: libjpeg warns and returns 0
}
static inline int extend(int v, int s) { return s == 0 ? 0 : (v < (1 << (s - 1)) ? v - (1 << s) + 1 : v); }

// ---- jidctint.c jpeg_idct_islow ----------------------------------------------------------------------------------------------
#define CONST_BITS 13
#define PASS1_BITS 2
#define FIX_0_298631336 2446
#define FIX_0_390180644 3196
#define FIX_0_541196100 4433
#define FIX_0_765366865 6270
#define FIX_0_899976223 7373
#define FIX_1_175875602 9633
#define FIX_1_501321110 12299
#define FIX_1_847759065 15137
#define FIX_1_961570560 16069
#define FIX_2_053119869 16819
#define FIX_2_562915447 20995
#define FIX_3_072711026 25172
static inline int64_t descale(int64_t x, int n) { return (x + ((int64_t)1 << (n - 1))) >> n; }
static inline uint8_t idct_range_limit(int64_t x) {   // range_limit[(x) & RANGE_MASK] of the post-IDCT table (jdmaster.c)
    const int i = (int)(x & 1023);
    if (i < 128) return (uint8_t)(128 + i);
    if (i < 512) return 255;
    if (i < 896) return 0;
    return (uint8_t)(i - 896);
}
static void idct_1d(const int64_t in[8], int64_t out[8], int shift) {
    int64_t z2 = in[2], z3 = in[6];
    int64_t z1 = (z2 + z3) * FIX_0_541196100;
    int64_t tmp2 = z1 + z3 * (-FIX_1_847759065);
    int64_t tmp3 = z1 + z2 * FIX_0_765366865;
    z2 = in[0]; z3 = in[4];
    int64_t tmp0 = (z2 + z3) * ((int64_t)1 << CONST_BITS);
    int64_t tmp1 = (z2 - z3) * ((int64_t)1 << CONST_BITS);
    const int64_t tmp10 = tmp0 + tmp3, tmp13 = tmp0 - tmp3, tmp11 = tmp1 + tmp2, tmp12 = tmp1 - tmp2;
    tmp0 = in[7]; tmp1 = in[5]; tmp2 = in[3]; tmp3 = in[1];
    z1 = tmp0 + tmp3; z2 = tmp1 + tmp2; z3 = tmp0 + tmp2;
    int64_t z4 = tmp1 + tmp3;
    const int64_t z5 = (z3 + z4) * FIX_1_175875602;
    tmp0 *= FIX_0_298631336; tmp1 *= FIX_2_053119869; tmp2 *= FIX_3_072711026; tmp3 *= FIX_1_501321110;
    z1 *= -FIX_0_899976223; z2 *= -FIX_2_562915447; z3 *= -FIX_1_961570560; z4 *= -FIX_0_390180644;
    z3 += z5; z4 += z5;
    tmp0 += z1 + z3; tmp1 += z2 + z4; tmp2 += z2 + z3; tmp3 += z1 + z4;
    out[0] = descale(tmp10 + tmp3, shift); out[7] = descale(tmp10 - tmp3, shift);
    out[1] = descale(tmp11 + tmp2, shift); out[6] = descale(tmp11 - tmp2, shift);
    out[2] = descale(tmp12 + tmp1, shift); out[5] = descale(tmp12 - tmp1, shift);
    out[3] = descale(tmp13 + tmp0, shift); out[4] = descale(tmp13 - tmp0, shift);
}
static void idct_islow(const int16_t coef[64], const uint16_t q[64], uint8_t* out, int stride) {
    int64_t ws[64];
    for (int c = 0; c < 8; ++c) {   // pass 1: columns
        int64_t in[8], o[8];
        for (int r = 0; r < 8; ++r) in[r] = (int64_t)coef[r * 8 + c] * q[r * 8 + c];
        idct_1d(in, o, CONST_BITS - PASS1_BITS);
        for (int r = 0; r < 8; ++r) ws[r * 8 + c] = o[r];
    }
    for (int r = 0; r < 8; ++r) {   // pass 2: rows
        int64_t o[8];
        idct_1d(ws + r * 8, o, CONST_BITS + PASS1_BITS + 3);
        for (int c = 0; c < 8; ++c) out[r * stride + c] = idct_range_limit(o[c]);
    }
}

static inline uint8_t clamp255(int v) { return (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v)); }

// status: 0 ok, 1 not a JPEG / truncated, 2 unsupported (progressive, arithmetic, 12-bit, CMYK, RGB, non-interleaved scans, odd sampling)
static int decode(const uint8_t* d, size_t n, std::vector<uint8_t>& rgb, int* H, int* W) {
    if (n < 4 || d[0] != 0xFF || d[1] != 0xD8) return 1;
    uint16_t qt[4][64];
    bool qt_present[4] = {false, false, false, false};
    HuffTable dc[4], ac[4];
    std::vector<Component> comp;
    int X = 0, Y = 0, ri = 0;
    bool adobe = false, jfif = false;
    int adobe_transform = -1;
    size_t pos = 2;
    bool sof = false;
    for (;;) {
        if (pos + 4 > n) return 1;
        if (d[pos] != 0xFF) return 1;
        while (pos < n && d[pos] == 0xFF) ++pos;
        if (pos >= n) return 1;
        const int m = d[pos++];
        if (m == 0xD8 || m == 0x01 || (m >= 0xD0 && m <= 0xD7)) continue;
        if (m == 0xD9) return 1;
        if (pos + 2 > n) return 1;
        const size_t L = ((size_t)d[pos] << 8) | d[pos + 1];
        if (L < 2 || pos + L > n) return 1;
        const uint8_t* s = d + pos + 2;
        const size_t sl = L - 2;
        if (m == 0xDB) {
            size_t i = 0;
            while (i < sl) {
                const int pq = s[i] >> 4, tq = s[i] & 15;
                ++i;
                if (tq > 3 || pq > 1) return 1;
                if (i + (pq ? 128 : 64) > sl) return 1;
                for (int k = 0; k < 64; ++k) {
                    const int v = pq ? ((s[i] << 8) | s[i + 1]) : s[i];
                    i += pq ? 2 : 1;
                    qt[tq][ZIGZAG[k]] = (uint16_t)v;
                }
                qt_present[tq] = true;
            }
        } else if (m == 0xC0 || m == 0xC1) {
            if (sl < 6 || s[0] != 8) return 2;
            Y = (s[1] << 8) | s[2]; X = (s[3] << 8) | s[4];
            const int nf = s[5];
            if (X <= 0 || Y <= 0) return 2;
            if (nf != 1 && nf != 3) return 2;
            if (sl < (size_t)6 + 3 * nf) return 1;
            comp.resize(nf);
            for (int i = 0; i < nf; ++i) {
                comp[i].id = s[6 + 3 * i]; comp[i].h = s[7 + 3 * i] >> 4; comp[i].v = s[7 + 3 * i] & 15; comp[i].tq = s[8 + 3 * i];
                if (comp[i].h < 1 || comp[i].h > 4 || comp[i].v < 1 || comp[i].v > 4 || comp[i].tq > 3) return 1;
            }
            sof = true;
        } else if (m == 0xC2 || m == 0xC3 || (m >= 0xC5 && m <= 0xCF && m != 0xC4 && m != 0xC8 && m != 0xCC)) {
            return 2;   // progressive, lossless, differential, arithmetic
        } else if (m == 0xCC) {
            return 2;
        } else if (m == 0xC4) {
            size_t i = 0;
            while (i < sl) {
                const int tc = s[i] >> 4, th = s[i] & 15;
                ++i;
                if (tc > 1 || th > 3 || i + 16 > sl) return 1;
                HuffTable& t = tc ? ac[th] : dc[th];
                int cnt = 0;
                t.bits[0] = 0;
                for (int k = 1; k <= 16; ++k) { t.bits[k] = s[i + k - 1]; cnt += t.bits[k]; }
                i += 16;
                if (cnt > 256 || i + cnt > sl) return 1;
                memcpy(t.vals, s + i, cnt);
                i += cnt;
                t.present = true;
                t.build();
            }
        } else if (m == 0xDD) {
            if (sl < 2) return 1;
            ri = (s[0] << 8) | s[1];
        } else if (m == 0xE0) {
            if (sl >= 5 && s[0] == 'J' && s[1] == 'F' && s[2] == 'I' && s[3] == 'F' && s[4] == 0) jfif = true;
        } else if (m == 0xEE) {
            if (sl >= 12 && s[0] == 'A' && s[1] == 'd' && s[2] == 'o' && s[3] == 'b' && s[4] == 'e') { adobe = true; adobe_transform = s[11]; }
        } else if (m == 0xDA) {
            if (!sof) return 1;
            const int ns = s[0];
            if (ns != (int)comp.size()) return 2;   // non-interleaved multi-scan files
            if (sl < (size_t)1 + 2 * ns + 3) return 1;
            for (int i = 0; i < ns; ++i) {
                const int cs = s[1 + 2 * i];
                int ci = -1;
                for (size_t k = 0; k < comp.size(); ++k) if (comp[k].id == cs) ci = (int)k;
                if (ci != i) return 2;
                comp[i].td = s[2 + 2 * i] >> 4; comp[i].ta = s[2 + 2 * i] & 15;
                if (comp[i].td > 3 || comp[i].ta > 3) return 1;
            }
            pos += L;
            break;
        }
        pos += L;
    }
    const int nc = (int)comp.size();
    // colour space (jdapimin.c default_decompress_parms): only YCbCr and grayscale are supported
    if (nc == 3) {
        bool ycc = true;
        if (jfif) ycc = true;
        else if (adobe) ycc = adobe_transform != 0;
        else if (comp[0].id == 'R' && comp[1].id == 'G' && comp[2].id == 'B') ycc = false;
        if (!ycc) return 2;
    }
    int max_h = 1, max_v = 1;
    for (auto& c : comp) { if (c.h > max_h) max_h = c.h; if (c.v > max_v) max_v = c.v; }
    if (nc == 1) { comp[0].h = comp[0].v = 1; max_h = max_v = 1; }   // a single-component scan is non-interleaved: 1 block per MCU
    else {
        if (comp[1].h != 1 || comp[1].v != 1 || comp[2].h != 1 || comp[2].v != 1) return 2;
        if (!((max_h == 1 || max_h == 2) && (max_v == 1 || max_v == 2))) return 2;
    }
    const int mcux = (X + 8 * max_h - 1) / (8 * max_h), mcuy = (Y + 8 * max_v - 1) / (8 * max_v);
    for (auto& c : comp) {
        if (!qt_present[c.tq] || !dc[c.td].present || !ac[c.ta].present) return 1;
        c.width_in_blocks = mcux * c.h; c.height_in_blocks = mcuy * c.v;
        c.ds_w = (X * c.h + max_h - 1) / max_h; c.ds_h = (Y * c.v + max_v - 1) / max_v;
        c.plane.assign((size_t)c.width_in_blocks * 8 * c.height_in_blocks * 8, 0);
    }
    // ---- entropy decode + IDCT (jdhuff.c decode_mcu / jdcoefct.c) ----------------------------------------------------------
    BitReader br{d + pos, d + n};
    int pred[3] = {0, 0, 0};
    int since_restart = 0;
    for (int my = 0; my < mcuy; ++my)
        for (int mx = 0; mx < mcux; ++mx) {
            if (ri && since_restart == ri) {
                br.restart();
                pred[0] = pred[1] = pred[2] = 0;
                since_restart = 0;
            }
            ++since_restart;
            for (int ci = 0; ci < nc; ++ci) {
                Component& c = comp[ci];
                for (int by = 0; by < c.v; ++by)
                    for (int bx = 0; bx < c.h; ++bx) {
                        int16_t coef[64];
                        memset(coef, 0, sizeof(coef));
                        int s = decode_symbol(br, dc[c.td]);
                        int diff = s ? extend(br.bits(s), s) : 0;
                        pred[ci] += diff;
                        coef[0] = (int16_t)pred[ci];
                        for (int k = 1; k < 64;) {
                            const int rs = decode_symbol(br, ac[c.ta]);
                            const int r = rs >> 4;
                            s = rs & 15;
                            if (s) {
                                k += r;
                                const int v = extend(br.bits(s), s);
                                if (k < 64) coef[ZIGZAG[k]] = (int16_t)v;
                                ++k;
                            } else {
                                if (r != 15) break;
                                k += 16;
                            }
                        }
                        const int bxx = mx * c.h + bx, byy = my * c.v + by;
                        const int stride = c.width_in_blocks * 8;
                        idct_islow(coef, qt[c.tq], c.plane.data() + (size_t)byy * 8 * stride + bxx * 8, stride);
                    }
            }
        }
    // ---- up-sampling (jdsample.c) + colour conversion (jdcolor.c) ---------------------------------------------------------------
    *H = Y; *W = X;
    rgb.assign((size_t)X * Y * 3, 0);
    if (nc == 1) {
        const int stride = comp[0].width_in_blocks * 8;
        for (int y = 0; y < Y; ++y)
            for (int x = 0; x < X; ++x) { const uint8_t v = comp[0].plane[(size_t)y * stride + x]; uint8_t* o = &rgb[((size_t)y * X + x) * 3]; o[0] = o[1] = o[2] = v; }
        return 0;
    }
    const int hs = max_h / comp[1].h, vs = max_v / comp[1].v;   // chroma expansion factors (1 or 2)
    const Component& cy = comp[0];
    const int ystride = cy.width_in_blocks * 8;
    std::vector<uint8_t> up[2];
    for (int k = 0; k < 2; ++k) {
        const Component& c = comp[1 + k];
        const int st = c.width_in_blocks * 8;
        const int w = c.ds_w, h = c.ds_h;
        const int ow = w * hs, oh = h * vs;
        up[k].assign((size_t)ow * oh, 0);
        auto in = [&](int y, int x) -> int { return c.plane[(size_t)y * st + x]; };
        const bool fancy_h2 = hs == 2 && w > 2;
        if (hs == 1 && vs == 1) {
            for (int y = 0; y < h; ++y) for (int x = 0; x < w; ++x) up[k][(size_t)y * ow + x] = (uint8_t)in(y, x);
        } else if (hs == 2 && vs == 1) {
            for (int y = 0; y < h; ++y) {
                uint8_t* o = &up[k][(size_t)y * ow];
                if (!fancy_h2) { for (int x = 0; x < w; ++x) o[2 * x] = o[2 * x + 1] = (uint8_t)in(y, x); continue; }
                for (int x = 0; x < w; ++x) {
                    const int v = in(y, x);
                    o[2 * x] = x == 0 ? (uint8_t)v : (uint8_t)((v * 3 + in(y, x - 1) + 1) >> 2);
                    o[2 * x + 1] = x == w - 1 ? (uint8_t)v : (uint8_t)((v * 3 + in(y, x + 1) + 2) >> 2);
                }
            }
        } else if (hs == 1 && vs == 2) {   // h1v2 fancy (libjpeg-turbo)
            for (int y = 0; y < h; ++y)
                for (int v = 0; v < 2; ++v) {
                    const int y1 = v == 0 ? (y > 0 ? y - 1 : 0) : (y < h - 1 ? y + 1 : h - 1);
                    const int bias = v == 0 ? 1 : 2;
                    uint8_t* o = &up[k][(size_t)(2 * y + v) * ow];
                    for (int x = 0; x < w; ++x) o[x] = (uint8_t)((in(y, x) * 3 + in(y1, x) + bias) >> 2);
                }
        } else {   // h2v2
            for (int y = 0; y < h; ++y)
                for (int v = 0; v < 2; ++v) {
                    uint8_t* o = &up[k][(size_t)(2 * y + v) * ow];
                    if (!fancy_h2) { for (int x = 0; x < w; ++x) o[2 * x] = o[2 * x + 1] = (uint8_t)in(y, x); continue; }
                    const int y1 = v == 0 ? (y > 0 ? y - 1 : 0) : (y < h - 1 ? y + 1 : h - 1);
                    auto colsum = [&](int x) { return in(y, x) * 3 + in(y1, x); };
                    for (int x = 0; x < w; ++x) {
                        const int t = colsum(x);
                        o[2 * x] = x == 0 ? (uint8_t)((t * 4 + 8) >> 4) : (uint8_t)((t * 3 + colsum(x - 1) + 8) >> 4);
                        o[2 * x + 1] = x == w - 1 ? (uint8_t)((t * 4 + 7) >> 4) : (uint8_t)((t * 3 + colsum(x + 1) + 7) >> 4);
                    }
                }
        }
    }
    const int cw = comp[1].ds_w * hs;
    for (int y = 0; y < Y; ++y)
        for (int x = 0; x < X; ++x) {
            const int yy = cy.plane[(size_t)y * ystride + x];
            const int cb = up[0][(size_t)y * cw + x] - 128, cr = up[1][(size_t)y * cw + x] - 128;
            // jdcolor.c build_ycc_rgb_table: SCALEBITS 16, ONE_HALF 32768
            const int r = yy + (int)((91881 * (int64_t)cr + 32768) >> 16);
            const int g = yy + (int)(((-22554) * (int64_t)cb + 32768 + (-46802) * (int64_t)cr) >> 16);
            const int b = yy + (int)((116130 * (int64_t)cb + 32768) >> 16);
            uint8_t* o = &rgb[((size_t)y * X + x) * 3];
            o[0] = clamp255(r); o[1] = clamp255(g); o[2] = clamp255(b);
        }
    return 0;
}

}  // namespace jorc

extern "C" {
// two-call protocol: out == nullptr -> only the status and dims
int orc_jpeg_decode(const uint8_t* bytes, size_t n, uint8_t* out, size_t cap, int* h, int* w) {
    std::vector<uint8_t> rgb;
    int H = 0, W = 0;
    const int st = jorc::decode(bytes, n, rgb, &H, &W);
    if (st != 0) return st;
    *h = H; *w = W;
    if (out) {
        if (cap < rgb.size()) return 3;
        memcpy(out, rgb.data(), rgb.size());
    }
    return 0;
}
}
