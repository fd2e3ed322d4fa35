// rec_batch.cu — K8 / K9: cls + rec batch building and the cls postprocess / 180-degree flip.
//   plan_batches  (host) : ordering + batching rules of ClsProcessor::process / RecProcessor::process
//                          (cls_processor.rs:135-140, rec_processor.rs:223-238) — sizes only, no pixels
//   build_batches (K8)   : ImageHelper::resize_norm_image + concatenate (image_helper.rs:176-209):
//                          thumbnail to 48 x resized_w, px/255, (v-.5)/.5, RGB planes, zero pad to img_w,
//                          written straight into the [n,3,48,img_w] batch tensor; rec reads the crop through
//                          its flip flag (rotate_180_in_place as an index transform: zero bytes moved)
//   cls_postprocess (K9) : first-max argmax over [n,2], label/score, flip flag when label==180 && score>=thresh
// Roofline: HBM; algorithmic bytes per line = 3*w*h (crop read) + 12*48*img_w (tensor write).
#include "common.cuh"
#include "thumbnail.cuh"

template <bool DIRECT> __device__ __forceinline__ unsigned bb_px(const unsigned char* __restrict__ base, int idx);
template <bool DIRECT>
struct FlipReaderT {   // generic-class reader: RGBX crop words, or the page itself for a direct crop (see bb_px)
    const unsigned char* __restrict__ base;
    unsigned w, h;
    int flip, rstride, rbase;
    __device__ __forceinline__ uchar3 operator()(unsigned x, unsigned y) const {
        if (flip) { x = w - 1 - x; y = h - 1 - y; }
        const unsigned p = bb_px<DIRECT>(base, (int)y * rstride + rbase + (int)x);
        return make_uchar3((unsigned char)(p & 0xFFu), (unsigned char)((p >> 8) & 0xFFu), (unsigned char)((p >> 16) & 0xFFu));
    }
};

// One block = one (line, 128-column chunk); one thread = one output column over all img_h rows.
// Everything that depends only on x is computed once per thread, everything that depends only on y once per block
// (shared memory); the normalisation `(px as f32 / 255 - .5) / .5` (image_helper.rs:200-203) comes from a 256-entry
// table built with the same three correctly-rounded operations.
//
// A line is classified on the host by its two scale ratios (x: crop_w / resized_w, y: crop_h / img_h); with a ratio
// <= 1 every thumbnail window on that axis is a single pixel or a fractional pair, with a ratio >= 1 it is a block of
// >= 1 pixels (thumbnail.cuh).  Each class runs a straight-line loop that evaluates exactly the expression of the
// imageops::thumbnail branch it can meet — no per-pixel branch chain:
//   BB_FF  x <= 1, y <= 1 (rec lines up to 48 px tall): all four branches collapse into the bilinear expression when a
//          1-px block axis is given the weights (1, 0) — x*1 = x, x*0 = 0, 0 + a = a are exact in f32
//   BB_BB  x >= 1, y >= 1 (taller lines): integer block mean (sum + n/2) / n, the division as a multiply-shift
//   BB_BF  x >  1, y <  1 (cls: wide lines squeezed into 192 columns): column sums of the two rows mixed with
//          (1-f)/n and f/n from a per-(row, n) table; rows that hit a source row exactly are 1 x n block means
//   BB_GEN anything else: generic thumbnail_pixel
// u8 -> f32 goes through the 2^23 mantissa trick (PRMT + FADD, exact) and f32 -> u8 truncation through FADD.RZ with
// 2^23 (exact for 0 <= v < 2^23): both stay on the full-rate pipes instead of the quarter-rate conversion unit.
#define BB_COLS 128
#define BB_MAX_H 64
#define BB_FF 0
#define BB_BB 1
#define BB_BF 2
#define BB_GEN 3
#define BB_BB2 4          // BB with every window at most 2 x 2
#define BB_BF4 5          // BF with every column block at most 4 wide
#define BB_BF_MAXN 8      // widest column block of the BF path (table width)
#define BB_BB_MAXN 32     // largest nx*ny of the BB path (multiply-shift exact for n < 64, see magic_div20)
struct ChunkDev { int line, x0; };
struct AxisS {
    int i0, i1;      // the two source indices the window touches (i1 == i0 for a 1-px block)
    int n;           // block length (0 => fractional case between i0 and i1)
    float fract;
};
__device__ __forceinline__ AxisS axis_small(const ThumbAxis a, unsigned size) {
    AxisS r;
    if (a.lo != a.hi) { r.n = (int)(a.hi - a.lo); r.i0 = (int)a.lo; r.i1 = (int)a.hi - 1; r.fract = 0.0f; }
    else { r.n = 0; r.i0 = (int)a.hi - 1; r.i1 = (int)((a.hi > size - 1) ? size - 1 : a.hi); r.fract = a.fract; }
    return r;
}
// per output row, 16 B: pixel offsets of the two source rows (flip applied) + the vertical weights (FF/BF) or the
// first row offset + row count (BB)
struct RowS { int o0, o1; float fv, omv; };

// k23 = 0x4B000000 held in a register (opaque to the optimiser) so that the PRMT selector can be the immediate
__device__ __forceinline__ float u8f(unsigned word, unsigned k23, unsigned sel) {   // exact float of byte `sel & 3` of word
    return __fsub_rn(__uint_as_float(__byte_perm(word, k23, sel)), 8388608.0f);   // (I2F on a byte costs SHF + LOP + I2FP here)
}
__device__ __forceinline__ float u32f(unsigned v) {                   // exact float of v < 2^23
    return __fsub_rn(__uint_as_float(0x4B000000u | v), 8388608.0f);
}
// lut[floor(v)] for 0 <= v < 256 (NumCast truncation): FADD.RZ leaves 0x4B000000 + floor(v) in the register; the shared
// address is (bits << 2) + (lut - (0x4B000000 << 2)) in wrapping 32-bit arithmetic, so no mask is needed
__device__ __forceinline__ float lut_trunc(float v, unsigned lut_biased) {
    float r;
    asm("ld.shared.f32 %0, [%1];" : "=f"(r) : "r"((__float_as_uint(__fadd_rz(v, 8388608.0f)) << 2) + lut_biased));
    return r;
}
// q = (s * M) >> 20 with M = ceil(2^20 / n) equals s / n whenever s * n < 2^20 (here s <= 255.5 n, n <= 32)
__host__ __device__ inline unsigned magic_div20(unsigned n) { return ((1u << 20) + n - 1) / n; }

struct BBShared {
    float lut[256];
    RowS row[BB_MAX_H];
    ThumbAxis ay[BB_MAX_H];                 // BB_GEN only
    float fb[BB_MAX_H][BB_BF_MAXN];         // BB_BF only: (1 - fract) / n and fract / n
    float ft[BB_MAX_H][BB_BF_MAXN];
    unsigned magic[BB_BB_MAXN + 1];
};

// Source pixel access.  DIRECT == false: the crop was materialised by crop_rows_kernel as RGBX words (one aligned load per pixel).
// DIRECT == true: the crop is a pure translation of the page by whole pixels with every bicubic window inside the page
// (CropDev::direct — axis-aligned boxes, the common case for text lines): cubic(p0,p1,p2,p3,0) == p1, so crop pixel (x, y) IS page
// pixel (x + tx, y + ty) and the batch build reads the page itself — the crop is never written nor re-read (K7 and K8 fused).
// A page pixel is 3 unaligned bytes: two aligned words + one funnel shift (byte 3 of the result is garbage; every consumer
// selects bytes 0..2 by PRMT or multiplies byte 3 by zero in DP4A).
template <bool DIRECT>
__device__ __forceinline__ unsigned bb_px(const unsigned char* __restrict__ base, int idx) {
    if (!DIRECT) return __ldg(reinterpret_cast<const unsigned*>(base) + idx);
    const unsigned char* a = base + 3 * (size_t)(unsigned)idx;
    const unsigned* w = reinterpret_cast<const unsigned*>(reinterpret_cast<uintptr_t>(a) & ~uintptr_t(3));
    return __funnelshift_r(__ldg(w), __ldg(w + 1), 8u * (unsigned)(reinterpret_cast<uintptr_t>(a) & 3u));
}
// two horizontally adjacent pixels idx, idx + 1 (DIRECT: 6 bytes out of three aligned words)
template <bool DIRECT>
__device__ __forceinline__ void bb_px2(const unsigned char* __restrict__ base, int idx, unsigned& p0, unsigned& p1) {
    if (!DIRECT) {
        const unsigned* w = reinterpret_cast<const unsigned*>(base) + idx;
        p0 = __ldg(w); p1 = __ldg(w + 1);
        return;
    }
    const unsigned char* a = base + 3 * (size_t)(unsigned)idx;
    const unsigned* w = reinterpret_cast<const unsigned*>(reinterpret_cast<uintptr_t>(a) & ~uintptr_t(3));
    const unsigned ph = (unsigned)(reinterpret_cast<uintptr_t>(a) & 3u);
    const unsigned w0 = __ldg(w), w1 = __ldg(w + 1), w2 = __ldg(w + 2);
    p0 = __funnelshift_r(w0, w1, 8u * ph);
    p1 = __funnelshift_r(ph == 0 ? w0 : w1, ph == 0 ? w1 : w2, 8u * ((ph + 3u) & 3u));
}

// the pixels of source columns c0 and c1 (equal or neighbours; cl = the left one) of the row at index `row`
template <bool DIRECT>
__device__ __forceinline__ void bb_ff_pair(const unsigned char* __restrict__ base, int row, int c0, int c1, int cl, bool swp, bool same, unsigned& pa, unsigned& pb) {
    if (!DIRECT) {
        const unsigned* w = reinterpret_cast<const unsigned*>(base) + row;
        pa = __ldg(w + c0); pb = __ldg(w + c1);
        return;
    }
    unsigned a, b;
    bb_px2<true>(base, row + cl, a, b);
    pa = swp ? b : a;
    pb = same ? pa : (swp ? a : b);
}
// the pixels of columns cl and cl + dx (dx = 0 or 1)
template <bool DIRECT>
__device__ __forceinline__ void bb_bb2_pair(const unsigned char* __restrict__ base, int idx, int dx, unsigned& pa, unsigned& pb) {
    if (!DIRECT) {
        const unsigned* w = reinterpret_cast<const unsigned*>(base) + idx;
        pa = __ldg(w); pb = __ldg(w + dx);
        return;
    }
    bb_px2<true>(base, idx, pa, pb);   // (for dx == 0 the second pixel is masked out of the sums by the caller)
}

template <bool DIRECT>
__device__ __forceinline__ void bb_body(BBShared& sh, const ChunkDev ck, const LineDev ln, const CropDev& c, const unsigned char* __restrict__ crop_pix,
                                        const int flip, int img_h, float* __restrict__ out, const unsigned k23) {
    const unsigned cw = (unsigned)c.w, chh = (unsigned)c.h;
    const int kind = ln.kind;
    // pixel (x, y) of the crop lives at index y * rstride + rbase + x of `pbase` (RGBX words, or 3-byte page pixels)
    const int rstride = DIRECT ? c.page_w : (int)cw;
    const int rbase = DIRECT ? c.ty * c.page_w + c.tx : 0;
    const unsigned char* __restrict__ pbase = DIRECT ? c.page : crop_pix + c.offset;
    for (int v = threadIdx.x; v < 256; v += BB_COLS) sh.lut[v] = __fdiv_rn(__fsub_rn(__fdiv_rn((float)v, 255.0f), 0.5f), 0.5f);
    if (threadIdx.x <= BB_BB_MAXN) sh.magic[threadIdx.x] = threadIdx.x ? magic_div20(threadIdx.x) : 0u;
    const float yr = __fdiv_rn((float)chh, (float)img_h);
    for (int y = threadIdx.x; y < img_h; y += BB_COLS) {
        const ThumbAxis ay = thumb_axis(y, yr, chh);
        RowS r;
        if (kind == BB_BB || kind == BB_BB2) {   // rows [lo, hi): the sums do not depend on the order, so a flip only moves the start
            const int ny = (int)(ay.hi - ay.lo);
            r.o0 = (flip ? (int)chh - (int)ay.hi : (int)ay.lo) * rstride + rbase;
            r.o1 = kind == BB_BB ? ny : r.o0 + (ny - 1) * rstride;   // BB: row count; BB2: offset of the second row (or the first again)
            r.fv = 0.0f; r.omv = 0.0f;
        } else {
            const AxisS a = axis_small(ay, chh);
            const int j0 = flip ? (int)chh - 1 - a.i0 : a.i0, j1 = flip ? (int)chh - 1 - a.i1 : a.i1;
            r.o0 = j0 * rstride + rbase; r.o1 = j1 * rstride + rbase;
            if (kind == BB_FF) { r.fv = a.n ? 0.0f : a.fract; r.omv = a.n ? 1.0f : __fsub_rn(1.0f, a.fract); }   // (1, 0) for a 1-px block row
            else { r.fv = a.n ? -1.0f : a.fract; r.omv = __fsub_rn(1.0f, a.fract); }                              // fv < 0 marks a block row (BF)
            if (kind == BB_BF || kind == BB_BF4)
                for (int k = 0; k < BB_BF_MAXN; ++k) {
                    sh.fb[y][k] = __fdiv_rn(r.omv, (float)(k + 1));
                    sh.ft[y][k] = __fdiv_rn(a.fract, (float)(k + 1));
                }
            if (kind == BB_GEN) sh.ay[y] = ay;
        }
        sh.row[y] = r;
    }
    __syncthreads();
    const int x = ck.x0 + threadIdx.x;
    if (x >= ln.img_w) return;
    const int stride = ln.img_w;
    const size_t plane = (size_t)img_h * stride;
    float* d0 = out + ln.dst_offset + x;
    float* d1 = d0 + plane;
    float* d2 = d1 + plane;
    if (x >= ln.resized_w) {
        for (int y = 0; y < img_h; ++y, d0 += stride, d1 += stride, d2 += stride) { *d0 = 0.0f; *d1 = 0.0f; *d2 = 0.0f; }
        return;
    }
    const unsigned lut_biased = (unsigned)__cvta_generic_to_shared(sh.lut) - (k23 << 2);
    const float xr = __fdiv_rn((float)cw, (float)ln.resized_w);
    const ThumbAxis ax = thumb_axis(x, xr, cw);
    // Pull the block's source tile into L2 before the row loops start.  The loops keep one output row of loads in flight
    // per thread, which hides an L2 hit but not a DRAM miss (ncu: 59 % of the stall samples sat on the first use of the
    // loaded words); every source row is used by some output row, so one prefetch per thread and row — at the thread's
    // first source column, neighbouring threads cover neighbouring columns — touches every sector of the tile.
    // Measured: 0.94 -> 0.88 ms per 256 pages.  (Two rows of loads in flight per thread: no further gain; a cp.async
    // shared-memory tile per band of rows: slower, 1.01 ms.)
    {
        const unsigned first = ax.lo != ax.hi ? ax.lo : ax.hi - 1, last = ax.lo != ax.hi ? ax.hi - 1 : (ax.hi > cw - 1 ? cw - 1 : ax.hi);
        const unsigned char* __restrict__ pf = pbase + (size_t)(DIRECT ? 3 : 4) * (size_t)(rbase + (int)(flip ? cw - 1 - last : first));
        for (unsigned r = 0; r < chh; ++r, pf += (size_t)(DIRECT ? 3 : 4) * rstride) asm volatile("prefetch.global.L2 [%0];" ::"l"(pf));
    }

    if (kind == BB_FF) {
        const AxisS xs = axis_small(ax, cw);
        // the two source columns are the same pixel or neighbours (i1 = i0 + 1; with the flip i1 lies LEFT of i0): one pair load per row
        const int c0 = flip ? (int)cw - 1 - xs.i0 : xs.i0, c1 = flip ? (int)cw - 1 - xs.i1 : xs.i1;
        const int cl = min(c0, c1);
        const bool swp = c1 < c0, same = c1 == c0;
        const float fhu = xs.n ? 0.0f : xs.fract, omfhu = xs.n ? 1.0f : __fsub_rn(1.0f, xs.fract);
        // software pipeline: the source words of row y+1 are in flight while row y is mixed and stored
        RowS r = sh.row[0];
        unsigned p00, p10, p01, p11;
        bb_ff_pair<DIRECT>(pbase, r.o0, c0, c1, cl, swp, same, p00, p10);
        bb_ff_pair<DIRECT>(pbase, r.o1, c0, c1, cl, swp, same, p01, p11);
#pragma unroll 2
        for (int y = 0; y < img_h; ++y, d0 += stride, d1 += stride, d2 += stride) {
            const RowS rn = sh.row[y + 1 < img_h ? y + 1 : y];
            unsigned n00, n10, n01, n11;
            bb_ff_pair<DIRECT>(pbase, rn.o0, c0, c1, cl, swp, same, n00, n10);
            bb_ff_pair<DIRECT>(pbase, rn.o1, c0, c1, cl, swp, same, n01, n11);
            const float f_tr = __fmul_rn(r.fv, fhu), f_tl = __fmul_rn(r.fv, omfhu), f_br = __fmul_rn(r.omv, fhu), f_bl = __fmul_rn(r.omv, omfhu);
#define RT_BL(sel) lut_trunc(__fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(f_br, u8f(p10, k23, sel)), __fmul_rn(f_tr, u8f(p11, k23, sel))), \
                                                 __fmul_rn(f_bl, u8f(p00, k23, sel))), __fmul_rn(f_tl, u8f(p01, k23, sel))), lut_biased)
            *d0 = RT_BL(0x7540u); *d1 = RT_BL(0x7541u); *d2 = RT_BL(0x7542u);
#undef RT_BL
            r = rn; p00 = n00; p10 = n10; p01 = n01; p11 = n11;
        }
        return;
    }
    if (kind == BB_BB2) {
        // windows of 1-2 x 1-2 pixels (scale ratios in [1, 2]): always four loads (a missing column / row repeats the
        // first one and is masked out of the sums), n = 1, 2 or 4 so the division is a shift; row y+1 prefetched
        const int nx = (int)(ax.hi - ax.lo);
        const int cl = flip ? (int)cw - (int)ax.hi : (int)ax.lo;   // left column of the window; the right one (nx == 2) is its neighbour
        const unsigned mx = nx > 1 ? 0xFFFFFFFFu : 0u;
        RowS r = sh.row[0];
        unsigned w00, w10, w01, w11;
        bb_bb2_pair<DIRECT>(pbase, r.o0 + cl, nx - 1, w00, w10);
        bb_bb2_pair<DIRECT>(pbase, r.o1 + cl, nx - 1, w01, w11);
#pragma unroll 2
        for (int y = 0; y < img_h; ++y, d0 += stride, d1 += stride, d2 += stride) {
            const RowS rn = sh.row[y + 1 < img_h ? y + 1 : y];
            unsigned n00, n10, n01, n11;
            bb_bb2_pair<DIRECT>(pbase, rn.o0 + cl, nx - 1, n00, n10);
            bb_bb2_pair<DIRECT>(pbase, rn.o1 + cl, nx - 1, n01, n11);
            const unsigned my = r.o1 != r.o0 ? 0xFFFFFFFFu : 0u;
            const unsigned shn = (mx & 1u) + (my & 1u), h2 = (1u << shn) >> 1;
            w10 &= mx; w01 &= my; w11 &= mx & my;
            unsigned a0 = h2, a1 = h2, a2 = h2;
            a0 = __dp4a(w00, 0x00000001u, a0); a1 = __dp4a(w00, 0x00000100u, a1); a2 = __dp4a(w00, 0x00010000u, a2);
            a0 = __dp4a(w10, 0x00000001u, a0); a1 = __dp4a(w10, 0x00000100u, a1); a2 = __dp4a(w10, 0x00010000u, a2);
            a0 = __dp4a(w01, 0x00000001u, a0); a1 = __dp4a(w01, 0x00000100u, a1); a2 = __dp4a(w01, 0x00010000u, a2);
            a0 = __dp4a(w11, 0x00000001u, a0); a1 = __dp4a(w11, 0x00000100u, a1); a2 = __dp4a(w11, 0x00010000u, a2);
            *d0 = sh.lut[a0 >> shn]; *d1 = sh.lut[a1 >> shn]; *d2 = sh.lut[a2 >> shn];
            r = rn; w00 = n00; w10 = n10; w01 = n01; w11 = n11;
        }
        return;
    }
    if (kind == BB_BB) {
        const int nx = (int)(ax.hi - ax.lo);
        const int cl = flip ? (int)cw - (int)ax.hi : (int)ax.lo;
        for (int y = 0; y < img_h; ++y, d0 += stride, d1 += stride, d2 += stride) {
            const RowS r = sh.row[y];
            const int ny = r.o1;
            int p = r.o0 + cl;
            unsigned a0 = 0, a1 = 0, a2 = 0;
            for (int j = 0; j < ny; ++j, p += rstride)
                for (int i = 0; i < nx; ++i) {
                    const unsigned w = bb_px<DIRECT>(pbase, p + i);
                    a0 = __dp4a(w, 0x00000001u, a0); a1 = __dp4a(w, 0x00000100u, a1); a2 = __dp4a(w, 0x00010000u, a2);
                }
            const unsigned n = (unsigned)(nx * ny), h2 = n >> 1, M = sh.magic[n];
            *d0 = sh.lut[((a0 + h2) * M) >> 20];
            *d1 = sh.lut[((a1 + h2) * M) >> 20];
            *d2 = sh.lut[((a2 + h2) * M) >> 20];
        }
        return;
    }
    if (kind == BB_BF4) {
        // column blocks of at most 4 pixels: a fixed set of 2 x 4 predicated loads per output row, those of row y+1 in
        // flight while row y is summed and mixed (consecutive output rows mostly re-read the same two source rows: L1 hits)
        const int nx = (int)(ax.hi - ax.lo);
        const int s0 = flip ? (int)cw - (int)ax.hi : (int)ax.lo;
        const unsigned h2 = (unsigned)nx >> 1, M = sh.magic[nx];
        const bool e1 = nx > 1, e2 = nx > 2, e3 = nx > 3;
        RowS r = sh.row[0];
#define RT_LDP(i) bb_px<DIRECT>(pbase, (i))
        unsigned p0 = RT_LDP(s0 + r.o0), p1 = e1 ? RT_LDP(s0 + r.o0 + 1) : 0u, p2 = e2 ? RT_LDP(s0 + r.o0 + 2) : 0u, p3 = e3 ? RT_LDP(s0 + r.o0 + 3) : 0u;
        unsigned q0 = RT_LDP(s0 + r.o1), q1 = e1 ? RT_LDP(s0 + r.o1 + 1) : 0u, q2 = e2 ? RT_LDP(s0 + r.o1 + 2) : 0u, q3 = e3 ? RT_LDP(s0 + r.o1 + 3) : 0u;
        for (int y = 0; y < img_h; ++y, d0 += stride, d1 += stride, d2 += stride) {
            const RowS rn = sh.row[y + 1 < img_h ? y + 1 : y];
            const int pn = s0 + rn.o0, qn = s0 + rn.o1;
            const unsigned np0 = RT_LDP(pn), np1 = e1 ? RT_LDP(pn + 1) : 0u, np2 = e2 ? RT_LDP(pn + 2) : 0u, np3 = e3 ? RT_LDP(pn + 3) : 0u;
            const unsigned nq0 = RT_LDP(qn), nq1 = e1 ? RT_LDP(qn + 1) : 0u, nq2 = e2 ? RT_LDP(qn + 2) : 0u, nq3 = e3 ? RT_LDP(qn + 3) : 0u;
            unsigned a0 = 0, a1 = 0, a2 = 0;
#define RT_ACC(w, x0, x1, x2) x0 = __dp4a(w, 0x00000001u, x0); x1 = __dp4a(w, 0x00000100u, x1); x2 = __dp4a(w, 0x00010000u, x2)
            RT_ACC(p0, a0, a1, a2); RT_ACC(p1, a0, a1, a2); RT_ACC(p2, a0, a1, a2); RT_ACC(p3, a0, a1, a2);
            if (r.fv < 0.0f) {           // the window is one source row: 1 x nx block mean (block-uniform branch)
                *d0 = sh.lut[((a0 + h2) * M) >> 20];
                *d1 = sh.lut[((a1 + h2) * M) >> 20];
                *d2 = sh.lut[((a2 + h2) * M) >> 20];
            } else {
                unsigned c0 = 0, c1 = 0, c2 = 0;
                RT_ACC(q0, c0, c1, c2); RT_ACC(q1, c0, c1, c2); RT_ACC(q2, c0, c1, c2); RT_ACC(q3, c0, c1, c2);
                const float fb = sh.fb[y][nx - 1], ft = sh.ft[y][nx - 1];
                *d0 = lut_trunc(__fadd_rn(__fmul_rn(fb, u32f(a0)), __fmul_rn(ft, u32f(c0))), lut_biased);
                *d1 = lut_trunc(__fadd_rn(__fmul_rn(fb, u32f(a1)), __fmul_rn(ft, u32f(c1))), lut_biased);
                *d2 = lut_trunc(__fadd_rn(__fmul_rn(fb, u32f(a2)), __fmul_rn(ft, u32f(c2))), lut_biased);
            }
#undef RT_ACC
            r = rn; p0 = np0; p1 = np1; p2 = np2; p3 = np3; q0 = nq0; q1 = nq1; q2 = nq2; q3 = nq3;
        }
        return;
    }
    if (kind == BB_BF) {
        const int nx = (int)(ax.hi - ax.lo);
        const int s0 = flip ? (int)cw - (int)ax.hi : (int)ax.lo;
        const unsigned h2 = (unsigned)nx >> 1, M = sh.magic[nx];
        for (int y = 0; y < img_h; ++y, d0 += stride, d1 += stride, d2 += stride) {
            const RowS r = sh.row[y];
            const int p = s0 + r.o0, q = s0 + r.o1;
            unsigned a0 = 0, a1 = 0, a2 = 0, c0 = 0, c1 = 0, c2 = 0;
            for (int i = 0; i < nx; ++i) {
                const unsigned w = RT_LDP(p + i), u = RT_LDP(q + i);
                a0 = __dp4a(w, 0x00000001u, a0); a1 = __dp4a(w, 0x00000100u, a1); a2 = __dp4a(w, 0x00010000u, a2);
                c0 = __dp4a(u, 0x00000001u, c0); c1 = __dp4a(u, 0x00000100u, c1); c2 = __dp4a(u, 0x00010000u, c2);
            }
            if (r.fv < 0.0f) {           // the window is one source row: 1 x nx block mean (block-uniform branch)
                *d0 = sh.lut[((a0 + h2) * M) >> 20];
                *d1 = sh.lut[((a1 + h2) * M) >> 20];
                *d2 = sh.lut[((a2 + h2) * M) >> 20];
            } else {
                const float fb = sh.fb[y][nx - 1], ft = sh.ft[y][nx - 1];
                *d0 = lut_trunc(__fadd_rn(__fmul_rn(fb, u32f(a0)), __fmul_rn(ft, u32f(c0))), lut_biased);
                *d1 = lut_trunc(__fadd_rn(__fmul_rn(fb, u32f(a1)), __fmul_rn(ft, u32f(c1))), lut_biased);
                *d2 = lut_trunc(__fadd_rn(__fmul_rn(fb, u32f(a2)), __fmul_rn(ft, u32f(c2))), lut_biased);
            }
        }
        return;
    }
    {
        const FlipReaderT<DIRECT> rd{pbase, cw, chh, flip, rstride, rbase};
        for (int y = 0; y < img_h; ++y, d0 += stride, d1 += stride, d2 += stride) {
            unsigned char px[3];
            thumbnail_pixel(rd, cw, chh, ax, sh.ay[y], px);
            *d0 = sh.lut[px[0]]; *d1 = sh.lut[px[1]]; *d2 = sh.lut[px[2]];
        }
    }
}
#undef RT_LDP

__global__ void __launch_bounds__(BB_COLS, 8) build_batches_kernel(const LineDev* __restrict__ lines, const ChunkDev* __restrict__ chunks,
                                                                 const CropDev* __restrict__ crops, const unsigned char* __restrict__ crop_pix,
                                                                 const int* __restrict__ flip_flags, int use_flip, int img_h,
                                                                 float* __restrict__ out, const unsigned k23, int allow_direct) {
    // k23 = 0x4B000000 comes in as a kernel parameter: a literal would be folded by ptxas into the PRMT immediate slot,
    // which pushes the byte selector into a register that has to be re-materialised for every conversion
    __shared__ BBShared sh;
    const ChunkDev ck = chunks[blockIdx.x];
    const LineDev ln = lines[ck.line];
    const CropDev& c = crops[ln.crop];
    const int flip = use_flip ? flip_flags[ln.crop] : 0;
    if (allow_direct && c.direct) bb_body<true>(sh, ck, ln, c, crop_pix, flip, img_h, out, k23);   // block-uniform
    else bb_body<false>(sh, ck, ln, c, crop_pix, flip, img_h, out, k23);
}

// ---- K9 ------------------------------------------------------------------------------------------------
struct ClsLine { const float* logits; int crop; int pad; };
__global__ void cls_post_kernel(const ClsLine* __restrict__ lines, int n, int ncls, int label0, int label1, float thresh,
                                int* __restrict__ flip_flags, int* __restrict__ out_label, float* __restrict__ out_score,
                                int* __restrict__ nan_flag) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float* r = lines[i].logits;
    int best = 0;
    float bv = r[0];
    bool nan = bv != bv;
    for (int k = 1; k < ncls; ++k) {
        const float v = r[k];
        nan |= (v != v);
        if (v > bv) { bv = v; best = k; }
    }
    if (nan) { *nan_flag = 1; }
    const int label = best == 0 ? label0 : label1;
    out_label[i] = label;
    out_score[i] = bv;
    // cls_processor.rs:164-166
    if (label == 180 && bv >= thresh) flip_flags[lines[i].crop] ^= 1;
}

retto_b200_status rt_cls_collect(retto_b200_ctx* ctx, int n, retto_b200_cls_result* h_results);

// ---- host ------------------------------------------------------------------------------------------------
extern "C" retto_b200_status retto_b200_plan_batches(const retto_b200_config* cfg, int32_t kind, const retto_b200_crop_info* crops,
                                                     int32_t n, retto_b200_line_job* h_lines, retto_b200_batch* h_batches,
                                                     int32_t* n_batches, uint64_t* total_floats) {
    if (!cfg || n < 0 || (n > 0 && (!crops || !h_lines || !h_batches)) || !n_batches || !total_floats) return RETTO_B200_ERR_INVALID_ARG;
    const int* shape = kind == 0 ? cfg->cls_image_shape : cfg->rec_image_shape;
    const int img_h = shape[1], img_w_cfg = shape[2];
    const int batch_num = std::max(1, kind == 0 ? cfg->cls_batch_num : cfg->rec_batch_num);
    // stable sort by Reverse(OrderedFloat(ori_ratio)), ori_ratio = h as f64 / w as f64 (image_helper.rs:79-82)
    std::vector<int> idx(n);
    for (int i = 0; i < n; ++i) idx[i] = i;
    std::stable_sort(idx.begin(), idx.end(), [&](int a, int b) {
        const double ra = (double)crops[a].h / (double)crops[a].w, rb = (double)crops[b].h / (double)crops[b].w;
        return ra > rb;  // descending; OrderedFloat total order (NaN greatest) is unreachable for w,h > 0
    });
    float max_wh_ratio = (float)img_w_cfg / (float)img_h;  // rec_processor.rs:226-227, carried across batches
    uint64_t off = 0;
    int nb = 0;
    for (int b0 = 0; b0 < n; b0 += batch_num) {
        const int nl = std::min(batch_num, n - b0);
        int img_w = img_w_cfg;
        if (kind == 1) {
            for (int k = 0; k < nl; ++k) {
                const retto_b200_crop_info& c = crops[idx[b0 + k]];
                const float wh = (float)c.w / (float)c.h;  // rec_processor.rs:234-236 (current dims == crop dims)
                if (wh > max_wh_ratio) max_wh_ratio = wh;
            }
            img_w = (int)(size_t)((float)img_h * max_wh_ratio);  // image_helper.rs:179
        }
        retto_b200_batch& bt = h_batches[nb++];
        bt.first_line = b0; bt.n = nl; bt.img_w = img_w; bt.max_wh_ratio = max_wh_ratio; bt.offset = off;
        for (int k = 0; k < nl; ++k) {
            const retto_b200_crop_info& c = crops[idx[b0 + k]];
            const double rw = std::ceil((double)img_h * (double)c.w / (double)c.h);  // image_helper.rs:183
            retto_b200_line_job& l = h_lines[b0 + k];
            l.crop = idx[b0 + k];
            l.img_w = img_w;
            l.resized_w = (int)std::min<size_t>((size_t)img_w, (size_t)rw);
            l.dst_offset = off + (uint64_t)k * 3 * img_h * img_w;
        }
        off += (uint64_t)nl * 3 * img_h * img_w;
    }
    *n_batches = nb;
    *total_floats = off;
    return RETTO_B200_OK;
}

// Host half of build_batches: validate + classify the lines, cut them into 128-column chunks (bucketed by class, the
// slowest class first so its blocks start early and the cheap ones fill the tail of the grid) and upload both tables.
// Only host-known facts are used (crop dims), so the session prepares the cls AND the rec batches while the crop
// kernels are still running; rt_build_batches_launch then only enqueues the kernel.
retto_b200_status rt_build_batches_prepare(retto_b200_ctx* ctx, int32_t kind, const retto_b200_line_job* h_lines, int32_t n_lines,
                                           uint64_t total_floats, float** d_base) {
    DevBuf& buf = kind == 0 ? ctx->d_batch_cls : ctx->d_batch_rec;
    RT_CUDA_OK(ctx, buf.ensure(std::max<size_t>((size_t)total_floats * 4, 16), ctx->stream));
    *d_base = buf.as<float>();
    ctx->bb_chunks[kind] = 0;
    if (n_lines == 0) return RETTO_B200_OK;
    const int img_h = kind == 0 ? ctx->cfg.cls_image_shape[1] : ctx->cfg.rec_image_shape[1];
    if (img_h > BB_MAX_H) { ctx->set_error("build_batches: image_shape height > 64 is not supported"); return RETTO_B200_ERR_UNSUPPORTED; }
    const bool force_generic = getenv("RETTO_B200_BB_GENERIC") != nullptr;   // tests: every line through thumbnail_pixel
    size_t n_chunks[6] = {0, 0, 0, 0, 0, 0};
    std::vector<LineDev>& lines = ctx->bb_lines;
    lines.resize(n_lines);
    for (int i = 0; i < n_lines; ++i) {
        const retto_b200_line_job& l = h_lines[i];
        if (l.crop < 0 || l.crop >= (int)ctx->crops.size() || l.img_w <= 0 || l.resized_w < 0 || l.resized_w > l.img_w ||
            l.dst_offset + (uint64_t)3 * img_h * l.img_w > total_floats || ctx->crops[l.crop].status != RETTO_B200_OK) {
            ctx->set_error("build_batches: bad line " + std::to_string(i));
            return RETTO_B200_ERR_INVALID_ARG;
        }
        // class of the line (see build_batches_kernel): integer comparisons decide on which side of 1 the f32 ratios fall
        // (the division is correctly rounded and monotone), ceil(ratio) bounds every window length on that axis
        const int cw = ctx->crops[l.crop].w, ch = ctx->crops[l.crop].h, rw = std::max(l.resized_w, 1);
        const int nx_max = (cw + rw - 1) / rw, ny_max = (ch + img_h - 1) / img_h;
        int k = BB_GEN;
        if (cw <= rw && ch <= img_h) k = BB_FF;
        else if (cw >= rw && ch >= img_h && nx_max <= 2 && ny_max <= 2) k = BB_BB2;
        else if (cw >= rw && ch >= img_h && nx_max * ny_max <= BB_BB_MAXN) k = BB_BB;
        else if (cw > rw && ch < img_h && nx_max <= 4) k = BB_BF4;
        else if (cw > rw && ch < img_h && nx_max <= BB_BF_MAXN) k = BB_BF;
        if (force_generic) k = BB_GEN;
        lines[i] = LineDev{l.crop, l.img_w, l.resized_w, k, l.dst_offset};
        n_chunks[k] += (size_t)(l.img_w + BB_COLS - 1) / BB_COLS;
    }
    const size_t total_chunks = n_chunks[0] + n_chunks[1] + n_chunks[2] + n_chunks[3] + n_chunks[4] + n_chunks[5];
    if (getenv("RETTO_B200_BB_STATS"))
        fprintf(stderr, "[build_batches kind %d] lines %d chunks FF %zu BB %zu BF %zu GEN %zu BB2 %zu BF4 %zu\n", kind, n_lines, n_chunks[BB_FF], n_chunks[BB_BB],
                n_chunks[BB_BF], n_chunks[BB_GEN], n_chunks[BB_BB2], n_chunks[BB_BF4]);
    const size_t lb = (sizeof(LineDev) * n_lines + 15) & ~size_t(15);
    std::vector<char>& blob = ctx->bb_blob;
    blob.resize(lb + sizeof(ChunkDev) * total_chunks);
    memcpy(blob.data(), lines.data(), sizeof(LineDev) * n_lines);
    ChunkDev* ck = reinterpret_cast<ChunkDev*>(blob.data() + lb);
    size_t cur[6];
    cur[BB_GEN] = 0; cur[BB_BF] = n_chunks[BB_GEN]; cur[BB_BB] = cur[BB_BF] + n_chunks[BB_BF]; cur[BB_BF4] = cur[BB_BB] + n_chunks[BB_BB];
    cur[BB_BB2] = cur[BB_BF4] + n_chunks[BB_BF4]; cur[BB_FF] = cur[BB_BB2] + n_chunks[BB_BB2];
    for (int i = 0; i < n_lines; ++i) {
        size_t& c = cur[lines[i].kind];
        for (int x0 = 0; x0 < lines[i].img_w; x0 += BB_COLS) ck[c++] = ChunkDev{i, x0};
    }
    DevBuf& dl = kind == 0 ? ctx->d_lines : ctx->d_lines_rec;
    RT_TRY(rt_upload(ctx, dl, blob.data(), blob.size()));
    ctx->bb_chunks[kind] = total_chunks;
    ctx->bb_chunk_off[kind] = lb;
    return RETTO_B200_OK;
}

retto_b200_status rt_build_batches_launch(retto_b200_ctx* ctx, int32_t kind) {
    if (ctx->bb_chunks[kind] == 0) return RETTO_B200_OK;
    DevBuf& buf = kind == 0 ? ctx->d_batch_cls : ctx->d_batch_rec;
    DevBuf& dl = kind == 0 ? ctx->d_lines : ctx->d_lines_rec;
    const int img_h = kind == 0 ? ctx->cfg.cls_image_shape[1] : ctx->cfg.rec_image_shape[1];
    const LineDev* d_lines = dl.as<LineDev>();
    const ChunkDev* d_chunks = reinterpret_cast<const ChunkDev*>(dl.as<char>() + ctx->bb_chunk_off[kind]);
    RT_LAUNCH_BEGIN(ctx, "build_batches_kernel");
    build_batches_kernel<<<(unsigned)ctx->bb_chunks[kind], BB_COLS, 0, ctx->stream>>>(d_lines, d_chunks, ctx->d_crop_descs.as<CropDev>(),
                                                                                    ctx->d_crop_pix.as<unsigned char>(), ctx->d_crop_flip.as<int>(),
                                                                                    kind == 1 ? 1 : 0, img_h, buf.as<float>(), 0x4B000000u, ctx->crops_lazy ? 1 : 0);
    RT_LAUNCH_CHECK(ctx);
    ctx->bb_chunks[kind] = 0;
    return RETTO_B200_OK;
}

extern "C" retto_b200_status retto_b200_build_batches(retto_b200_ctx* ctx, int32_t kind, const retto_b200_line_job* h_lines, int32_t n_lines,
                                                      uint64_t total_floats, float** d_base) {
    RtDeviceGuard _dg(ctx);
    if (!ctx || n_lines < 0 || (n_lines > 0 && !h_lines) || !d_base || (kind != 0 && kind != 1)) return RETTO_B200_ERR_INVALID_ARG;
    RT_TRY(rt_build_batches_prepare(ctx, kind, h_lines, n_lines, total_floats, d_base));
    return rt_build_batches_launch(ctx, kind);
}

// enqueue K9 + the result read-back; with defer == true the caller collects after its next stream sync
retto_b200_status rt_cls_postprocess_ptrs(retto_b200_ctx* ctx, const std::vector<const float*>& logits, const int32_t* crop_index, int n,
                                          retto_b200_cls_result* h_results, bool defer) {
    if (n == 0) return RETTO_B200_OK;
    std::vector<ClsLine> lines(n);
    for (int i = 0; i < n; ++i) {
        if (crop_index[i] < 0 || crop_index[i] >= (int)ctx->crops.size()) { ctx->set_error("cls_postprocess: bad crop index"); return RETTO_B200_ERR_INVALID_ARG; }
        lines[i] = ClsLine{logits[i], crop_index[i], 0};
    }
    RT_TRY(rt_upload(ctx, ctx->d_cls_idx, lines.data(), sizeof(ClsLine) * n));
    RT_CUDA_OK(ctx, ctx->d_cls_out.ensure(sizeof(int) * (2 * (size_t)n + 1), ctx->stream));
    int* d_label = ctx->d_cls_out.as<int>();
    float* d_score = reinterpret_cast<float*>(d_label + n);
    int* d_nan = d_label + 2 * n;
    RT_CUDA_OK(ctx, cudaMemsetAsync(d_nan, 0, sizeof(int), ctx->stream));
    RT_LAUNCH_BEGIN(ctx, "cls_post_kernel");
    cls_post_kernel<<<(n + 127) / 128, 128, 0, ctx->stream>>>(ctx->d_cls_idx.as<ClsLine>(), n, 2, ctx->cfg.cls_label[0], ctx->cfg.cls_label[1],
                                                              ctx->cfg.cls_thresh, ctx->d_crop_flip.as<int>(), d_label, d_score, d_nan);
    RT_LAUNCH_CHECK(ctx);
    RT_CUDA_OK(ctx, ctx->h_cls.ensure(sizeof(int) * (2 * (size_t)n + 1)));
    RT_CUDA_OK(ctx, cudaMemcpyAsync(ctx->h_cls.p, d_label, sizeof(int) * (2 * (size_t)n + 1), cudaMemcpyDeviceToHost, ctx->stream));
    if (defer) return RETTO_B200_OK;
    RT_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
    return rt_cls_collect(ctx, n, h_results);
}
// after a stream sync: results of the last rt_cls_postprocess_ptrs
retto_b200_status rt_cls_collect(retto_b200_ctx* ctx, int n, retto_b200_cls_result* h_results) {
    if (n == 0) return RETTO_B200_OK;
    const int* hl = ctx->h_cls.as<int>();
    const float* hs = reinterpret_cast<const float*>(hl + n);
    for (int i = 0; i < n; ++i) { h_results[i].label = hl[i]; h_results[i].score = hs[i]; }
    if (hl[2 * n]) { ctx->set_error("cls_postprocess: NaN logits (reference: argmax().unwrap() panics, cls_processor.rs:113)"); return RETTO_B200_ERR_NAN_LOGITS; }
    return RETTO_B200_OK;
}

extern "C" retto_b200_status retto_b200_cls_postprocess(retto_b200_ctx* ctx, const float* d_logits, int32_t n, const int32_t* h_crop_index,
                                                        retto_b200_cls_result* h_results) {
    RtDeviceGuard _dg(ctx);
    if (!ctx || n < 0 || (n > 0 && (!d_logits || !h_crop_index || !h_results))) return RETTO_B200_ERR_INVALID_ARG;
    std::vector<const float*> ptrs(n);
    for (int i = 0; i < n; ++i) ptrs[i] = d_logits + 2 * (size_t)i;
    return rt_cls_postprocess_ptrs(ctx, ptrs, h_crop_index, n, h_results, false);
}
