// det_preprocess.cu — K1: DetProcessor::preprocess (det_processor.rs:256-274) as one fused kernel:
//   resize_either/thumbnail (image_helper.rs:150-174) -> rgb2bgr (:211-221) -> normalize
//   (det_processor.rs:151-155: x as f32 * scale, - mean, / std; three separately rounded f32 ops)
//   -> HWC->CHW (:157-160) -> contiguous NCHW (ort_worker.rs:191 as_standard_layout).
// plus the stand-alone batched thumbnail used by ImageHelper::resize_both (image_helper.rs:106-148).
//
// Roofline: HBM.  Algorithmic bytes per page = 3*H*W (u8 in) + 12*H'*W' (f32 out).
// Identity-size fast path (the common case: short side >= 736): each thread owns 16 pixels =
// 48 B = three 128-bit loads, and writes four 128-bit stores to each of the 3 planes.
#include "common.cuh"
#include "thumbnail.cuh"

struct DetPreDev {
    const unsigned char* src; int h, w;
    float* dst; int oh, ow;
    float xr, yr;   // w / ow, h / oh as f32 (thumbnail's x_ratio / y_ratio), computed once on the host: IEEE division, same bits
};
// the page a work unit belongs to: one binary search per block (thread 0), per-thread search only for the units of a
// block that straddle a page boundary
__device__ __forceinline__ int unit_page(const int* __restrict__ unit_prefix, int n, int u, int* s_first) {
    if (threadIdx.x == 0) *s_first = rt_find_segment(unit_prefix, n, blockIdx.x * blockDim.x);
    __syncthreads();
    const int p = *s_first;
    return u < unit_prefix[p + 1] ? p : rt_find_segment(unit_prefix, n, u);
}

struct NormParams {
    float scale, mean[3], stdv[3];  // tensor channel c (B,G,R) uses mean[c], std[c]
};

__device__ __forceinline__ float norm1(unsigned char v, float scale, float mean, float stdv) {
    return __fdiv_rn(__fsub_rn(__fmul_rn((float)v, scale), mean), stdv);
}

// work unit = 512 consecutive pixels (identity resize: pixel index == output index), one warp per unit:
//   3 x LDG.128 per lane, fully coalesced (1536 contiguous bytes per warp) -> shared memory,
//   then per plane 4 x STG.128 per lane, each store instruction writing 512 contiguous bytes.
// unit_prefix[p] = units before page p.  Normalisation comes from 256-entry tables built with the
// reference's three separately rounded f32 operations (bit-identical to per-pixel evaluation).
__global__ void __launch_bounds__(256) det_pre_identity_kernel(const DetPreDev* __restrict__ pages, const int* __restrict__ unit_prefix,
                                                                int n_pages, int total_units, NormParams np) {
    __shared__ float s_lut[3][256];
    __shared__ __align__(16) unsigned s_px[8][384];   // 8 warps x 1536 bytes
    for (int i = threadIdx.x; i < 768; i += blockDim.x) {
        const int c = i >> 8, v = i & 255;
        s_lut[c][v] = norm1((unsigned char)v, np.scale, np.mean[c], np.stdv[c]);
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int u = blockIdx.x * 8 + wib;
    if (u >= total_units) return;
    const int p = rt_find_segment(unit_prefix, n_pages, u);
    const DetPreDev pg = pages[p];
    const size_t pix0 = (size_t)(u - unit_prefix[p]) * 512;
    const size_t plane = (size_t)pg.oh * pg.ow;
    const uint4* src = reinterpret_cast<const uint4*>(pg.src + pix0 * 3);
    uint4* sw = reinterpret_cast<uint4*>(s_px[wib]);
#pragma unroll
    for (int k = 0; k < 3; ++k) sw[k * 32 + lane] = __ldcs(src + k * 32 + lane);
    __syncwarp();
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        // pixels q*128 + lane*4 .. +3 = 12 bytes = words 3*(q*32+lane) .. +2 (bank = 3*lane mod 32: conflict-free)
        const unsigned* w3 = s_px[wib] + 3 * (q * 32 + lane);
        const unsigned w0 = w3[0], w1 = w3[1], w2 = w3[2];
        // bytes: R0 G0 B0 R1 | G1 B1 R2 G2 | B2 R3 G3 B3
        const unsigned char r0 = w0 & 0xff, g0 = (w0 >> 8) & 0xff, b0 = (w0 >> 16) & 0xff, r1 = w0 >> 24;
        const unsigned char g1 = w1 & 0xff, b1 = (w1 >> 8) & 0xff, r2 = (w1 >> 16) & 0xff, g2 = w1 >> 24;
        const unsigned char b2 = w2 & 0xff, r3 = (w2 >> 8) & 0xff, g3 = (w2 >> 16) & 0xff, b3 = w2 >> 24;
        const size_t o = pix0 + q * 128 + lane * 4;
        __stcs(reinterpret_cast<float4*>(pg.dst + o), make_float4(s_lut[0][b0], s_lut[0][b1], s_lut[0][b2], s_lut[0][b3]));
        __stcs(reinterpret_cast<float4*>(pg.dst + plane + o), make_float4(s_lut[1][g0], s_lut[1][g1], s_lut[1][g2], s_lut[1][g3]));
        __stcs(reinterpret_cast<float4*>(pg.dst + 2 * plane + o), make_float4(s_lut[2][r0], s_lut[2][r1], s_lut[2][r2], s_lut[2][r3]));
    }
}

// general path: one thread per 4 consecutive output pixels of one row (out_w is a multiple of 32)
__global__ void __launch_bounds__(256) det_pre_resize_kernel(const DetPreDev* __restrict__ pages, const int* __restrict__ unit_prefix,
                                                              int n_pages, int total_units, NormParams np) {
    // normalisation table: the reference's three separately rounded f32 operations per (channel, byte value)
    __shared__ float s_lut[3][256];
    for (int i = threadIdx.x; i < 768; i += blockDim.x) {
        const int c = i >> 8, v = i & 255;
        s_lut[c][v] = norm1((unsigned char)v, np.scale, np.mean[c], np.stdv[c]);
    }
    __shared__ int s_first;
    const int u = blockIdx.x * blockDim.x + threadIdx.x;
    const int p = unit_page(unit_prefix, n_pages, min(u, total_units - 1), &s_first);   // (also the barrier for s_lut)
    if (u >= total_units) return;
    const DetPreDev pg = pages[p];
    const int lu = u - unit_prefix[p];
    const int upr = pg.ow >> 2;
    const int oy = lu / upr, ox0 = (lu - oy * upr) << 2;
    const float xr = pg.xr, yr = pg.yr;
    const ThumbAxis ay = thumb_axis(oy, yr, (unsigned)pg.h);
    const PlainReader rd{pg.src, (unsigned)pg.w};
    const size_t plane = (size_t)pg.oh * pg.ow;
    float o[3][4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const ThumbAxis ax = thumb_axis(ox0 + i, xr, (unsigned)pg.w);
        unsigned char px[3];
        thumbnail_pixel(rd, (unsigned)pg.w, (unsigned)pg.h, ax, ay, px);
        o[0][i] = s_lut[0][px[2]];
        o[1][i] = s_lut[1][px[1]];
        o[2][i] = s_lut[2][px[0]];
    }
    const size_t off = (size_t)oy * pg.ow + ox0;
#pragma unroll
    for (int c = 0; c < 3; ++c)
        __stcs(reinterpret_cast<float4*>(pg.dst + c * plane + off), make_float4(o[c][0], o[c][1], o[c][2], o[c][3]));
}

// Resizes with both ratios <= 2 (every window at most 2 x 2 — everything resize_either produces from a resize_both-ed page:
// up-scaling of a short side below limit_side_len, and the slight up / down scaling of each axis to its multiple of 32),
// laid out like build_batches (rec_batch.cu): one block = 128 output columns x 64 output rows of one page, one thread =
// one output COLUMN.  Everything that depends only on x is computed once per thread, everything that depends only on y
// once per block (shared memory), so the per-pixel work is the window arithmetic alone — the generic kernel spends most
// of its ~180 instructions per pixel on axes, unit lookup and addressing (ncu: 65 % issue utilisation).  The four
// branches of imageops::thumbnail are written out for windows of one or two pixels (branch (i) is the integer mean
// (s + n/2) / n with n in {1, 2, 4}); u8 -> f32 through the 2^23 mantissa trick, f32 -> u8 truncation through FADD.RZ.
#define RS_COLS 128
#define RS_ROWS 64
__device__ __forceinline__ float dp_u8f(unsigned b) { return __fsub_rn(__uint_as_float(0x4B000000u | b), 8388608.0f); }   // exact float of b < 2^23
struct RsRow { unsigned o0, o1; float f, omf, f2, omf2; int kind, pad; };   // kind 0: rows (o0, o1) mixed with f; 1 / 2: block of 1 / 2 rows
__global__ void __launch_bounds__(RS_COLS) det_pre_resize_cols_kernel(const DetPreDev* __restrict__ pages, const int* __restrict__ block_prefix,
                                                                      int n_pages, NormParams np) {
    __shared__ float s_lut[3][256];
    __shared__ RsRow s_row[RS_ROWS];
    __shared__ int s_page;
    for (int i = threadIdx.x; i < 768; i += RS_COLS) {
        const int c = i >> 8, v = i & 255;
        s_lut[c][v] = norm1((unsigned char)v, np.scale, np.mean[c], np.stdv[c]);
    }
    if (threadIdx.x == 0) s_page = rt_find_segment(block_prefix, n_pages, (int)blockIdx.x);
    __syncthreads();
    const int p = s_page;
    const DetPreDev pg = pages[p];
    const int nbx = (pg.ow + RS_COLS - 1) / RS_COLS;
    const int lb = (int)blockIdx.x - block_prefix[p];
    const int by = lb / nbx, bx = lb - by * nbx;
    const int y0 = by * RS_ROWS, rows = min(RS_ROWS, pg.oh - y0);
    const unsigned W = (unsigned)pg.w, H = (unsigned)pg.h;
    if ((int)threadIdx.x < rows) {
        const ThumbAxis ay = thumb_axis(y0 + (int)threadIdx.x, pg.yr, H);
        const bool yb = ay.lo != ay.hi;
        const unsigned r0 = yb ? ay.lo : ay.hi - 1, r1 = yb ? ay.hi - 1 : (ay.hi > H - 1 ? H - 1 : ay.hi);
        RsRow r;
        r.o0 = r0 * W * 3u; r.o1 = r1 * W * 3u;
        r.f = ay.fract; r.omf = __fsub_rn(1.0f, ay.fract);
        r.f2 = __fdiv_rn(r.f, 2.0f); r.omf2 = __fdiv_rn(r.omf, 2.0f);
        r.kind = yb ? (ay.hi - ay.lo > 1 ? 2 : 1) : 0; r.pad = 0;
        s_row[threadIdx.x] = r;
    }
    __syncthreads();
    const int x = bx * RS_COLS + (int)threadIdx.x;
    if (x >= pg.ow) return;
    const ThumbAxis ax = thumb_axis(x, pg.xr, W);
    const bool xb = ax.lo != ax.hi;
    const unsigned c0 = xb ? ax.lo : ax.hi - 1, c1 = xb ? ax.hi - 1 : (ax.hi > W - 1 ? W - 1 : ax.hi);
    const bool nx2 = xb && (ax.hi - ax.lo > 1);
    const float fh = ax.fract, omh = __fsub_rn(1.0f, ax.fract);
    const float fh2 = __fdiv_rn(fh, 2.0f), omh2 = __fdiv_rn(omh, 2.0f);
    const unsigned char* __restrict__ sa = pg.src + 3u * c0;
    const unsigned char* __restrict__ sb = pg.src + 3u * c1;
    const size_t plane = (size_t)pg.oh * pg.ow;
    float* dst = pg.dst + (size_t)y0 * pg.ow + x;
    for (int yy = 0; yy < rows; ++yy, dst += pg.ow) {
        const RsRow r = s_row[yy];
        const bool yb = r.kind != 0, ny2 = r.kind == 2;
        const unsigned char* p00 = sa + r.o0;
        const unsigned char* p10 = sb + r.o0;
        const unsigned char* p01 = sa + r.o1;
        const unsigned char* p11 = sb + r.o1;
        unsigned q[3];
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) {
            const unsigned v00 = __ldg(p00 + ch), v10 = __ldg(p10 + ch), v01 = __ldg(p01 + ch), v11 = __ldg(p11 + ch);
            if (xb && yb) {
                const unsigned sh = (nx2 ? 1u : 0u) + (ny2 ? 1u : 0u);
                const unsigned sum = v00 + (nx2 ? v10 : 0u) + (ny2 ? v01 : 0u) + ((nx2 && ny2) ? v11 : 0u);
                q[ch] = (sum + ((1u << sh) >> 1)) >> sh;
            } else {
                float v;
                if (!xb && !yb) {
                    const float f_tr = __fmul_rn(r.f, fh), f_tl = __fmul_rn(r.f, omh), f_br = __fmul_rn(r.omf, fh), f_bl = __fmul_rn(r.omf, omh);
                    v = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(f_br, dp_u8f(v10)), __fmul_rn(f_tr, dp_u8f(v11))), __fmul_rn(f_bl, dp_u8f(v00))),
                                  __fmul_rn(f_tl, dp_u8f(v01)));
                } else if (!xb) {   // columns (c0, c1) mixed, summed over the 1-2 rows of the block: fl = (1 - fx) / ny, fr = fx / ny
                    const float fr = ny2 ? fh2 : fh, fl = ny2 ? omh2 : omh;
                    v = __fadd_rn(__fmul_rn(fl, dp_u8f(v00 + (ny2 ? v01 : 0u))), __fmul_rn(fr, dp_u8f(v10 + (ny2 ? v11 : 0u))));
                } else {            // rows (o0, o1) mixed, summed over the 1-2 columns of the block: fb = (1 - fy) / nx, ft = fy / nx
                    const float ft = nx2 ? r.f2 : r.f, fb = nx2 ? r.omf2 : r.omf;
                    v = __fadd_rn(__fmul_rn(fb, dp_u8f(v00 + (nx2 ? v10 : 0u))), __fmul_rn(ft, dp_u8f(v01 + (nx2 ? v11 : 0u))));
                }
                q[ch] = __float_as_uint(__fadd_rz(v, 8388608.0f)) & 0xFFu;   // NumCast truncation of 0 <= v < 256
            }
        }
        __stcs(dst, s_lut[0][q[2]]);              // tensor planes are B, G, R
        __stcs(dst + plane, s_lut[1][q[1]]);
        __stcs(dst + 2 * plane, s_lut[2][q[0]]);
    }
}

// scalar fallback for output widths that are not a multiple of 4 (only reachable through a direct API
// call with hand-picked dims; resize_either always yields multiples of 32)
__global__ void det_pre_scalar_kernel(DetPreDev pg, NormParams np) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= pg.oh * pg.ow) return;
    const int oy = i / pg.ow, ox = i - oy * pg.ow;
    const float xr = __fdiv_rn((float)pg.w, (float)pg.ow), yr = __fdiv_rn((float)pg.h, (float)pg.oh);
    const PlainReader rd{pg.src, (unsigned)pg.w};
    unsigned char px[3];
    thumbnail_pixel(rd, (unsigned)pg.w, (unsigned)pg.h, thumb_axis(ox, xr, (unsigned)pg.w), thumb_axis(oy, yr, (unsigned)pg.h), px);
    const size_t plane = (size_t)pg.oh * pg.ow;
    pg.dst[i] = norm1(px[2], np.scale, np.mean[0], np.stdv[0]);
    pg.dst[plane + i] = norm1(px[1], np.scale, np.mean[1], np.stdv[1]);
    pg.dst[2 * plane + i] = norm1(px[0], np.scale, np.mean[2], np.stdv[2]);
}

// stand-alone thumbnail (u8 HWC -> u8 HWC): one thread per PXT consecutive output pixels of a row (4 when the output
// width is a multiple of 4 — resize_both always yields multiples of 32 — so the row axis and the index arithmetic are
// shared and the 12 result bytes leave as three aligned words)
struct ResizeDev {
    const unsigned char* src; int h, w;
    unsigned char* dst; int oh, ow;
    float xr, yr;
    int pxt;
};
// Down-scaling with both ratios in [1, 3] (resize_both's reduction of a > max_side_len page: ratio <= 4096 / 2000): every
// window is a BLOCK of 1-3 x 1-3 pixels, so imageops::thumbnail is its branch (i) everywhere — the integer mean
// (s + n/2) / n, the division as a multiply-shift (exact for s * n < 2^20).  Same layout as det_pre_resize_cols_kernel:
// one block = 128 output columns x 64 rows, one thread = one output column, the row windows in shared memory.
struct TcRow { unsigned o0; int ny; };
__global__ void __launch_bounds__(RS_COLS) thumbnail_cols_kernel(const ResizeDev* __restrict__ jobs, const int* __restrict__ block_prefix, int n_jobs) {
    __shared__ TcRow s_row[RS_ROWS];
    __shared__ int s_job;
    if (threadIdx.x == 0) s_job = rt_find_segment(block_prefix, n_jobs, (int)blockIdx.x);
    __syncthreads();
    const int p = s_job;
    const ResizeDev jb = jobs[p];
    const int nbx = (jb.ow + RS_COLS - 1) / RS_COLS;
    const int lb = (int)blockIdx.x - block_prefix[p];
    const int by = lb / nbx, bx = lb - by * nbx;
    const int y0 = by * RS_ROWS, rows = min(RS_ROWS, jb.oh - y0);
    const unsigned W = (unsigned)jb.w, H = (unsigned)jb.h;
    if ((int)threadIdx.x < rows) {
        const ThumbAxis ay = thumb_axis(y0 + (int)threadIdx.x, jb.yr, H);
        s_row[threadIdx.x] = TcRow{ay.lo * W * 3u, (int)(ay.hi - ay.lo)};
    }
    __syncthreads();
    const int xr_ = bx * RS_COLS + (int)threadIdx.x;
    const bool act = xr_ < jb.ow;
    const int x = act ? xr_ : jb.ow - 1;          // lanes past the row stay alive for the shuffles below (they store nothing)
    const ThumbAxis ax = thumb_axis(x, jb.xr, W);
    const unsigned nx = ax.hi - ax.lo;
    const unsigned char* __restrict__ sa = jb.src + 3u * ax.lo;
    const unsigned M1 = ((1u << 20) + nx - 1) / nx, M2 = ((1u << 20) + 2 * nx - 1) / (2 * nx), M3 = ((1u << 20) + 3 * nx - 1) / (3 * nx);
    const unsigned stride = W * 3u;
    // Stores: the three result bytes of a thread sit at a 3-byte stride — as byte stores a warp row is three instructions touching
    // the same three sectors.  When rows are word-aligned (ow % 4 == 0: everything resize_both produces) the four lanes of a quad
    // exchange pixels by shuffle and three of them store one aligned word each: one instruction, 96 contiguous bytes per warp.
    const bool wide = (jb.ow & 3) == 0 && (reinterpret_cast<uintptr_t>(jb.dst) & 3u) == 0;
    const unsigned q4 = threadIdx.x & 3u;
    unsigned char* dst = jb.dst + ((size_t)y0 * jb.ow + x) * 3;
    unsigned* dstw = reinterpret_cast<unsigned*>(jb.dst + ((size_t)y0 * jb.ow + (x & ~3)) * 3) + q4;
    for (int yy = 0; yy < rows; ++yy, dst += (size_t)jb.ow * 3, dstw += (size_t)jb.ow * 3 / 4) {
        const TcRow r = s_row[yy];
        const unsigned char* q = sa + r.o0;
        unsigned s0 = 0, s1 = 0, s2 = 0;
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            if (j < r.ny) {   // block-uniform
#pragma unroll
                for (unsigned i = 0; i < 3; ++i)
                    if (i < nx) { s0 += __ldg(q + 3 * i); s1 += __ldg(q + 3 * i + 1); s2 += __ldg(q + 3 * i + 2); }
            }
            q += stride;
        }
        const unsigned n = nx * (unsigned)r.ny, h2 = n >> 1, M = r.ny == 1 ? M1 : (r.ny == 2 ? M2 : M3);
        const unsigned c0 = ((s0 + h2) * M) >> 20, c1 = ((s1 + h2) * M) >> 20, c2 = ((s2 + h2) * M) >> 20;
        if (wide) {
            const unsigned px = c0 | (c1 << 8) | (c2 << 16);
            const unsigned pn = __shfl_down_sync(0xffffffffu, px, 1);
            if (act && q4 < 3u) *dstw = (px >> (8u * q4)) | (pn << (24u - 8u * q4));
        } else if (act) {
            dst[0] = (unsigned char)c0; dst[1] = (unsigned char)c1; dst[2] = (unsigned char)c2;
        }
    }
}
__global__ void __launch_bounds__(256) thumbnail_kernel(const ResizeDev* __restrict__ jobs, const int* __restrict__ unit_prefix, int n_jobs,
                                                         int total_units) {
    __shared__ int s_first;
    const int u = blockIdx.x * blockDim.x + threadIdx.x;
    const int p = unit_page(unit_prefix, n_jobs, min(u, total_units - 1), &s_first);
    if (u >= total_units) return;
    const ResizeDev jb = jobs[p];
    const int lu = u - unit_prefix[p];
    const PlainReader rd{jb.src, (unsigned)jb.w};
    if (jb.pxt == 4) {
        const int upr = jb.ow >> 2;
        const int oy = lu / upr, ox0 = (lu - oy * upr) << 2;
        const ThumbAxis ay = thumb_axis(oy, jb.yr, (unsigned)jb.h);
        unsigned char px[4][3];
#pragma unroll
        for (int i = 0; i < 4; ++i) thumbnail_pixel(rd, (unsigned)jb.w, (unsigned)jb.h, thumb_axis(ox0 + i, jb.xr, (unsigned)jb.w), ay, px[i]);
        unsigned char* d = jb.dst + ((size_t)oy * jb.ow + ox0) * 3;
        if ((((uintptr_t)d) & 3) == 0) {
            unsigned* dw = reinterpret_cast<unsigned*>(d);
            dw[0] = px[0][0] | (px[0][1] << 8) | (px[0][2] << 16) | ((unsigned)px[1][0] << 24);
            dw[1] = px[1][1] | (px[1][2] << 8) | (px[2][0] << 16) | ((unsigned)px[2][1] << 24);
            dw[2] = px[2][2] | (px[3][0] << 8) | (px[3][1] << 16) | ((unsigned)px[3][2] << 24);
        } else {
#pragma unroll
            for (int i = 0; i < 4; ++i) { d[3 * i] = px[i][0]; d[3 * i + 1] = px[i][1]; d[3 * i + 2] = px[i][2]; }
        }
        return;
    }
    const int oy = lu / jb.ow, ox = lu - oy * jb.ow;
    unsigned char px[3];
    thumbnail_pixel(rd, (unsigned)jb.w, (unsigned)jb.h, thumb_axis(ox, jb.xr, (unsigned)jb.w), thumb_axis(oy, jb.yr, (unsigned)jb.h), px);
    unsigned char* d = jb.dst + (size_t)lu * 3;
    d[0] = px[0]; d[1] = px[1]; d[2] = px[2];
}

// ---- host -----------------------------------------------------------------------------------------
static inline float round_half_away_f(float v) { return roundf(v); }

extern "C" retto_b200_status retto_b200_resize_both_plan(int32_t ori_h, int32_t ori_w, int32_t max_len, int32_t min_len, int32_t dims[4],
                                                         int32_t* n_steps) {
    // image_helper.rs:106-148 — both branches use the ORIGINAL dims for their size math
    if (ori_h <= 0 || ori_w <= 0 || !dims || !n_steps) return RETTO_B200_ERR_INVALID_ARG;
    int n = 0;
    const float h = (float)ori_h, w = (float)ori_w;
    if (std::max(ori_h, ori_w) > max_len) {
        const float scale = (float)max_len / std::max(h, w);
        const uint32_t rh = std::max<uint32_t>((uint32_t)floorf(h * scale) / 32u, 1u) * 32u;
        const uint32_t rw = std::max<uint32_t>((uint32_t)floorf(w * scale) / 32u, 1u) * 32u;
        dims[2 * n] = (int)rh; dims[2 * n + 1] = (int)rw; ++n;
    }
    if (std::min(ori_h, ori_w) < min_len) {
        const float scale = (float)min_len / std::min(h, w);
        const uint32_t rh = (uint32_t)round_half_away_f(floorf(h * scale) / 32.0f) * 32u;
        const uint32_t rw = (uint32_t)round_half_away_f(floorf(w * scale) / 32.0f) * 32u;
        dims[2 * n] = (int)rh; dims[2 * n + 1] = (int)rw; ++n;
    }
    *n_steps = n;
    return RETTO_B200_OK;
}

extern "C" retto_b200_status retto_b200_resize_either_plan(int32_t h, int32_t w, int32_t limit_type, int32_t limit_len, int32_t* out_h,
                                                           int32_t* out_w) {
    // image_helper.rs:150-174
    if (h <= 0 || w <= 0 || !out_h || !out_w) return RETTO_B200_ERR_INVALID_ARG;
    float ratio = 1.0f;
    if (limit_type == 1) { if (std::max(w, h) > limit_len) ratio = (float)limit_len / (float)std::max(w, h); }
    else { if (std::min(w, h) < limit_len) ratio = (float)limit_len / (float)std::min(w, h); }
    *out_h = (int)((uint32_t)round_half_away_f(floorf((float)h * ratio) / 32.0f) * 32u);
    *out_w = (int)((uint32_t)round_half_away_f(floorf((float)w * ratio) / 32.0f) * 32u);
    return RETTO_B200_OK;
}

template <class Dev>
static retto_b200_status upload_with_prefix(retto_b200_ctx* ctx, DevBuf& buf, const std::vector<Dev>& v, const std::vector<int>& prefix,
                                            const Dev** d_v, const int** d_prefix) {
    std::vector<char> blob(v.size() * sizeof(Dev) + prefix.size() * sizeof(int));
    memcpy(blob.data(), v.data(), v.size() * sizeof(Dev));
    memcpy(blob.data() + v.size() * sizeof(Dev), prefix.data(), prefix.size() * sizeof(int));
    RT_TRY(rt_upload(ctx, buf, blob.data(), blob.size()));
    *d_v = buf.as<Dev>();
    *d_prefix = reinterpret_cast<const int*>(buf.as<char>() + v.size() * sizeof(Dev));
    return RETTO_B200_OK;
}

extern "C" retto_b200_status retto_b200_det_preprocess(retto_b200_ctx* ctx, const retto_b200_det_pre_desc* h_descs, int32_t n) {
    RtDeviceGuard _dg(ctx);
    if (!ctx || (!h_descs && n > 0) || n < 0) return RETTO_B200_ERR_INVALID_ARG;
    NormParams np;
    np.scale = ctx->cfg.det_scale;
    for (int c = 0; c < 3; ++c) { np.mean[c] = ctx->cfg.det_mean[c]; np.stdv[c] = ctx->cfg.det_std[c]; }
    std::vector<DetPreDev> ident, rs, cols;
    std::vector<int> ident_pre{0}, rs_pre{0}, cols_pre{0};
    const bool generic_only = getenv("RETTO_B200_DETPRE_GENERIC") != nullptr;   // tests / A-B: every resize through thumbnail_pixel
    for (int i = 0; i < n; ++i) {
        const retto_b200_det_pre_desc& d = h_descs[i];
        if (!d.d_rgb || !d.d_out || d.h <= 0 || d.w <= 0 || d.out_h <= 0 || d.out_w <= 0) {
            ctx->set_error("det_preprocess: bad descriptor " + std::to_string(i));
            return RETTO_B200_ERR_INVALID_ARG;
        }
        const DetPreDev dv{d.d_rgb, d.h, d.w, d.d_out, d.out_h, d.out_w, (float)d.w / (float)d.out_w, (float)d.h / (float)d.out_h};
        const long long px = (long long)d.out_h * d.out_w;
        const bool aligned = ((uintptr_t)d.d_rgb % 16 == 0) && ((uintptr_t)d.d_out % 16 == 0);
        if (d.out_h == d.h && d.out_w == d.w && (px % 512 == 0) && aligned) {
            ident.push_back(dv);
            ident_pre.push_back(ident_pre.back() + (int)(px / 512));
        } else if (!generic_only && d.w <= 2LL * d.out_w && d.h <= 2LL * d.out_h && (long long)d.h * d.w * 3 < 0xffffffffLL &&
                   cols_pre.back() + (long long)((d.out_w + RS_COLS - 1) / RS_COLS) * ((d.out_h + RS_ROWS - 1) / RS_ROWS) < 0x7fffffffLL) {
            cols.push_back(dv);   // both f32 ratios <= 2 (correctly rounded division is monotone): windows of at most 2 x 2
            cols_pre.push_back(cols_pre.back() + ((d.out_w + RS_COLS - 1) / RS_COLS) * ((d.out_h + RS_ROWS - 1) / RS_ROWS));
        } else if (d.out_w % 4 == 0 && ((uintptr_t)d.d_out % 16 == 0)) {
            rs.push_back(dv);
            rs_pre.push_back(rs_pre.back() + (int)(px / 4));
        } else {
            RT_LAUNCH_BEGIN(ctx, "det_pre_scalar_kernel");
            det_pre_scalar_kernel<<<(unsigned)((px + 255) / 256), 256, 0, ctx->stream>>>(dv, np);
            RT_LAUNCH_CHECK(ctx);
        }
    }
    if (!ident.empty()) {
        const DetPreDev* dv; const int* dp;
        RT_TRY(upload_with_prefix(ctx, ctx->d_stage, ident, ident_pre, &dv, &dp));
        const int total = ident_pre.back();
        RT_LAUNCH_BEGIN(ctx, "det_pre_identity_kernel");
        det_pre_identity_kernel<<<(total + 7) / 8, 256, 0, ctx->stream>>>(dv, dp, (int)ident.size(), total, np);
        RT_LAUNCH_CHECK(ctx);
    }
    if (!cols.empty()) {
        const DetPreDev* dv; const int* dp;
        RT_TRY(upload_with_prefix(ctx, ctx->d_stage_cols, cols, cols_pre, &dv, &dp));
        RT_LAUNCH_BEGIN(ctx, "det_pre_resize_cols_kernel");
        det_pre_resize_cols_kernel<<<cols_pre.back(), RS_COLS, 0, ctx->stream>>>(dv, dp, (int)cols.size(), np);
        RT_LAUNCH_CHECK(ctx);
    }
    if (!rs.empty()) {
        const DetPreDev* dv; const int* dp;
        RT_TRY(upload_with_prefix(ctx, ctx->d_stage2, rs, rs_pre, &dv, &dp));
        const int total = rs_pre.back();
        RT_LAUNCH_BEGIN(ctx, "det_pre_resize_kernel");
        det_pre_resize_kernel<<<(total + 255) / 256, 256, 0, ctx->stream>>>(dv, dp, (int)rs.size(), total, np);
        RT_LAUNCH_CHECK(ctx);
    }
    return RETTO_B200_OK;
}

extern "C" retto_b200_status retto_b200_thumbnail(retto_b200_ctx* ctx, const retto_b200_resize_desc* h_descs, int32_t n) {
    RtDeviceGuard _dg(ctx);
    if (!ctx || (!h_descs && n > 0) || n < 0) return RETTO_B200_ERR_INVALID_ARG;
    if (n == 0) return RETTO_B200_OK;
    std::vector<ResizeDev> jobs, cjobs;
    std::vector<int> pre{0}, cpre{0};
    const bool generic_only = getenv("RETTO_B200_DETPRE_GENERIC") != nullptr;
    for (int i = 0; i < n; ++i) {
        const retto_b200_resize_desc& d = h_descs[i];
        if (!d.d_src || !d.d_dst || d.h <= 0 || d.w <= 0 || d.out_h <= 0 || d.out_w <= 0) {
            ctx->set_error("thumbnail: bad descriptor " + std::to_string(i));
            return RETTO_B200_ERR_INVALID_ARG;
        }
        const int pxt = (d.out_w % 4 == 0) ? 4 : 1;
        const ResizeDev jd{d.d_src, d.h, d.w, d.d_dst, d.out_h, d.out_w, (float)d.w / (float)d.out_w, (float)d.h / (float)d.out_h, pxt};
        const long long nblk = (long long)((d.out_w + RS_COLS - 1) / RS_COLS) * ((d.out_h + RS_ROWS - 1) / RS_ROWS);
        if (!generic_only && d.w >= d.out_w && d.h >= d.out_h && d.w <= 3LL * d.out_w && d.h <= 3LL * d.out_h && (long long)d.h * d.w * 3 < 0xffffffffLL &&
            cpre.back() + nblk < 0x7fffffffLL) {
            cjobs.push_back(jd);   // both ratios in [1, 3]: block windows of at most 3 x 3
            cpre.push_back(cpre.back() + (int)nblk);
        } else {
            jobs.push_back(jd);
            pre.push_back(pre.back() + d.out_h * d.out_w / pxt);
        }
    }
    if (!cjobs.empty()) {
        const ResizeDev* dv; const int* dp;
        RT_TRY(upload_with_prefix(ctx, ctx->d_stage_tcols, cjobs, cpre, &dv, &dp));
        RT_LAUNCH_BEGIN(ctx, "thumbnail_cols_kernel");
        thumbnail_cols_kernel<<<cpre.back(), RS_COLS, 0, ctx->stream>>>(dv, dp, (int)cjobs.size());
        RT_LAUNCH_CHECK(ctx);
    }
    if (jobs.empty()) return RETTO_B200_OK;
    const ResizeDev* dv; const int* dp;
    RT_TRY(upload_with_prefix(ctx, ctx->d_stage3, jobs, pre, &dv, &dp));
    const int total = pre.back();
    RT_LAUNCH_BEGIN(ctx, "thumbnail_kernel");
    thumbnail_kernel<<<(total + 255) / 256, 256, 0, ctx->stream>>>(dv, dp, (int)jobs.size(), total);
    RT_LAUNCH_CHECK(ctx);
    return RETTO_B200_OK;
}
