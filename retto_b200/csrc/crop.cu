// crop.cu — K7: ImageHelper::get_crop_img (image_helper.rs:223-249): perspective rotate-crop of each text
// box with imageproc's bicubic warp (white border), then rotate270 when h/w >= 1.5.
//   setup kernel : 1 thread per crop — 8x8 DLT system in f64 (Gaussian elimination with partial pivoting,
//                  same operation order as the oracle), cast to f32, cofactor inverse, class detection
//   warp kernel  : 1 thread per output pixel — inverse map, 4x4 Catmull-Rom taps with the u8 truncation
//                  between the horizontal and the vertical pass, gather through the read-only path
// Roofline: HBM/L2 gather; algorithmic bytes = sum over boxes of 6*w*h (3 B read + 3 B written per pixel).
#include "common.cuh"
#include "db_geom.cuh"  // side_len

static __device__ bool solve8(double A[8][9]) {
    for (int col = 0; col < 8; ++col) {
        int piv = col;
        double best = fabs(A[col][col]);
        for (int r = col + 1; r < 8; ++r) { const double v = fabs(A[r][col]); if (v > best) { best = v; piv = r; } }
        if (best == 0.0) return false;
        if (piv != col) for (int c = 0; c < 9; ++c) { const double t = A[piv][c]; A[piv][c] = A[col][c]; A[col][c] = t; }
        for (int r = col + 1; r < 8; ++r) {
            const double f = __ddiv_rn(A[r][col], A[col][col]);
            if (f == 0.0) continue;
            for (int c = col; c < 9; ++c) A[r][c] = __dsub_rn(A[r][c], __dmul_rn(f, A[col][c]));
        }
    }
    for (int r = 7; r >= 0; --r) {
        double s = A[r][8];
        for (int c = r + 1; c < 8; ++c) s = __dsub_rn(s, __dmul_rn(A[r][c], A[c][8]));
        A[r][8] = __ddiv_rn(s, A[r][r]);
    }
    return true;
}

__global__ void crop_setup_kernel(CropDev* __restrict__ crops, int n, const int* __restrict__ n_dev, int* __restrict__ any_indirect) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (n_dev) n = min(n, *n_dev);   // device-built table: n is the capacity, *n_dev the count (0 on overflow)
    if (i >= n) return;
    CropDev c = crops[i];
    const float* box = c.box;
    // widths / heights (points.rs:125-169), W = max(width_brc, width_tlc), H = max(height_brc, height_tlc)
    const float w_brc = side_len(box[6], box[7], box[4], box[5]);
    const float w_tlc = side_len(box[0], box[1], box[2], box[3]);
    const float h_brc = side_len(box[2], box[3], box[4], box[5]);
    const float h_tlc = side_len(box[0], box[1], box[6], box[7]);
    const float W = fmaxf(w_brc, w_tlc), H = fmaxf(h_brc, h_tlc);
    const double fx[4] = {box[0], box[2], box[4], box[6]}, fy[4] = {box[1], box[3], box[5], box[7]};
    const double tx[4] = {0.0, (double)W, (double)W, 0.0}, ty[4] = {0.0, 0.0, (double)H, (double)H};
    double A[8][9];
    for (int k = 0; k < 4; ++k) {
        double* r0 = A[2 * k];
        double* r1 = A[2 * k + 1];
        r0[0] = 0; r0[1] = 0; r0[2] = 0; r0[3] = -fx[k]; r0[4] = -fy[k]; r0[5] = -1.0;
        r0[6] = __dmul_rn(ty[k], fx[k]); r0[7] = __dmul_rn(ty[k], fy[k]); r0[8] = -ty[k];
        r1[0] = fx[k]; r1[1] = fy[k]; r1[2] = 1.0; r1[3] = 0; r1[4] = 0; r1[5] = 0;
        r1[6] = __dmul_rn(-tx[k], fx[k]); r1[7] = __dmul_rn(-tx[k], fy[k]); r1[8] = tx[k];
    }
    int status = RETTO_B200_OK;
    float inv[9];
    int cls = 2;
    if (!solve8(A)) status = RETTO_B200_ERR_DEGENERATE_QUAD;
    else {
        float t[9];
        for (int k = 0; k < 8; ++k) t[k] = (float)A[k][8];
        t[8] = 1.0f;
        if (fabsf(t[6]) < 1e-10f && fabsf(t[7]) < 1e-10f && fabsf(__fsub_rn(t[8], 1.0f)) < 1e-10f) {
            if (fabsf(__fsub_rn(t[0], 1.0f)) < 1e-10f && fabsf(t[1]) < 1e-10f && fabsf(t[3]) < 1e-10f && fabsf(__fsub_rn(t[4], 1.0f)) < 1e-10f) cls = 0;
            else cls = 1;
        }
        const float t00 = t[0], t01 = t[1], t02 = t[2], t10 = t[3], t11 = t[4], t12 = t[5], t20 = t[6], t21 = t[7], t22 = t[8];
#define M2(a, b, c, d) __fsub_rn(__fmul_rn(a, b), __fmul_rn(c, d))
        const float m00 = M2(t11, t22, t12, t21);
        const float m01 = M2(t10, t22, t12, t20);
        const float m02 = M2(t10, t21, t11, t20);
        const float det = __fadd_rn(__fsub_rn(__fmul_rn(t00, m00), __fmul_rn(t01, m01)), __fmul_rn(t02, m02));
        if (fabsf(det) < 1e-10f) status = RETTO_B200_ERR_DEGENERATE_QUAD;
        else {
            const float m10 = M2(t01, t22, t02, t21);
            const float m11 = M2(t00, t22, t02, t20);
            const float m12 = M2(t00, t21, t01, t20);
            const float m20 = M2(t01, t12, t02, t11);
            const float m21 = M2(t00, t12, t02, t10);
            const float m22 = M2(t00, t11, t01, t10);
#undef M2
            const float iv[9] = {__fdiv_rn(m00, det), __fdiv_rn(-m10, det), __fdiv_rn(m20, det), __fdiv_rn(-m01, det), __fdiv_rn(m11, det),
                                 __fdiv_rn(-m21, det), __fdiv_rn(m02, det), __fdiv_rn(-m12, det), __fdiv_rn(m22, det)};
            for (int k = 0; k < 8; ++k) inv[k] = __fdiv_rn(iv[k], iv[8]);
            inv[8] = 1.0f;
        }
    }
    if (status == RETTO_B200_OK) { for (int k = 0; k < 9; ++k) crops[i].t[k] = inv[k]; }
    crops[i].cls = cls;
    crops[i].status = status;
    // direct crops: translation by whole pixels, not rotated, the 4x4 bicubic window of EVERY pixel inside the page (the conditions of
    // crop_rows_kernel's copy path with no white pixel) — the crop is then exactly the page rectangle at (tx, ty), and the batch build may
    // read the page instead of a materialised copy
    int direct = 0, itx = 0, ity = 0;
    if (status == RETTO_B200_OK && cls == 0 && !c.rot) {
        const float tx = inv[2], ty = inv[5];
        if (tx == floorf(tx) && ty == floorf(ty) && fabsf(tx) < 1e6f && fabsf(ty) < 1e6f) {
            itx = (int)tx; ity = (int)ty;
            if (itx >= 1 && itx + c.w - 1 <= c.page_w - 4 && ity >= 1 && ity + c.h - 1 + 3 < c.page_h) direct = 1;
        }
    }
    crops[i].direct = direct; crops[i].tx = itx; crops[i].ty = ity; crops[i].pad = 0;
    if (!direct && any_indirect) *any_indirect = 1;   // some crop needs crop_rows_kernel even in the session's lazy mode
}

__device__ __forceinline__ unsigned char clamp_u8_trunc(float x) { return x < 255.0f ? (x > 0.0f ? (unsigned char)x : 0) : 255; }
__device__ __forceinline__ float cubic(float p0, float p1, float p2, float p3, float x) {
    // p1 + 0.5 * x * (p2 - p0 + x * (2.0 * p0 - 5.0 * p1 + 4.0 * p2 - p3 + x * (3.0 * (p1 - p2) + p3 - p0)))
    const float a = __fsub_rn(__fadd_rn(__fmul_rn(3.0f, __fsub_rn(p1, p2)), p3), p0);
    const float b = __fadd_rn(__fsub_rn(__fadd_rn(__fsub_rn(__fmul_rn(2.0f, p0), __fmul_rn(5.0f, p1)), __fmul_rn(4.0f, p2)), p3), __fmul_rn(x, a));
    const float c = __fadd_rn(__fsub_rn(p2, p0), __fmul_rn(x, b));
    return __fadd_rn(p1, __fmul_rn(__fmul_rn(0.5f, x), c));
}

// bicubic sample of one output pixel (x, y) of the un-rotated warp (imageproc warp_into + interpolate_bicubic)
__device__ __forceinline__ void crop_pixel(const CropDev& c, int x, int y, unsigned char rgb[3]) {
    const float xf = (float)x, yf = (float)y;
    float px, py;
    if (c.cls == 2) {
        const float d = __fadd_rn(__fadd_rn(__fmul_rn(c.t[6], xf), __fmul_rn(c.t[7], yf)), c.t[8]);
        px = __fdiv_rn(__fadd_rn(__fadd_rn(__fmul_rn(c.t[0], xf), __fmul_rn(c.t[1], yf)), c.t[2]), d);
        py = __fdiv_rn(__fadd_rn(__fadd_rn(__fmul_rn(c.t[3], xf), __fmul_rn(c.t[4], yf)), c.t[5]), d);
    } else if (c.cls == 1) {
        px = __fadd_rn(__fadd_rn(__fmul_rn(c.t[0], xf), __fmul_rn(c.t[1], yf)), c.t[2]);
        py = __fadd_rn(__fadd_rn(__fmul_rn(c.t[3], xf), __fmul_rn(c.t[4], yf)), c.t[5]);
    } else {
        px = __fadd_rn(xf, c.t[2]);
        py = __fadd_rn(yf, c.t[5]);
    }
    rgb[0] = rgb[1] = rgb[2] = 255;
    const float left = __fsub_rn(floorf(px), 1.0f), right = __fadd_rn(left, 4.0f);
    const float top = __fsub_rn(floorf(py), 1.0f), bottom = __fadd_rn(top, 4.0f);
    if (!(left < 0.0f || right >= (float)c.page_w || top < 0.0f || bottom >= (float)c.page_h) && isfinite(px) && isfinite(py)) {
        const float xw = __fsub_rn(px, __fadd_rn(left, 1.0f)), yw = __fsub_rn(py, __fadd_rn(top, 1.0f));
        const unsigned l = (unsigned)left, tp = (unsigned)top;
        float col[3][4];
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const unsigned char* row = c.page + ((size_t)(tp + r) * c.page_w + l) * 3;
            unsigned char v[12];
#pragma unroll
            for (int k = 0; k < 12; ++k) v[k] = __ldg(row + k);
#pragma unroll
            for (int ch = 0; ch < 3; ++ch)
                col[ch][r] = (float)clamp_u8_trunc(cubic((float)v[ch], (float)v[3 + ch], (float)v[6 + ch], (float)v[9 + ch], xw));
        }
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) rgb[ch] = clamp_u8_trunc(cubic(col[ch][0], col[ch][1], col[ch][2], col[ch][3], yw));
    }
}

// One warp per row of the un-rotated warp of one crop (row_prefix over all crops, built on the host from the crop
// dims).  Translation-class crops with an integral offset (axis-aligned boxes — the common case for text lines)
// sample exact pixel centres: both cubic weights are 0 and cubic(p0,p1,p2,p3,0) == p1, so the row is a guarded byte
// copy, fully coalesced on both sides (same white-border rule: any tap of the 4x4 window outside the page -> white).
// Everything else (affine / projective / rotate270) takes the bicubic path, lanes striding the row.
struct CropTotals { int n, rows, overflow, pad; unsigned long long bytes; };   // written by crop_scan_kernel
template <bool VEC>
__global__ void __launch_bounds__(256) crop_rows_kernel(const CropDev* __restrict__ crops, const int* __restrict__ row_prefix, int n_crops,
                                                         int total_rows, unsigned char* __restrict__ pix, const CropTotals* __restrict__ totals, int lazy,
                                                         const int* __restrict__ any_indirect) {
    if (lazy && any_indirect && *any_indirect == 0) return;   // every crop of the batch is read from its page by the batch build: nothing to materialise
    if (totals) {   // device-built descriptor table: the sizes come from the device, the grid from the host's row hint
        if (totals->overflow) return;
        n_crops = totals->n; total_rows = totals->rows;
    }
    const int lane = threadIdx.x & 31;
    const int ru = blockIdx.x * 8 + (threadIdx.x >> 5);
    // the eight rows of a block mostly belong to one crop: one binary search per block (13 dependent loads), then a short walk
    __shared__ int s_k0;
    if (threadIdx.x == 0) s_k0 = rt_find_segment(row_prefix, n_crops, min(blockIdx.x * 8, max(total_rows - 1, 0)));
    __syncthreads();
    if (ru >= total_rows) return;
    int k = s_k0;
    while (k + 1 < n_crops && row_prefix[k + 1] <= ru) ++k;
    const CropDev& c = crops[k];
    if (c.status != RETTO_B200_OK) return;
    if (lazy && c.direct) return;   // session path: the batch build reads the page itself (rec_batch.cu bb_px<true>); retto_b200_crop_fetch materialises on demand
    const int y = ru - row_prefix[k];
    const int w = c.rot ? c.h : c.w, h = c.rot ? c.w : c.h;  // un-rotated warp size
    if (c.cls == 0 && !c.rot && c.t[2] == floorf(c.t[2]) && c.t[5] == floorf(c.t[5]) && fabsf(c.t[2]) < 1e6f && fabsf(c.t[5]) < 1e6f) {
        const int tx = (int)c.t[2], ty = (int)c.t[5];
        const int iy = y + ty;
        const bool row_ok = !(iy - 1 < 0 || iy + 3 >= c.page_h);
        const int xlo = max(0, 1 - tx), xhi = min(w - 1, c.page_w - 4 - tx);   // columns whose 4x4 window is inside the page
        unsigned* dst = reinterpret_cast<unsigned*>(pix + c.offset) + (size_t)y * w;
        const unsigned char* src = c.page + ((size_t)iy * c.page_w + tx) * 3;
        if (VEC) {
            auto one = [&](int xx) {   // scalar pixel (row head / tail around the 16-byte-aligned groups)
                unsigned o = 0xFFFFFFFFu;
                if (row_ok && xx >= xlo && xx <= xhi) { const unsigned char* sp = src + 3 * xx; o = __ldg(sp) | (__ldg(sp + 1) << 8) | (__ldg(sp + 2) << 16) | 0xFF000000u; }
                dst[xx] = o;
            };
            // groups of four pixels whose RGBX words form one aligned 16-byte store; their 12 source bytes are read as four
            // aligned words (the byte phase of the row is the same for every group) and re-cut with funnel shifts / PRMT
            // (0.280 -> 0.265 ms per 256 pages against the byte-load path below, at the same 40 registers; an unrolled
            // variant at 48 registers lost a resident block per SM and was slower than both)
            const int head = min(w, (int)((4u - ((unsigned)((size_t)y * w) & 3u)) & 3u));
            const int G = (w - head) >> 2;
            if (lane < head) one(lane);
            { const int t0 = head + 4 * G; if (t0 + lane < w) one(t0 + lane); }
            const unsigned sh8 = (unsigned)((uintptr_t)src + 3u * (unsigned)head) & 3u;   // (src + 3 x) & 3 for x = head + 4 g
            const unsigned* wsrc = reinterpret_cast<const unsigned*>(src + 3 * head - sh8);
            const unsigned shift = 8u * sh8;
#pragma unroll 1
            for (int g = lane; g < G; g += 32) {
                const int x = head + 4 * g;
                uint4 o = make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu);
                if (row_ok && x + 3 >= xlo && x <= xhi) {
                    const unsigned* wp = wsrc + 3 * g;
                    const unsigned w0 = __ldg(wp), w1 = __ldg(wp + 1), w2 = __ldg(wp + 2), w3 = __ldg(wp + 3);
                    const unsigned v0 = __funnelshift_r(w0, w1, shift), v1 = __funnelshift_r(w1, w2, shift), v2 = __funnelshift_r(w2, w3, shift);
                    const unsigned p0 = v0 | 0xFF000000u, p1 = __byte_perm(v0, v1, 0x0543) | 0xFF000000u, p2 = __byte_perm(v1, v2, 0x0432) | 0xFF000000u,
                                   p3 = (v2 >> 8) | 0xFF000000u;
                    if (x >= xlo && x <= xhi) o.x = p0;
                    if (x + 1 >= xlo && x + 1 <= xhi) o.y = p1;
                    if (x + 2 >= xlo && x + 2 <= xhi) o.z = p2;
                    if (x + 3 >= xlo && x + 3 <= xhi) o.w = p3;
                }
                *reinterpret_cast<uint4*>(dst + x) = o;
            }
            return;
        }
        // four pixels per lane per trip: the twelve byte loads are independent and issued before the first store
        uchar4* dst4 = reinterpret_cast<uchar4*>(dst);
        for (int x = lane; x < w; x += 128) {
            uchar4 o[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int xx = x + 32 * k;
                o[k] = make_uchar4(255, 255, 255, 255);
                if (row_ok && xx >= xlo && xx <= xhi) { const unsigned char* sp = src + 3 * xx; o[k].x = __ldg(sp); o[k].y = __ldg(sp + 1); o[k].z = __ldg(sp + 2); }
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) if (x + 32 * k < w) dst4[x + 32 * k] = o[k];
        }
        return;
    }
    for (int x = lane; x < w; x += 32) {
        unsigned char rgb[3];
        crop_pixel(c, x, y, rgb);
        size_t o;
        if (c.rot) o = (size_t)(w - 1 - x) * h + y;  // rotate270: out(y, w-1-x) = in(x, y), out is h wide
        else o = (size_t)y * w + x;
        reinterpret_cast<uchar4*>(pix + c.offset)[o] = make_uchar4(rgb[0], rgb[1], rgb[2], 255);
    }
}

// ---- descriptor table built on the device (session path) ---------------------------------------------------------
// The host needs the boxes to plan the cls / rec batches, but the GPU does not need the host to start cropping: two small
// kernels turn the packed boxes of the batch into the CropDev table (dims with the same IEEE operations as rt_crop_dims,
// byte offsets and row prefix by a block scan), and the setup / row kernels are enqueued behind it BEFORE the host
// waits for the boxes — the 0.15 ms the host spends on its copy of the table no longer leaves the GPU idle.
// Capacities (table entries, pixel bytes) are those of the grow-only buffers at enqueue time; if the batch does not
// fit, the kernels do nothing and the host — which computes the same sizes — falls back to the host-built table.
struct CropPageDev { const uint8_t* page; int h, w; int pad; };
// pass 1, one thread per box: everything of the descriptor but its offset
__global__ void __launch_bounds__(128) crop_dims_kernel(const retto_b200_box* __restrict__ boxes, const int* __restrict__ box_off, int n_pages,
                                                         const CropPageDev* __restrict__ pages, CropDev* __restrict__ crops, int cap_crops, int max_boxes,
                                                         int* __restrict__ flip_flags) {
    const int n = box_off[n_pages];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (n > cap_crops || n > max_boxes || i >= n) return;
    const float* box = boxes[i].xy;
    const float w_brc = side_len(box[6], box[7], box[4], box[5]);
    const float w_tlc = side_len(box[0], box[1], box[2], box[3]);
    const float h_brc = side_len(box[2], box[3], box[4], box[5]);
    const float h_tlc = side_len(box[0], box[1], box[6], box[7]);
    const float W = fmaxf(w_brc, w_tlc), H = fmaxf(h_brc, h_tlc);
    const unsigned uw = (unsigned)W, uh = (unsigned)H;
    const int rot = (uw > 0) && (__fdiv_rn((float)uh, (float)uw) >= 1.5f);
    int cw = rot ? (int)uh : (int)uw, ch = rot ? (int)uw : (int)uh, status = RETTO_B200_OK;
    const long long px = (long long)cw * ch;
    if (px <= 0 || px > 0x3fffffffLL) { status = RETTO_B200_ERR_DEGENERATE_QUAD; cw = ch = 0; }
    const int pg = rt_find_segment(box_off, n_pages, i);
    CropDev c;
    c.page = pages[pg].page; c.page_h = pages[pg].h; c.page_w = pages[pg].w;
#pragma unroll
    for (int k = 0; k < 8; ++k) c.box[k] = box[k];
#pragma unroll
    for (int k = 0; k < 9; ++k) c.t[k] = 0.0f;
    c.cls = 0; c.w = cw; c.h = ch; c.rot = rot; c.status = status; c.offset = 0;
    c.direct = 0; c.tx = 0; c.ty = 0; c.pad = 0;
    crops[i] = c;
    flip_flags[i] = 0;
}
// pass 2, one block: byte offsets and row prefix (exclusive scans over the boxes), totals
__global__ void __launch_bounds__(1024) crop_scan_kernel(const int* __restrict__ box_off, int n_pages, CropDev* __restrict__ crops, int* __restrict__ row_prefix,
                                                          int cap_crops, unsigned long long cap_bytes, CropTotals* __restrict__ totals, int max_boxes) {
    __shared__ unsigned long long s_b[32];
    __shared__ int s_r[32];
    __shared__ unsigned long long s_carry_b;
    __shared__ int s_carry_r;
    const int n = box_off[n_pages];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (n > cap_crops || n > max_boxes) {   // more boxes than the table (or the packed box array) holds
        if (threadIdx.x == 0) { totals->n = 0; totals->rows = 0; totals->overflow = 1; totals->bytes = 0; }
        return;
    }
    if (threadIdx.x == 0) { s_carry_b = 0; s_carry_r = 0; row_prefix[0] = 0; }
    __syncthreads();
    for (int base = 0; base < n; base += 1024) {
        const int i = base + threadIdx.x;
        int rows = 0;
        unsigned long long bytes = 0;
        if (i < n) {
            const int cw = crops[i].w, ch = crops[i].h;
            rows = crops[i].rot ? cw : ch;
            bytes = ((unsigned long long)cw * ch * 4 + 15) & ~15ULL;
        }
        unsigned long long xb = bytes;
        int xr = rows;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned long long tb = __shfl_up_sync(0xffffffffu, xb, o);
            const int t2 = __shfl_up_sync(0xffffffffu, xr, o);
            if (lane >= o) { xb += tb; xr += t2; }
        }
        if (lane == 31) { s_b[w] = xb; s_r[w] = xr; }
        __syncthreads();
        if (w == 0) {
            unsigned long long tb = s_b[lane];
            int t2 = s_r[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned long long ub = __shfl_up_sync(0xffffffffu, tb, o);
                const int u2 = __shfl_up_sync(0xffffffffu, t2, o);
                if (lane >= o) { tb += ub; t2 += u2; }
            }
            s_b[lane] = tb; s_r[lane] = t2;
        }
        __syncthreads();
        const unsigned long long incl_b = s_carry_b + (w ? s_b[w - 1] : 0ULL) + xb;
        const int incl_r = s_carry_r + (w ? s_r[w - 1] : 0) + xr;
        if (i < n) { crops[i].offset = incl_b - bytes; row_prefix[i + 1] = incl_r; }
        __syncthreads();
        if (threadIdx.x == 1023) { s_carry_b = incl_b; s_carry_r = incl_r; }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const unsigned long long tot = s_carry_b;
        const bool over = tot > cap_bytes;
        totals->n = over ? 0 : n; totals->rows = over ? 0 : s_carry_r; totals->overflow = over ? 1 : 0; totals->bytes = tot;
    }
}

// host: crop dims (same IEEE operations as the setup kernel; size planning only)
static inline float side_len_h(float ax, float ay, float bx, float by) {
    const double dx = (double)(ax - bx), dy = (double)(ay - by);
    return (float)std::sqrt(dx * dx + dy * dy);
}
void rt_crop_dims(const float box[8], int* cw, int* ch, int* rot) {
    const float w_brc = side_len_h(box[6], box[7], box[4], box[5]);
    const float w_tlc = side_len_h(box[0], box[1], box[2], box[3]);
    const float h_brc = side_len_h(box[2], box[3], box[4], box[5]);
    const float h_tlc = side_len_h(box[0], box[1], box[6], box[7]);
    const float W = std::max(w_brc, w_tlc), H = std::max(h_brc, h_tlc);
    const uint32_t w = (uint32_t)W, h = (uint32_t)H;
    const int r = (w > 0) && ((float)h / (float)w >= 1.5f);
    *rot = r;
    *cw = r ? (int)h : (int)w;
    *ch = r ? (int)w : (int)h;
}

// async part: descriptors built in place in a pinned staging slot and uploaded, projection setup + row kernel
// enqueued, status read-back enqueued (no sync).  `get(i, &page, &page_h, &page_w)` returns the box of crop i.
template <class Get>
static retto_b200_status crop_launch_impl(retto_b200_ctx* ctx, int n, retto_b200_crop_info* h_infos, Get get) {
    ctx->crops.clear();
    ctx->crop_dev_check = false;
    if (n == 0) return RETTO_B200_OK;
    ctx->crops.resize(n);
    const size_t desc_bytes = sizeof(CropDev) * (size_t)n, blob_bytes = desc_bytes + sizeof(int) * ((size_t)n + 1);
    int slot = -1;
    void* sp = nullptr;
    RT_TRY(rt_stage_begin(ctx, blob_bytes, &slot, &sp));
    CropDev* hd = reinterpret_cast<CropDev*>(sp);
    int* prefix = reinterpret_cast<int*>(reinterpret_cast<char*>(sp) + desc_bytes);
    prefix[0] = 0;
    unsigned long long off = 0;
    for (int i = 0; i < n; ++i) {
        CropDev& c = hd[i];
        const float* box = get(i, &c.page, &c.page_h, &c.page_w);
        if (!c.page || c.page_h <= 0 || c.page_w <= 0) { ctx->stage_slots[slot].busy = false; ctx->crops.clear(); ctx->set_error("crop_boxes: bad job " + std::to_string(i)); return RETTO_B200_ERR_INVALID_ARG; }
        memcpy(c.box, box, sizeof(float) * 8);
        for (int k = 0; k < 9; ++k) c.t[k] = 0.0f;
        c.cls = 0; c.direct = 0; c.tx = 0; c.ty = 0; c.pad = 0;
        rt_crop_dims(c.box, &c.w, &c.h, &c.rot);
        c.status = RETTO_B200_OK;
        c.offset = off;
        const long long px = (long long)c.w * c.h;
        const int rows = c.rot ? c.w : c.h;
        if (px <= 0 || px > 0x3fffffffLL || prefix[i] + (long long)rows > 0x7fffffffLL) { c.status = RETTO_B200_ERR_DEGENERATE_QUAD; c.w = c.h = 0; prefix[i + 1] = prefix[i]; }
        else { prefix[i + 1] = prefix[i] + rows; off += ((unsigned long long)px * 4 + 15) & ~15ULL; }   // crops are stored RGBX (4 B/px): one aligned word per pixel
        ctx->crops[i] = retto_b200_ctx::CropHost{c.w, c.h, c.rot, c.status, c.offset};
        h_infos[i].w = c.w; h_infos[i].h = c.h; h_infos[i].rotated270 = c.rot; h_infos[i].status = c.status; h_infos[i].offset = c.offset;
    }
    const int rows = prefix[n];
    cudaStream_t st = ctx->stream;
    RT_CUDA_OK(ctx, ctx->d_crop_pix.ensure((size_t)std::max<unsigned long long>(off, 16), st));
    RT_CUDA_OK(ctx, ctx->d_crop_flip.ensure(sizeof(int) * (size_t)n, st));
    RT_CUDA_OK(ctx, cudaMemsetAsync(ctx->d_crop_flip.p, 0, sizeof(int) * (size_t)n, st));
    RT_TRY(rt_stage_commit(ctx, ctx->d_crop_descs, slot, blob_bytes));
    CropDev* d_crops = ctx->d_crop_descs.as<CropDev>();
    const int* d_prefix = reinterpret_cast<const int*>(ctx->d_crop_descs.as<char>() + desc_bytes);
    RT_LAUNCH_BEGIN(ctx, "crop_setup_kernel");
    RT_CUDA_OK(ctx, ctx->d_crop_any.ensure(16, st));
    RT_CUDA_OK(ctx, cudaMemsetAsync(ctx->d_crop_any.p, 0, 4, st));
    crop_setup_kernel<<<(n + 63) / 64, 64, 0, st>>>(d_crops, n, nullptr, ctx->d_crop_any.as<int>());
    RT_LAUNCH_CHECK(ctx);
    if (rows > 0) {
        RT_LAUNCH_BEGIN(ctx, "crop_rows_kernel");
        if (!getenv("RETTO_B200_CROP_SCALAR")) crop_rows_kernel<true><<<(rows + 7) / 8, 256, 0, st>>>(d_crops, d_prefix, n, rows, ctx->d_crop_pix.as<unsigned char>(), nullptr, ctx->crops_lazy ? 1 : 0, ctx->d_crop_any.as<int>());
        else crop_rows_kernel<false><<<(rows + 7) / 8, 256, 0, st>>>(d_crops, d_prefix, n, rows, ctx->d_crop_pix.as<unsigned char>(), nullptr, ctx->crops_lazy ? 1 : 0, ctx->d_crop_any.as<int>());
        RT_LAUNCH_CHECK(ctx);
    }
    ctx->crop_launch.d_prefix = d_prefix; ctx->crop_launch.n = n; ctx->crop_launch.rows = rows; ctx->crop_launch.d_totals = nullptr;
    // projection degeneracy is only known on the device: statuses (strided gather of one int per crop) come back async
    RT_CUDA_OK(ctx, ctx->h_crops.ensure(sizeof(int) * (size_t)n));
    RT_CUDA_OK(ctx, cudaMemcpy2DAsync(ctx->h_crops.p, sizeof(int), &d_crops[0].status, sizeof(CropDev), sizeof(int), n, cudaMemcpyDeviceToHost, st));
    return RETTO_B200_OK;
}
retto_b200_status rt_crop_launch(retto_b200_ctx* ctx, const retto_b200_crop_job* h_jobs, int n, retto_b200_crop_info* h_infos) {
    return crop_launch_impl(ctx, n, h_infos, [&](int i, const uint8_t** page, int* ph, int* pw) -> const float* {
        *page = h_jobs[i].d_page; *ph = h_jobs[i].page_h; *pw = h_jobs[i].page_w;
        return h_jobs[i].box.xy;
    });
}
// session variant: boxes in page order (box_off[p] .. box_off[p+1]), one page descriptor per page — no job array
retto_b200_status rt_crop_launch_pages(retto_b200_ctx* ctx, const retto_b200_box* h_boxes, const int32_t* box_off, int n_pages,
                                       const uint8_t* const* page_ptr, const int* page_h, const int* page_w, retto_b200_crop_info* h_infos) {
    int p = 0;
    return crop_launch_impl(ctx, box_off[n_pages], h_infos, [&](int i, const uint8_t** page, int* ph, int* pw) -> const float* {
        while (i >= box_off[p + 1]) ++p;   // boxes are visited in order
        *page = page_ptr[p]; *ph = page_h[p]; *pw = page_w[p];
        return h_boxes[i].xy;
    });
}
// statuses; do_sync == false when the caller has synchronised the stream since rt_crop_launch (session.cu defers this
// to the end of the page batch so that the host keeps running ahead of the GPU)
retto_b200_status rt_crop_finish(retto_b200_ctx* ctx, retto_b200_crop_info* h_infos, bool do_sync) {
    const int n = (int)ctx->crops.size();
    if (n == 0) return RETTO_B200_OK;
    if (do_sync) RT_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
    const int* hs = ctx->h_crops.as<int>();
    retto_b200_status ret = RETTO_B200_OK;
    if (ctx->crop_dev_check) {
        ctx->crop_dev_check = false;
        for (int i = 0; i < n; ++i)
            if (hs[n + 2 * i] != ctx->crops[i].w || hs[n + 2 * i + 1] != ctx->crops[i].h) {
                ctx->set_error("crop_boxes: host and device disagree on the size of crop " + std::to_string(i));
                return RETTO_B200_ERR_CUDA;
            }
    }
    for (int i = 0; i < n; ++i) {
        ctx->crops[i].status = hs[i];
        h_infos[i].status = hs[i];
        if (hs[i] != RETTO_B200_OK) { ctx->set_error("crop_boxes: degenerate quad " + std::to_string(i) + " (reference: from_control_points().unwrap() panics)"); ret = RETTO_B200_ERR_DEGENERATE_QUAD; }
    }
    return ret;
}

// session path, step 1 (before the host has the boxes): descriptor table, projection setup and the row kernel, all sized
// on the device.  `d_boxes` / `d_box_off` are the packed boxes of det_postprocess (box_off has n_pages + 1 entries).
retto_b200_status rt_crop_enqueue_device(retto_b200_ctx* ctx, const retto_b200_box* d_boxes, const int* d_box_off, int n_pages,
                                         const uint8_t* const* page_ptr, const int* page_h, const int* page_w, int crops_hint, int max_boxes) {
    cudaStream_t st = ctx->stream;
    ctx->crop_dev_cap = 0;
    if (n_pages <= 0) return RETTO_B200_OK;
    // capacities: grow-only buffers; the first batches of a context may fall back to the host-built table
    const int cap_crops = std::max(crops_hint, 1024);
    const size_t desc_bytes = sizeof(CropDev) * (size_t)cap_crops;
    const size_t tab_bytes = desc_bytes + sizeof(int) * ((size_t)cap_crops + 1) + 64;
    RT_CUDA_OK(ctx, ctx->d_crop_descs.ensure(tab_bytes, st));
    RT_CUDA_OK(ctx, ctx->d_crop_flip.ensure(sizeof(int) * (size_t)cap_crops, st));
    RT_CUDA_OK(ctx, ctx->d_crop_pix.ensure(1 << 20, st));
    std::vector<CropPageDev> pt(n_pages);
    for (int i = 0; i < n_pages; ++i) pt[i] = CropPageDev{page_ptr[i], page_h[i], page_w[i], 0};
    RT_TRY(rt_upload(ctx, ctx->d_crop_pages, pt.data(), sizeof(CropPageDev) * (size_t)n_pages));
    CropDev* d_crops = ctx->d_crop_descs.as<CropDev>();
    int* d_prefix = reinterpret_cast<int*>(ctx->d_crop_descs.as<char>() + desc_bytes);
    CropTotals* d_tot = reinterpret_cast<CropTotals*>(ctx->d_crop_descs.as<char>() + ((desc_bytes + sizeof(int) * ((size_t)cap_crops + 1) + 15) & ~size_t(15)));
    RT_LAUNCH_BEGIN(ctx, "crop_dims_kernel");
    crop_dims_kernel<<<(cap_crops + 127) / 128, 128, 0, st>>>(d_boxes, d_box_off, n_pages, ctx->d_crop_pages.as<CropPageDev>(), d_crops, cap_crops, max_boxes,
                                                             ctx->d_crop_flip.as<int>());
    RT_LAUNCH_CHECK(ctx);
    RT_LAUNCH_BEGIN(ctx, "crop_scan_kernel");
    crop_scan_kernel<<<1, 1024, 0, st>>>(d_box_off, n_pages, d_crops, d_prefix, cap_crops, (unsigned long long)ctx->d_crop_pix.cap, d_tot, max_boxes);
    RT_LAUNCH_CHECK(ctx);
    RT_LAUNCH_BEGIN(ctx, "crop_setup_kernel");
    RT_CUDA_OK(ctx, ctx->d_crop_any.ensure(16, st));
    RT_CUDA_OK(ctx, cudaMemsetAsync(ctx->d_crop_any.p, 0, 4, st));
    crop_setup_kernel<<<(cap_crops + 63) / 64, 64, 0, st>>>(d_crops, cap_crops, &d_tot->n, ctx->d_crop_any.as<int>());
    RT_LAUNCH_CHECK(ctx);
    // one warp per row like the host-sized launch, the grid sized from the largest batch seen so far (+25 %); a batch with more
    // rows than that falls back to the host-built table (a persistent grid-stride variant was measured 0.03-0.05 ms slower:
    // more registers, no dynamic balancing of the rows)
    const int row_hint = std::max(ctx->crop_rows_seen_max + ctx->crop_rows_seen_max / 4, 8192);
    ctx->crop_dev_row_cap = (row_hint + 7) / 8 * 8;
    RT_LAUNCH_BEGIN(ctx, "crop_rows_kernel");
    if (!getenv("RETTO_B200_CROP_SCALAR")) crop_rows_kernel<true><<<(row_hint + 7) / 8, 256, 0, st>>>(d_crops, d_prefix, 0, 0, ctx->d_crop_pix.as<unsigned char>(), d_tot, ctx->crops_lazy ? 1 : 0, ctx->d_crop_any.as<int>());
    else crop_rows_kernel<false><<<(row_hint + 7) / 8, 256, 0, st>>>(d_crops, d_prefix, 0, 0, ctx->d_crop_pix.as<unsigned char>(), d_tot, ctx->crops_lazy ? 1 : 0, ctx->d_crop_any.as<int>());
    RT_LAUNCH_CHECK(ctx);
    ctx->crop_launch.d_prefix = d_prefix; ctx->crop_launch.n = 0; ctx->crop_launch.rows = ctx->crop_dev_row_cap; ctx->crop_launch.d_totals = d_tot;
    ctx->crop_dev_cap = cap_crops;
    ctx->crop_dev_cap_bytes = ctx->d_crop_pix.cap;
    ctx->crop_dev_desc_bytes = desc_bytes;
    return RETTO_B200_OK;
}
// step 2 (the host has the boxes): the host's copy of the sizes; *fits == false when the device table did not hold the
// batch (the caller then runs rt_crop_launch_pages).  Enqueues the status read-back.
retto_b200_status rt_crop_adopt_device(retto_b200_ctx* ctx, const retto_b200_box* h_boxes, int n, retto_b200_crop_info* h_infos, bool* fits) {
    *fits = false;
    ctx->crops.clear();
    if (n == 0) { *fits = true; return RETTO_B200_OK; }
    if (ctx->crop_dev_cap <= 0 || n > ctx->crop_dev_cap) return RETTO_B200_OK;
    ctx->crops.resize(n);
    unsigned long long off = 0;
    for (int i = 0; i < n; ++i) {
        int cw, ch, rot, status = RETTO_B200_OK;
        rt_crop_dims(h_boxes[i].xy, &cw, &ch, &rot);
        const long long px = (long long)cw * ch;
        const unsigned long long o = off;
        if (px <= 0 || px > 0x3fffffffLL) { status = RETTO_B200_ERR_DEGENERATE_QUAD; cw = ch = 0; }
        else off += ((unsigned long long)px * 4 + 15) & ~15ULL;
        ctx->crops[i] = retto_b200_ctx::CropHost{cw, ch, rot, status, o};
        h_infos[i].w = cw; h_infos[i].h = ch; h_infos[i].rotated270 = rot; h_infos[i].status = status; h_infos[i].offset = o;
    }
    {
        long long rows = 0;
        for (int i = 0; i < n; ++i) rows += ctx->crops[i].rot ? ctx->crops[i].w : ctx->crops[i].h;
        ctx->crop_rows_seen_max = (int)std::min<long long>(std::max<long long>(ctx->crop_rows_seen_max, rows), 1 << 28);
        if (rows > ctx->crop_dev_row_cap) { ctx->crops.clear(); return RETTO_B200_OK; }   // the row grid was too small: host-built table
    }
    if (off > ctx->crop_dev_cap_bytes) {   // the device saw the same total and did nothing: grow with headroom for the next batch
        RT_CUDA_OK(ctx, ctx->d_crop_pix.ensure((size_t)(off + off / 4), ctx->stream));
        ctx->crops.clear();
        return RETTO_B200_OK;
    }
    CropDev* d_crops = ctx->d_crop_descs.as<CropDev>();
    RT_CUDA_OK(ctx, ctx->h_crops.ensure(sizeof(int) * 3 * (size_t)n));
    // status + (w, h) of every crop: the host's sizes must be the device's (same IEEE operations; checked in rt_crop_finish)
    RT_CUDA_OK(ctx, cudaMemcpy2DAsync(ctx->h_crops.p, sizeof(int), &d_crops[0].status, sizeof(CropDev), sizeof(int), n, cudaMemcpyDeviceToHost, ctx->stream));
    RT_CUDA_OK(ctx, cudaMemcpy2DAsync(ctx->h_crops.as<int>() + n, 2 * sizeof(int), &d_crops[0].w, sizeof(CropDev), 2 * sizeof(int), n, cudaMemcpyDeviceToHost, ctx->stream));
    ctx->crop_dev_check = true;
    *fits = true;
    return RETTO_B200_OK;
}

extern "C" retto_b200_status retto_b200_crop_boxes(retto_b200_ctx* ctx, const retto_b200_crop_job* h_jobs, int32_t n,
                                                   retto_b200_crop_info* h_infos) {
    RtDeviceGuard _dg(ctx);
    if (!ctx || n < 0 || (n > 0 && (!h_jobs || !h_infos))) return RETTO_B200_ERR_INVALID_ARG;
    ctx->crops_lazy = getenv("RETTO_B200_CROP_LAZY") != nullptr;   // stage API: crops are materialised (tests may ask for the session's lazy mode)
    RT_TRY(rt_crop_launch(ctx, h_jobs, n, h_infos));
    return rt_crop_finish(ctx, h_infos, true);
}

__global__ void crop_flip_copy_kernel(const unsigned char* __restrict__ src, unsigned char* __restrict__ dst, int n_px, int flip) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_px) return;
    const int s = flip ? n_px - 1 - i : i;
    dst[3 * i] = src[4 * s]; dst[3 * i + 1] = src[4 * s + 1]; dst[3 * i + 2] = src[4 * s + 2];   // RGBX -> RGB
}

extern "C" retto_b200_status retto_b200_crop_fetch(retto_b200_ctx* ctx, int32_t i, uint8_t* h_out) {
    RtDeviceGuard _dg(ctx);
    if (!ctx || i < 0 || i >= (int)ctx->crops.size() || !h_out) return RETTO_B200_ERR_INVALID_ARG;
    const retto_b200_ctx::CropHost& c = ctx->crops[i];
    const int n_px = c.w * c.h;
    if (n_px == 0) return RETTO_B200_OK;
    if (ctx->crops_lazy && ctx->crop_launch.d_prefix) {   // direct crops were left to the batch build: materialise them now
        const retto_b200_ctx::CropLaunch& cl = ctx->crop_launch;
        RT_LAUNCH_BEGIN(ctx, "crop_rows_kernel");
        crop_rows_kernel<true><<<(std::max(cl.rows, 1) + 7) / 8, 256, 0, ctx->stream>>>(ctx->d_crop_descs.as<CropDev>(), cl.d_prefix, cl.n, cl.rows, ctx->d_crop_pix.as<unsigned char>(),
                                                                                      reinterpret_cast<const CropTotals*>(cl.d_totals), 0, nullptr);
        RT_LAUNCH_CHECK(ctx);
        ctx->crops_lazy = false;
    }
    int flip = 0;
    RT_CUDA_OK(ctx, cudaMemcpyAsync(&flip, ctx->d_crop_flip.as<int>() + i, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    RT_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
    RT_CUDA_OK(ctx, ctx->d_stage3.ensure((size_t)n_px * 3, ctx->stream));
    // rotate_180_in_place (image_helper.rs:268-286) is applied lazily: materialise it here
    RT_LAUNCH_BEGIN(ctx, "crop_flip_copy_kernel");
    crop_flip_copy_kernel<<<(n_px + 255) / 256, 256, 0, ctx->stream>>>(ctx->d_crop_pix.as<unsigned char>() + c.offset, ctx->d_stage3.as<unsigned char>(), n_px, flip);
    RT_LAUNCH_CHECK(ctx);
    RT_CUDA_OK(ctx, cudaMemcpyAsync(h_out, ctx->d_stage3.p, (size_t)n_px * 3, cudaMemcpyDeviceToHost, ctx->stream));
    RT_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
    return RETTO_B200_OK;
}
