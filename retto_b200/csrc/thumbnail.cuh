// thumbnail.cuh — device restatement of image 0.25.6 `imageops::thumbnail` (box filter with the
// fractional up-scaling branches), the only resampler retto uses (image_helper.rs:124,139,168,184).
// One call computes the 3 channels of one output pixel from an HWC u8 source.  All float math uses
// explicitly rounded intrinsics so no FMA contraction can change a truncation.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

struct ThumbAxis {  // per output coordinate: window [lo, hi) or fractional pair (hi-1, hi) when lo == hi
    unsigned lo, hi;
    float fract;    // (fract(lof) + fract(hif)) / 2   — only meaningful when lo == hi
};

__device__ __forceinline__ float rt_fract(float v) { return __fsub_rn(v, truncf(v)); }

__device__ __forceinline__ ThumbAxis thumb_axis(int out_i, float ratio, unsigned size) {
    const float lof = __fmul_rn((float)out_i, ratio);
    const float hif = __fadd_rn(lof, ratio);
    unsigned lo = (unsigned)ceilf(lof);
    lo = lo > size - 1 ? size - 1 : lo;
    unsigned hi = (unsigned)ceilf(hif);
    hi = hi < lo ? lo : (hi > size ? size : hi);
    ThumbAxis a;
    a.lo = lo;
    a.hi = hi;
    a.fract = __fmul_rn(__fadd_rn(rt_fract(lof), rt_fract(hif)), 0.5f);   // "/ 2.0": halving is exact, same bits as the division
    return a;
}

__device__ __forceinline__ unsigned char f32_to_u8_numcast(float v) {
    if (!(v > -1.0f && v < 256.0f)) return v >= 256.0f ? 255 : 0;
    return (unsigned char)v;  // truncation
}

// Reader: functor (x, y) -> uchar3-like struct with .x .y .z ; lets callers read through a flip.
template <class Reader>
__device__ __forceinline__ void thumbnail_pixel(const Reader& rd, unsigned w, unsigned h, const ThumbAxis ax, const ThumbAxis ay,
                                                unsigned char out[3]) {
    const unsigned left = ax.lo, right = ax.hi, bottom = ay.lo, top = ay.hi;
    if (bottom != top && left != right) {
        unsigned s0 = 0, s1 = 0, s2 = 0;
        for (unsigned y = bottom; y < top; ++y)
            for (unsigned x = left; x < right; ++x) {
                const uchar3 p = rd(x, y);
                s0 += p.x; s1 += p.y; s2 += p.z;
            }
        const unsigned n = (right - left) * (top - bottom);
        const unsigned r = n >> 1;
        unsigned v0, v1, v2;
        if (n == 1) { v0 = s0; v1 = s1; v2 = s2; }
        else if (n < 4096) {
            // (s + r) / n without three integer divisions: floor((x + 0.5) * (1/n)) is exact here — (x + 0.5) / n is at least
            // 0.5/n away from every integer, and the two roundings move the product by less than 255.5 * 2^-22 < 0.5/4096
            const float inv = __frcp_rn((float)n);
            v0 = __float2uint_rd(__fmul_rn(__fadd_rn((float)(s0 + r), 0.5f), inv));
            v1 = __float2uint_rd(__fmul_rn(__fadd_rn((float)(s1 + r), 0.5f), inv));
            v2 = __float2uint_rd(__fmul_rn(__fadd_rn((float)(s2 + r), 0.5f), inv));
        } else { v0 = (s0 + r) / n; v1 = (s1 + r) / n; v2 = (s2 + r) / n; }
        out[0] = (unsigned char)(v0 > 255 ? 255 : v0);
        out[1] = (unsigned char)(v1 > 255 ? 255 : v1);
        out[2] = (unsigned char)(v2 > 255 ? 255 : v2);
    } else if (bottom != top) {  // horizontal fraction between columns (right-1, right)
        const unsigned l = right - 1;
        const unsigned l1 = (l + 1 > w - 1) ? w - 1 : l + 1;
        unsigned a0 = 0, a1 = 0, a2 = 0, b0 = 0, b1 = 0, b2 = 0;
        for (unsigned y = bottom; y < top; ++y) {
            const uchar3 p = rd(l, y), q = rd(l1, y);
            a0 += p.x; a1 += p.y; a2 += p.z;
            b0 += q.x; b1 += q.y; b2 += q.z;
        }
        const float cnt = (float)(top - bottom);
        const float fr = __fdiv_rn(ax.fract, cnt);
        const float fl = __fdiv_rn(__fsub_rn(1.0f, ax.fract), cnt);
        out[0] = f32_to_u8_numcast(__fadd_rn(__fmul_rn(fl, (float)a0), __fmul_rn(fr, (float)b0)));
        out[1] = f32_to_u8_numcast(__fadd_rn(__fmul_rn(fl, (float)a1), __fmul_rn(fr, (float)b1)));
        out[2] = f32_to_u8_numcast(__fadd_rn(__fmul_rn(fl, (float)a2), __fmul_rn(fr, (float)b2)));
    } else if (left != right) {  // vertical fraction between rows (top-1, top)
        const unsigned b = top - 1;
        const unsigned b1r = (b + 1 > h - 1) ? h - 1 : b + 1;
        unsigned a0 = 0, a1 = 0, a2 = 0, c0 = 0, c1 = 0, c2 = 0;
        for (unsigned x = left; x < right; ++x) {
            const uchar3 p = rd(x, b), q = rd(x, b1r);
            a0 += p.x; a1 += p.y; a2 += p.z;
            c0 += q.x; c1 += q.y; c2 += q.z;
        }
        const float cnt = (float)(right - left);
        const float ft = __fdiv_rn(ay.fract, cnt);
        const float fb = __fdiv_rn(__fsub_rn(1.0f, ay.fract), cnt);
        out[0] = f32_to_u8_numcast(__fadd_rn(__fmul_rn(fb, (float)a0), __fmul_rn(ft, (float)c0)));
        out[1] = f32_to_u8_numcast(__fadd_rn(__fmul_rn(fb, (float)a1), __fmul_rn(ft, (float)c1)));
        out[2] = f32_to_u8_numcast(__fadd_rn(__fmul_rn(fb, (float)a2), __fmul_rn(ft, (float)c2)));
    } else {
        const unsigned l = right - 1, b = top - 1;
        const unsigned l1 = (l + 1 > w - 1) ? w - 1 : l + 1;
        const unsigned b1r = (b + 1 > h - 1) ? h - 1 : b + 1;
        const float fv = ay.fract, fh = ax.fract;
        const float f_tr = __fmul_rn(fv, fh);
        const float f_tl = __fmul_rn(fv, __fsub_rn(1.0f, fh));
        const float f_br = __fmul_rn(__fsub_rn(1.0f, fv), fh);
        const float f_bl = __fmul_rn(__fsub_rn(1.0f, fv), __fsub_rn(1.0f, fh));
        const uchar3 kbl = rd(l, b), ktl = rd(l, b1r), kbr = rd(l1, b), ktr = rd(l1, b1r);
#define RT_MIX(bl, tl, br, tr)                                                                                     \
    f32_to_u8_numcast(__fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(f_br, (float)(br)), __fmul_rn(f_tr, (float)(tr))),   \
                                          __fmul_rn(f_bl, (float)(bl))),                                           \
                                __fmul_rn(f_tl, (float)(tl))))
        out[0] = RT_MIX(kbl.x, ktl.x, kbr.x, ktr.x);
        out[1] = RT_MIX(kbl.y, ktl.y, kbr.y, ktr.y);
        out[2] = RT_MIX(kbl.z, ktl.z, kbr.z, ktr.z);
#undef RT_MIX
    }
}

struct PlainReader {
    const unsigned char* __restrict__ src;
    unsigned w;
    __device__ __forceinline__ uchar3 operator()(unsigned x, unsigned y) const {
        const unsigned char* p = src + ((size_t)y * w + x) * 3;
        return make_uchar3(__ldg(p), __ldg(p + 1), __ldg(p + 2));
    }
};
