// rt_fmath.h — deterministic f64 elementary functions (atan2 / sincos / acos) built only from
// IEEE-754 correctly-rounded primitives (+ - * / sqrt fma), so the SAME bits come out of an
// sm_100a kernel and of x86-64 host code.
//
// Why this exists: the DB post-process geometry of the reference runs
//   imageproc::geometry::min_area_rect   (atan2 -> fmod -> sin/cos -> floor/ceil)   det_processor.rs:180
//   Clipper ClipperOffset::DoOffset      (acos, sin, cos, atan2 -> Round)            det_processor.rs:241-242
// in f64 on the host libm.  floor/ceil/Round make the integer box corners discontinuous in the
// last ulp of those calls, and CUDA's libdevice (1-2 ulp) is not bit-identical to glibc.  The
// functions below evaluate in double-double (~100 bits) and round once, i.e. they return the
// correctly rounded result except for astronomically rare ties; glibc's own results are within
// 0.55 ulp, so the two agree wherever glibc is correctly rounded.  tests/test_fmath.py measures
// the agreement with glibc exhaustively over the domain the path can reach (integer edge vectors
// of a <= 4096 px page).
//
// Compile rules: device TUs that include this need `-fmad=false`; host TUs `-ffp-contract=off`.
// fma() is used explicitly for the exact product split and must stay a real fused operation.
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define RT_HD __host__ __device__ __forceinline__
#define RT_HD_BIG static __host__ __device__ __noinline__   // big bodies: one copy per TU keeps kernels small (I-cache, registers)
#define RT_NOUNROLL _Pragma("unroll 1")
#else
#define RT_HD inline
#define RT_HD_BIG static inline
#define RT_NOUNROLL
#endif

namespace rtm {

struct dd {
    double hi, lo;
};

RT_HD dd two_sum(double a, double b) {
    double s = a + b;
    double bb = s - a;
    double e = (a - (s - bb)) + (b - bb);
    return dd{s, e};
}
RT_HD dd quick_two_sum(double a, double b) {
    double s = a + b;
    double e = b - (s - a);
    return dd{s, e};
}
RT_HD dd two_prod(double a, double b) {
    double p = a * b;
    double e = fma(a, b, -p);
    return dd{p, e};
}
RT_HD dd dd_from(double a) { return dd{a, 0.0}; }
RT_HD dd dd_neg(dd a) { return dd{-a.hi, -a.lo}; }
RT_HD dd dd_add(dd a, dd b) {
    dd s = two_sum(a.hi, b.hi);
    dd t = two_sum(a.lo, b.lo);
    s.lo = s.lo + t.hi;
    s = quick_two_sum(s.hi, s.lo);
    s.lo = s.lo + t.lo;
    return quick_two_sum(s.hi, s.lo);
}
RT_HD dd dd_sub(dd a, dd b) { return dd_add(a, dd_neg(b)); }
RT_HD dd dd_add_d(dd a, double b) {
    dd s = two_sum(a.hi, b);
    s.lo = s.lo + a.lo;
    return quick_two_sum(s.hi, s.lo);
}
RT_HD dd dd_mul(dd a, dd b) {
    dd p = two_prod(a.hi, b.hi);
    p.lo = p.lo + (a.hi * b.lo + a.lo * b.hi);
    return quick_two_sum(p.hi, p.lo);
}
RT_HD dd dd_mul_d(dd a, double b) {
    dd p = two_prod(a.hi, b);
    p.lo = p.lo + a.lo * b;
    return quick_two_sum(p.hi, p.lo);
}
RT_HD dd dd_scale2(dd a, double pow2) { return dd{a.hi * pow2, a.lo * pow2}; }  // exact
RT_HD dd dd_div(dd a, dd b) {
    double q1 = a.hi / b.hi;
    dd r = dd_sub(a, dd_mul_d(b, q1));
    double q2 = r.hi / b.hi;
    r = dd_sub(r, dd_mul_d(b, q2));
    double q3 = r.hi / b.hi;
    dd q = quick_two_sum(q1, q2);
    return dd_add_d(q, q3);
}
RT_HD dd dd_div_d(dd a, double b) { return dd_div(a, dd_from(b)); }
RT_HD dd dd_sqrt(dd a) {
    if (a.hi <= 0.0) return dd{0.0, 0.0};
    double s0 = sqrt(a.hi);
    dd s0s = two_prod(s0, s0);
    dd r = dd_sub(a, s0s);
    double corr = r.hi / (2.0 * s0);
    return quick_two_sum(s0, corr);
}
RT_HD double dd_round(dd a) { return a.hi + a.lo; }

// pi/2 split into three doubles (hi + mid + lo carries > 150 bits)
#define RT_PIO2_HI 1.5707963267948966
#define RT_PIO2_MD 6.123233995736766e-17
#define RT_PIO2_LO -1.4973849048591698e-33
#define RT_PI_D 3.141592653589793
#define RT_PIO2_D 1.5707963267948966

// sin and cos of a double-double argument, |a| small multiples of pi (no huge-argument reduction:
// the path only produces |a| <= 2*pi).
RT_HD_BIG void sincos_dd(dd a, dd* s_out, dd* c_out) {
    // quadrant reduction: a = k*(pi/2) + r, |r| <= pi/4
    double kf = floor(a.hi * 0.6366197723675814 + 0.5);
    dd r = a;
    if (kf != 0.0) {
        r = dd_sub(r, two_prod(kf, RT_PIO2_HI));
        r = dd_sub(r, two_prod(kf, RT_PIO2_MD));
        r = dd_sub(r, two_prod(kf, RT_PIO2_LO));
    }
    int k = ((int)kf) & 3;
    // halve 4 times: |r'| <= pi/64
    dd x = dd_scale2(r, 0.0625);
    dd x2 = dd_mul(x, x);
    // Taylor: sin x = x - x^3/3! + ... (to x^17);  cos x - 1 = -x^2/2! + x^4/4! - ... (to x^16)
    // term ratios 1/((2i)(2i+1)) and 1/((2i-1)(2i)) as double-double constants (a dd multiply instead of a dd divide)
    const double SR_HI[8] = {0.16666666666666666, 0.05, 0.023809523809523808, 0.013888888888888888, 0.00909090909090909,
                             0.00641025641025641, 0.004761904761904762, 0.003676470588235294};
    const double SR_LO[8] = {9.25185853854297e-18, -2.7755575615628915e-18, 1.32169407693471e-18, 7.709882115452476e-19,
                             4.415659757031872e-19, 2.2240044563805217e-19, -4.295505750037808e-19, 5.102127870520021e-20};
    const double CR_HI[7] = {0.08333333333333333, 0.03333333333333333, 0.017857142857142856, 0.011111111111111112,
                             0.007575757575757576, 0.005494505494505495, 0.004166666666666667};
    const double CR_LO[7] = {4.625929269271485e-18, 4.625929269271486e-19, 9.912705577010326e-19, -4.2404351634988616e-19,
                             -2.1026951223961299e-19, -4.2891514515910067e-19, 5.782411586589357e-20};
    dd term = x;   // x^(2i+1)/(2i+1)!
    dd s = x;
    RT_NOUNROLL
    for (int i = 1; i <= 8; ++i) {
        term = dd_mul(dd_mul(term, x2), dd{SR_HI[i - 1], SR_LO[i - 1]});
        s = (i & 1) ? dd_sub(s, term) : dd_add(s, term);
    }
    dd cterm = dd_scale2(x2, 0.5);  // x^2/2!
    dd cm1 = dd_neg(cterm);
    RT_NOUNROLL
    for (int i = 2; i <= 8; ++i) {
        cterm = dd_mul(dd_mul(cterm, x2), dd{CR_HI[i - 2], CR_LO[i - 2]});
        cm1 = (i & 1) ? dd_sub(cm1, cterm) : dd_add(cm1, cterm);
    }
    // double-angle x4:  sin 2x = 2 s (1 + cm1);  cos 2x - 1 = -2 s^2
    RT_NOUNROLL
    for (int i = 0; i < 4; ++i) {
        dd c = dd_add_d(cm1, 1.0);
        dd s2 = dd_scale2(dd_mul(s, c), 2.0);
        dd cm = dd_scale2(dd_mul(s, s), -2.0);
        s = s2;
        cm1 = cm;
    }
    dd c = dd_add_d(cm1, 1.0);
    dd so, co;
    switch (k) {
        case 0: so = s; co = c; break;
        case 1: so = c; co = dd_neg(s); break;
        case 2: so = dd_neg(s); co = dd_neg(c); break;
        default: so = dd_neg(c); co = s; break;
    }
    *s_out = so;
    *c_out = co;
}

RT_HD void rt_sincos(double a, double* s, double* c) {
    if (a == 0.0) {
        *s = a;
        *c = 1.0;
        return;
    }
    dd sd, cd;
    sincos_dd(dd_from(a), &sd, &cd);
    *s = dd_round(sd);
    *c = dd_round(cd);
}
RT_HD double rt_sin(double a) {
    double s, c;
    rt_sincos(a, &s, &c);
    return s;
}
RT_HD double rt_cos(double a) {
    double s, c;
    rt_sincos(a, &s, &c);
    return c;
}

// atan2 of double-double (y, x); returns dd.  (x, y) != (0, 0).
RT_HD_BIG dd atan2_dd(dd y, dd x) {
    double ax = fabs(x.hi), ay = fabs(y.hi);
    // crude double guess (max error ~1e-5 rad)
    double mn = ax < ay ? ax : ay, mx = ax < ay ? ay : ax;
    double t = mn / mx;
    double t2 = t * t;
    double p = t * (0.99997726 + t2 * (-0.33262347 + t2 * (0.19354346 + t2 * (-0.11643287 + t2 * (0.05265332 + t2 * -0.01172120)))));
    double th = (ay > ax) ? (RT_PIO2_D - p) : p;
    if (x.hi < 0.0) th = RT_PI_D - th;
    if (y.hi < 0.0) th = -th;
    // one Newton-like correction:  theta* = th + atan( (y c - x s) / (x c + y s) )
    dd s, c;
    sincos_dd(dd_from(th), &s, &c);
    dd num = dd_sub(dd_mul(y, c), dd_mul(x, s));
    dd den = dd_add(dd_mul(x, c), dd_mul(y, s));
    dd d = dd_div(num, den);
    dd d2 = dd_mul(d, d);
    // atan(d) = d (1 - d^2/3 + d^4/5 - d^6/7 + d^8/9 - d^10/11),  |d| <= ~1e-3 (d^13/13 negligible)
    const dd R3{0.3333333333333333, 1.850371707708594e-17}, R5{0.2, -1.1102230246251566e-17}, R7{0.14285714285714285, 7.93016446160826e-18},
        R9{0.1111111111111111, 6.1679056923619804e-18}, R11{0.09090909090909091, -2.523234146875356e-18};
    dd poly = dd_sub(R9, dd_mul(d2, R11));                               // 1/9 - d2/11
    poly = dd_sub(R7, dd_mul(d2, poly));                                 // 1/7 - d2 (...)
    poly = dd_sub(R5, dd_mul(d2, poly));                                 // 1/5 - d2 (...)
    poly = dd_sub(R3, dd_mul(d2, poly));                                 // 1/3 - d2 (...)
    poly = dd_sub(dd_from(1.0), dd_mul(d2, poly));                       // 1 - d2 (...)
    dd corr = dd_mul(d, poly);
    return dd_add(dd_from(th), corr);
}

RT_HD double rt_atan2(double y, double x) {
    if (y == 0.0) {
        if (x > 0.0 || x == 0.0) return y;  // +-0 (x==0: atan2(0,0)=0 by convention; unreachable on the path)
        return (signbit(y) ? -RT_PI_D : RT_PI_D);
    }
    if (x == 0.0) return y > 0.0 ? RT_PIO2_D : -RT_PIO2_D;
    return dd_round(atan2_dd(dd_from(y), dd_from(x)));
}

// acos(v), |v| <= 1
RT_HD_BIG double rt_acos(double v) {
    if (v >= 1.0) return 0.0;
    if (v <= -1.0) return RT_PI_D;
    dd om = two_sum(1.0, -v);
    dd op = two_sum(1.0, v);
    dd w = dd_mul(om, op);
    dd sq = dd_sqrt(w);
    return dd_round(atan2_dd(sq, dd_from(v)));
}

}  // namespace rtm
