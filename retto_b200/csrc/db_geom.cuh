// db_geom.cuh — per-box geometry of DBNet postprocessing as device functions (one warp per box):
//   convex hull (imageproc::geometry::convex_hull order), rotating calipers / min_area_rect,
//   draw_polygon_mut scan-fill + Bresenham row coverage, box_score_fast, geo area/length,
//   Clipper round-join offset, scale_and_clip.  Reference call sites: det_processor.rs:176-252,
//   295-320; points.rs:125-194.  Recalled third-party semantics are restated in
//   oracle/retto_oracle.cpp (same section names) — the two are written independently but must agree
//   bit for bit; trig goes through rt_fmath.h on both sides.
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>

#include "rt_fmath.h"

#define RT_FULL 0xffffffffu

// ---- convex hull from per-row extremes ---------------------------------------------------------
// rows: n entries sorted by strictly increasing y, each (y, xmin, xmax).  Output order == imageproc's
// Graham scan: starts at (xmin of the first row, y0) and walks in the direction where
// orient(p,q,r) = (q.y-p.y)(r.x-q.x) - (q.x-p.x)(r.y-q.y) < 0 ("CounterClockwise"), i.e. along the
// top to the right, down the right side, back up the left side; collinear points dropped.
// Serial (call from one lane).  `out` needs room for 2*n points.  Returns the hull size.
template <class RowFn>
static __device__ int hull_from_rows(int n, RowFn row, int2* out) {
    auto turn_ok = [](int2 p, int2 q, int2 r) -> bool {  // strictly "CounterClockwise" in imageproc's sense
        const long long v = (long long)(q.y - p.y) * (long long)(r.x - q.x) - (long long)(q.x - p.x) * (long long)(r.y - q.y);
        return v < 0;
    };
    // Two independent monotone chains (Andrew): the right side of the hull only involves the row
    // maxima, the left side only the row minima; each chain has its own stack floor so that the
    // second scan can never pop vertices of the first.  Chain end points (first/last row extremes)
    // are always hull vertices, and the junctions are strictly convex (horizontal top/bottom edge).
    int m = 0;
    int y0, a0, b0;
    row(0, y0, a0, b0);
    out[m++] = make_int2(a0, y0);            // start point S
    int base = (a0 == b0) ? 0 : 1;           // S doubles as the right chain's first point when the top row is 1 px
    // right chain, top -> bottom, over (xmax, y)
    for (int i = 0; i < n; ++i) {
        int y, a, b;
        row(i, y, a, b);
        const int2 p = make_int2(b, y);
        if (i == 0 && base == 0) continue;   // already there as S
        while (m - base >= 2 && !turn_ok(out[m - 2], out[m - 1], p)) --m;
        out[m++] = p;
    }
    // left chain, bottom -> top, over (xmin, y); ends at S, which is dropped
    {
        int yl, al, bl;
        row(n - 1, yl, al, bl);
        base = (al == bl) ? m - 1 : m;       // bottom row of 1 px: the right chain's last point starts the left chain
    }
    for (int i = n - 1; i >= 0; --i) {
        int y, a, b;
        row(i, y, a, b);
        const int2 p = make_int2(a, y);
        if (i == n - 1 && base == m - 1) continue;
        while (m - base >= 2 && !turn_ok(out[m - 2], out[m - 1], p)) --m;
        if (i == 0) break;                   // p == S closes the polygon: pop against it but never store it (keeps m <= 2n)
        out[m++] = p;
    }
    // all points collinear: the chains are [S, E] and [E, S] -> [S, E]
    return m;
}

// ---- min_area_rect ---------------------------------------------------------------------------------
// Rotating calipers over hull.windows(2) (closing edge not visited), whole warp cooperates: lane l
// evaluates edges l, l+32, ...; the winner is the first edge with the strictly smallest area.
// q[8] = tl.x tl.y tr.x tr.y br.x br.y bl.x bl.y as doubles holding integers (floor/ceil applied).
static __device__ void warp_min_area_rect(const int2* hull, int n, double q[8]) {
    const int lane = threadIdx.x & 31;
    if (n == 1) {
#pragma unroll
        for (int i = 0; i < 4; ++i) { q[2 * i] = (double)hull[0].x; q[2 * i + 1] = (double)hull[0].y; }
        return;
    }
    if (n == 2) {
        q[0] = hull[0].x; q[1] = hull[0].y; q[2] = hull[1].x; q[3] = hull[1].y;
        q[4] = hull[1].x; q[5] = hull[1].y; q[6] = hull[0].x; q[7] = hull[0].y;
        return;
    }
    const double PI = 3.14159265358979323846;
    double best_area = 1.7976931348623157e308;
    int best_i = 0x7fffffff;
    double b_s = 0, b_c = 1, b_minx = 0, b_maxx = 0, b_miny = 0, b_maxy = 0;
    for (int e = lane; e < n - 1; e += 32) {
        const int2 a = hull[e], b = hull[e + 1];
        // an earlier edge with the same vector has the same angle, hence the same rectangle and area — and the earlier edge wins
        // the "first strictly smallest" rule anyway: skip the double-double trig (the hull of a rasterised slanted side repeats a
        // handful of step vectors many times)
        {
            const int dxi = b.x - a.x, dyi = b.y - a.y;
            bool dup = false;
            for (int k = 0; k < e && !dup; ++k) dup = (hull[k + 1].x - hull[k].x == dxi) && (hull[k + 1].y - hull[k].y == dyi);
            if (dup) continue;
        }
        const double ex = (double)b.x - (double)a.x, ey = (double)b.y - (double)a.y;
        const double ang = fabs(fmod(__dadd_rn(rtm::rt_atan2(ey, ex), PI), PI / 2.0));
        double s, c;
        rtm::rt_sincos(ang, &s, &c);
        double min_x = 1.7976931348623157e308, max_x = -1.7976931348623157e308;
        double min_y = 1.7976931348623157e308, max_y = -1.7976931348623157e308;
        for (int k = 0; k < n; ++k) {
            const double px = (double)hull[k].x, py = (double)hull[k].y;
            const double rx = __dadd_rn(__dmul_rn(px, c), __dmul_rn(py, s));
            const double ry = __dsub_rn(__dmul_rn(py, c), __dmul_rn(px, s));
            min_x = fmin(min_x, rx); max_x = fmax(max_x, rx);
            min_y = fmin(min_y, ry); max_y = fmax(max_y, ry);
        }
        const double area = __dmul_rn(__dsub_rn(max_x, min_x), __dsub_rn(max_y, min_y));
        if (area < best_area) {  // within a lane edges come in increasing order: strict < keeps the first
            best_area = area; best_i = e; b_s = s; b_c = c;
            b_minx = min_x; b_maxx = max_x; b_miny = min_y; b_maxy = max_y;
        }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        const double oa = __shfl_xor_sync(RT_FULL, best_area, off);
        const int oi = __shfl_xor_sync(RT_FULL, best_i, off);
        const double os = __shfl_xor_sync(RT_FULL, b_s, off), oc = __shfl_xor_sync(RT_FULL, b_c, off);
        const double o1 = __shfl_xor_sync(RT_FULL, b_minx, off), o2 = __shfl_xor_sync(RT_FULL, b_maxx, off);
        const double o3 = __shfl_xor_sync(RT_FULL, b_miny, off), o4 = __shfl_xor_sync(RT_FULL, b_maxy, off);
        if (oa < best_area || (oa == best_area && oi < best_i)) {
            best_area = oa; best_i = oi; b_s = os; b_c = oc; b_minx = o1; b_maxx = o2; b_miny = o3; b_maxy = o4;
        }
    }
    // all lanes hold the winner; corners via invert_rotation: (x c - y s, y c + x s)
    double rx[4], ry[4];
    const double cx[4] = {b_maxx, b_minx, b_minx, b_maxx};
    const double cy[4] = {b_miny, b_miny, b_maxy, b_maxy};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        rx[i] = __dsub_rn(__dmul_rn(cx[i], b_c), __dmul_rn(cy[i], b_s));
        ry[i] = __dadd_rn(__dmul_rn(cy[i], b_c), __dmul_rn(cx[i], b_s));
    }
    // stable sort of 4 by x (insertion)
#pragma unroll
    for (int i = 1; i < 4; ++i) {
        const double vx = rx[i], vy = ry[i];
        int j = i;
        while (j > 0 && vx < rx[j - 1]) { rx[j] = rx[j - 1]; ry[j] = ry[j - 1]; --j; }
        rx[j] = vx; ry[j] = vy;
    }
    const int i1 = ry[1] > ry[0] ? 0 : 1;
    const int i2 = ry[3] > ry[2] ? 2 : 3;
    const int i3 = ry[3] > ry[2] ? 3 : 2;
    const int i4 = ry[1] > ry[0] ? 1 : 0;
    q[0] = floor(rx[i1]); q[1] = floor(ry[i1]);
    q[2] = ceil(rx[i2]);  q[3] = floor(ry[i2]);
    q[4] = ceil(rx[i3]);  q[5] = ceil(ry[i3]);
    q[6] = floor(rx[i4]); q[7] = ceil(ry[i4]);
}

// det_processor.rs:166-186: sside = min(|tl-tr|, |bl-br|) in f32
__device__ __forceinline__ float euclid_f32(float ax, float ay, float bx, float by) {
    const float dx = __fsub_rn(ax, bx), dy = __fsub_rn(ay, by);
    return __fsqrt_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)));
}
__device__ __forceinline__ float sside_of(const double q[8]) {
    const float s1 = euclid_f32((float)q[0], (float)q[1], (float)q[2], (float)q[3]);
    const float s2 = euclid_f32((float)q[6], (float)q[7], (float)q[4], (float)q[5]);
    return fminf(s1, s2);
}

// ---- draw_polygon_mut row coverage -----------------------------------------------------------------
// For canvas row y (relative coords, canvas bw x bh, quad px/py relative) produce the disjoint, sorted
// x-intervals the reference mask has set on that row: scan-fill pairs + Bresenham edge pixels.
struct RowCover {
    int n;
    int a[8], b[8];
};

__device__ __forceinline__ void cover_add(RowCover& rc, int a, int b, int bw) {
    if (a < 0) a = 0;
    if (b > bw - 1) b = bw - 1;
    if (a > b) return;
    rc.a[rc.n] = a; rc.b[rc.n] = b; rc.n++;
}

static __device__ void polygon_row_cover(const int px[4], const int py[4], int bw, int bh, int y, RowCover& rc) {
    rc.n = 0;
    // scan-fill intersections
    int inter[8];
    int ni = 0;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const int x0 = px[e], y0 = py[e], x1 = px[(e + 1) & 3], y1 = py[(e + 1) & 3];
        if ((y0 <= y && y1 >= y) || (y1 <= y && y0 >= y)) {
            if (y0 == y1) { inter[ni++] = x0; inter[ni++] = x1; }
            else if (y0 == y || y1 == y) {
                if (y1 > y) inter[ni++] = x0;
                if (y0 > y) inter[ni++] = x1;
            } else {
                const float fraction = __fdiv_rn((float)(y - y0), (float)(y1 - y0));
                const float v = __fadd_rn((float)x0, __fmul_rn(fraction, (float)(x1 - x0)));
                inter[ni++] = (int)roundf(v);
            }
        }
    }
    for (int i = 1; i < ni; ++i) {
        const int v = inter[i];
        int j = i;
        while (j > 0 && inter[j - 1] > v) { inter[j] = inter[j - 1]; --j; }
        inter[j] = v;
    }
    for (int k = 0; k + 1 < ni; k += 2) {
        int from = min(inter[k], bw), to = min(inter[k + 1], bw - 1);
        if (from < bw && to >= 0) cover_add(rc, max(0, from), max(0, to), bw);
    }
    // Bresenham edges (imageproc BresenhamLineIter, f32 error term == exact half-integer arithmetic)
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        int x0 = px[e], y0 = py[e], x1 = px[(e + 1) & 3], y1 = py[(e + 1) & 3];
        const bool steep = abs(y1 - y0) > abs(x1 - x0);
        if (steep) { int t = x0; x0 = y0; y0 = t; t = x1; x1 = y1; y1 = t; }
        if (x0 > x1) { int t = x0; x0 = x1; x1 = t; t = y0; y0 = y1; y1 = t; }
        const int dx = x1 - x0, dy = abs(y1 - y0);   // |coordinates| <= a few thousand: all products below fit in int32
        const int ystep = y0 < y1 ? 1 : -1;
        // n_k = number of minor-axis steps before emitting major-axis step k:
        //   n_k = 0 if 2k*dy - dx <= 0 else ceil((2k*dy - dx) / (2dx))
        if (steep) {
            // major axis = canvas y; one pixel on row y if x0 <= y <= x1
            if (y < x0 || y > x1) continue;
            const int k = y - x0, A = 2 * k * dy - dx;
            const int nk = (A <= 0 || dx == 0) ? 0 : (int)((unsigned)(A + 2 * dx - 1) / (unsigned)(2 * dx));
            const int xx = y0 + ystep * nk;
            cover_add(rc, xx, xx, bw);
        } else {
            // major axis = canvas x; row y is hit for k with n_k == t
            const int t = (y - y0) * ystep;
            if (t < 0 || t > dy) continue;
            int klo, khi;
            if (dy == 0) { klo = 0; khi = dx; }
            else {
                klo = (t == 0) ? 0 : (int)((unsigned)(dx * (2 * t - 1)) / (unsigned)(2 * dy)) + 1;
                khi = (int)((unsigned)(dx * (2 * t + 1)) / (unsigned)(2 * dy));   // k_lo(t+1) - 1
                if (khi > dx) khi = dx;
            }
            if (klo > khi) continue;
            cover_add(rc, x0 + klo, x0 + khi, bw);
        }
    }
    // sort by a, merge overlapping and touching intervals (the pixels and their order are what matters)
    for (int i = 1; i < rc.n; ++i) {
        const int va = rc.a[i], vb = rc.b[i];
        int j = i;
        while (j > 0 && rc.a[j - 1] > va) { rc.a[j] = rc.a[j - 1]; rc.b[j] = rc.b[j - 1]; --j; }
        rc.a[j] = va; rc.b[j] = vb;
    }
    int m = 0;
    for (int i = 0; i < rc.n; ++i) {
        if (m > 0 && rc.a[i] <= rc.b[m - 1] + 1) { if (rc.b[i] > rc.b[m - 1]) rc.b[m - 1] = rc.b[i]; }
        else { rc.a[m] = rc.a[i]; rc.b[m] = rc.b[i]; ++m; }
    }
    rc.n = m;
}

// box_score_fast (det_processor.rs:188-221): mean of the probability map over the polygon mask,
// accumulated in the reference's order (raster order, sequential f32).  The warp loads 128 pixels coalesced
// (double buffered), parks them in a warp-private shared-memory line, and every lane replays the same sequential
// add chain from broadcast LDS.128 reads (8 per 32 pixels; a shuffle per pixel would be bound by the 1/clk/SM
// shuffle pipe), so the result is bit-identical to the scalar fold.  sbuf: 128 floats, cov: 32*17 ints, both private
// to the warp.
// Returns false where the reference panics (poly[0] == poly[3]).
static __device__ bool warp_box_score(const float* __restrict__ pred, int h, int w, const int qx[4], const int qy[4], float* score,
                                      float* sbuf, int* cov) {
    const int lane = threadIdx.x & 31;
    if (qx[0] == qx[3] && qy[0] == qy[3]) return false;
    int x_min = min(min(qx[0], qx[1]), min(qx[2], qx[3])), x_max = max(max(qx[0], qx[1]), max(qx[2], qx[3]));
    int y_min = min(min(qy[0], qy[1]), min(qy[2], qy[3])), y_max = max(max(qy[0], qy[1]), max(qy[2], qy[3]));
    x_min = min(max(x_min, 0), w - 1); x_max = min(max(x_max, 0), w - 1);
    y_min = min(max(y_min, 0), h - 1); y_max = min(max(y_max, 0), h - 1);
    const int bw = x_max - x_min + 1, bh = y_max - y_min + 1;
    int px[4], py[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) { px[i] = qx[i] - x_min; py[i] = qy[i] - y_min; }
    float acc = 0.0f;
    unsigned long long count = 0;
    const float4* sb4 = reinterpret_cast<const float4*>(sbuf);
    // row covers of 32 consecutive rows are computed in parallel (one row per lane) into shared memory, then the
    // rows are consumed in raster order
    int* cov_n = cov;                 // [32]
    int* cov_ab = cov + 32;           // [32][16]: a[8] then b[8]
    for (int y_base = 0; y_base < bh; y_base += 32) {
        {
            RowCover rc;
            rc.n = 0;
            if (y_base + lane < bh) polygon_row_cover(px, py, bw, bh, y_base + lane, rc);
            cov_n[lane] = rc.n;
            const float* prow = pred + (size_t)(y_base + lane + y_min) * w + x_min;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                if (k < rc.n) {
                    cov_ab[lane * 16 + k] = rc.a[k]; cov_ab[lane * 16 + 8 + k] = rc.b[k];
                    // the first load of every segment is exposed (its chain cannot start before it lands): pull the rows of this
                    // 32-row group into L2 now, one request per 128-byte line
                    for (int x = rc.a[k]; x <= rc.b[k]; x += 32) asm volatile("prefetch.global.L2 [%0];" ::"l"(prow + x));
                }
            }
        }
        __syncwarp();
        const int nrows = min(32, bh - y_base);
        for (int r = 0; r < nrows; ++r) {
            const int y = y_base + r;
            const int nseg = cov_n[r];
            const float* row = pred + (size_t)(y + y_min) * w + x_min;
            for (int s = 0; s < nseg; ++s) {
                const int xa = cov_ab[r * 16 + s], xb = cov_ab[r * 16 + 8 + s];
                count += (unsigned long long)(xb - xa + 1);
                float v[4], nv[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int xi = xa + 32 * q + lane;
                    v[q] = (xi <= xb) ? __ldg(row + xi) : 0.0f;
                }
                for (int x = xa; x <= xb; x += 128) {
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        sbuf[32 * q + lane] = v[q];
                        const int xi = x + 128 + 32 * q + lane;
                        nv[q] = (xi <= xb) ? __ldg(row + xi) : 0.0f;   // next step's loads fly during this step's chain
                    }
                    __syncwarp();
                    const int n = min(128, xb - x + 1);
                    if (n == 128) {
#pragma unroll
                        for (int k = 0; k < 32; ++k) {
                            const float4 t = sb4[k];
                            acc = __fadd_rn(acc, t.x); acc = __fadd_rn(acc, t.y); acc = __fadd_rn(acc, t.z); acc = __fadd_rn(acc, t.w);
                        }
                    } else {
                        // the slots past the segment's end hold +0.0 (the predicated loads above), and acc + (+0.0) == acc bit for
                        // bit (acc is never -0.0: it starts at +0.0 and x + (-x) rounds to +0.0) — whole groups of 16, no scalar tail
                        const int groups = (n + 15) >> 4;
                        for (int gq = 0; gq < groups; ++gq) {
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                const float4 t = sb4[4 * gq + k];
                                acc = __fadd_rn(acc, t.x); acc = __fadd_rn(acc, t.y); acc = __fadd_rn(acc, t.z); acc = __fadd_rn(acc, t.w);
                            }
                        }
                    }
                    __syncwarp();
#pragma unroll
                    for (int q = 0; q < 4; ++q) v[q] = nv[q];
                }
            }
        }
        __syncwarp();
    }
    *score = count > 0 ? __fdiv_rn(acc, (float)count) : 0.0f;
    return true;
}

// box_score_fast on a map that holds NaN / +-Inf.  The reference's fold is sum += v * (m as f32) over every pixel of the
// bounding box in raster order, then sum / count: a NaN anywhere in the box, or an Inf outside the polygon (Inf * 0),
// makes the sum NaN; Infs inside the polygon make it +-Inf (NaN when both signs occur); finite partial sums cannot
// overflow (|v| <= FLT_MAX is excluded by the flag, ordinary maps hold [0, 1]).  Returns `finite_score` when the box
// holds no non-finite value.  Rare path: lanes take rows, each scans its row serially.
static __device__ __noinline__ float warp_box_score_nonfinite(const float* __restrict__ pred, int h, int w, const int qx[4], const int qy[4], float finite_score) {
    const int lane = threadIdx.x & 31;
    int x_min = min(min(qx[0], qx[1]), min(qx[2], qx[3])), x_max = max(max(qx[0], qx[1]), max(qx[2], qx[3]));
    int y_min = min(min(qy[0], qy[1]), min(qy[2], qy[3])), y_max = max(max(qy[0], qy[1]), max(qy[2], qy[3]));
    x_min = min(max(x_min, 0), w - 1); x_max = min(max(x_max, 0), w - 1);
    y_min = min(max(y_min, 0), h - 1); y_max = min(max(y_max, 0), h - 1);
    const int bw = x_max - x_min + 1, bh = y_max - y_min + 1;
    int px[4], py[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) { px[i] = qx[i] - x_min; py[i] = qy[i] - y_min; }
    unsigned f = 0;   // bit 0: NaN result, bit 1: +Inf inside the polygon, bit 2: -Inf inside the polygon
    for (int y = lane; y < bh; y += 32) {
        RowCover rc;
        rc.n = 0;
        polygon_row_cover(px, py, bw, bh, y, rc);
        const float* row = pred + (size_t)(y + y_min) * w + x_min;
        for (int x = 0; x < bw; ++x) {
            const float v = __ldg(row + x);
            if (fabsf(v) <= 3.402823466e+38f) continue;
            bool m = false;
            for (int k = 0; k < rc.n; ++k) m |= (x >= rc.a[k] && x <= rc.b[k]);
            if (v != v || !m) f |= 1u;
            else f |= v > 0.0f ? 2u : 4u;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) f |= __shfl_xor_sync(0xffffffffu, f, o);
    if (f == 0) return finite_score;
    if ((f & 1u) || (f & 6u) == 6u) return CUDART_NAN_F;
    return (f & 2u) ? CUDART_INF_F : -CUDART_INF_F;
}

// ---- unclip (det_processor.rs:223-252) ---------------------------------------------------------------
// geo 0.30 unsigned_area (f32 shoelace, shifted by the first vertex) and Euclidean length (f32 sum of
// hypotf, restated as (float)sqrt((double)dx*dx + (double)dy*dy)); distance = area * ratio / perimeter.
static __device__ float unclip_distance(const int qx[4], const int qy[4], float ratio) {
    float bx[4], by[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) { bx[i] = (float)qx[i]; by[i] = (float)qy[i]; }
    float tmp = 0.0f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float sx = __fsub_rn(bx[i], bx[0]), sy = __fsub_rn(by[i], by[0]);
        const float ex = __fsub_rn(bx[(i + 1) & 3], bx[0]), ey = __fsub_rn(by[(i + 1) & 3], by[0]);
        tmp = __fadd_rn(tmp, __fsub_rn(__fmul_rn(sx, ey), __fmul_rn(sy, ex)));
    }
    const float area = fabsf(__fdiv_rn(tmp, 2.0f));
    float perim = 0.0f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float dx = __fsub_rn(bx[i], bx[(i + 1) & 3]), dy = __fsub_rn(by[i], by[(i + 1) & 3]);
        const double d2 = __dadd_rn(__dmul_rn((double)dx, (double)dx), __dmul_rn((double)dy, (double)dy));
        perim = __fadd_rn(perim, (float)__dsqrt_rn(d2));
    }
    return __fdiv_rn(__fmul_rn(area, ratio), perim);
}

__device__ __forceinline__ long long clipper_round(double v) { return v < 0 ? (long long)(__dsub_rn(v, 0.5)) : (long long)(__dadd_rn(v, 0.5)); }

// Clipper 6.4.2 ClipperOffset (jtRound, etClosedPolygon) for one quad; serial.  Writes up to max_out
// points; returns the count, 0 when Clipper drops the path (< 3 distinct vertices), -1 on overflow.
static __device__ int clipper_offset_round(const int qx[4], const int qy[4], double delta, double arc_tol, int2* out, int max_out) {
    long long X[4], Y[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) { X[i] = qx[(i + 1) & 3]; Y[i] = qy[(i + 1) & 3]; }  // geo-clipper: ring minus its first point
    int highI = 3;
    while (highI > 0 && X[0] == X[highI] && Y[0] == Y[highI]) --highI;
    long long cx[4], cy[4];
    int len = 0;
    cx[len] = X[0]; cy[len] = Y[0]; ++len;
    for (int i = 1; i <= highI; ++i)
        if (cx[len - 1] != X[i] || cy[len - 1] != Y[i]) { cx[len] = X[i]; cy[len] = Y[i]; ++len; }
    if (len < 3) return 0;
    double a = 0;
    for (int i = 0, j = len - 1; i < len; ++i) {
        a = __dadd_rn(a, __dmul_rn(__dadd_rn((double)cx[j], (double)cx[i]), __dsub_rn((double)cy[j], (double)cy[i])));
        j = i;
    }
    const double area = __dmul_rn(-a, 0.5);
    if (!(area >= 0)) {
        for (int i = 0; i < len / 2; ++i) {
            long long t = cx[i]; cx[i] = cx[len - 1 - i]; cx[len - 1 - i] = t;
            t = cy[i]; cy[i] = cy[len - 1 - i]; cy[len - 1 - i] = t;
        }
    }
    int m = 0;
    if (fabs(delta) < 1.0e-20) {
        for (int i = 0; i < len; ++i) out[m++] = make_int2((int)cx[i], (int)cy[i]);
        return m;
    }
    const double pi = 3.141592653589793238;
    const double two_pi = __dmul_rn(pi, 2.0);
    double yv;
    if (arc_tol <= 0.0) yv = 0.25;
    else if (arc_tol > __dmul_rn(fabs(delta), 0.25)) yv = __dmul_rn(fabs(delta), 0.25);
    else yv = arc_tol;
    double steps = __ddiv_rn(pi, rtm::rt_acos(__dsub_rn(1.0, __ddiv_rn(yv, fabs(delta)))));
    if (steps > __dmul_rn(fabs(delta), pi)) steps = __dmul_rn(fabs(delta), pi);
    double m_sin, m_cos;
    rtm::rt_sincos(__ddiv_rn(two_pi, steps), &m_sin, &m_cos);
    const double steps_per_rad = __ddiv_rn(steps, two_pi);
    if (delta < 0.0) m_sin = -m_sin;
    double nx[4], ny[4];
    for (int j = 0; j < len; ++j) {
        const int j2 = (j + 1) % len;
        double Dx = (double)(cx[j2] - cx[j]), dy = (double)(cy[j2] - cy[j]);
        if (Dx == 0 && dy == 0) { nx[j] = 0; ny[j] = 0; continue; }
        const double f = __ddiv_rn(1.0, __dsqrt_rn(__dadd_rn(__dmul_rn(Dx, Dx), __dmul_rn(dy, dy))));
        Dx = __dmul_rn(Dx, f); dy = __dmul_rn(dy, f);
        nx[j] = dy; ny[j] = -Dx;
    }
#define RT_PUSH(xx, yy)                                                                 \
    do {                                                                                \
        if (m >= max_out) return -1;                                                    \
        out[m++] = make_int2((int)clipper_round(xx), (int)clipper_round(yy));           \
    } while (0)
    int k = len - 1;
    for (int j = 0; j < len; ++j) {
        double sinA = __dsub_rn(__dmul_rn(nx[k], ny[j]), __dmul_rn(nx[j], ny[k]));
        bool done = false;
        if (fabs(__dmul_rn(sinA, delta)) < 1.0) {
            const double cosA = __dadd_rn(__dmul_rn(nx[k], nx[j]), __dmul_rn(ny[j], ny[k]));
            if (cosA > 0) {
                RT_PUSH(__dadd_rn((double)cx[j], __dmul_rn(nx[k], delta)), __dadd_rn((double)cy[j], __dmul_rn(ny[k], delta)));
                done = true;
            }
        } else if (sinA > 1.0) sinA = 1.0;
        else if (sinA < -1.0) sinA = -1.0;
        if (!done) {
            if (__dmul_rn(sinA, delta) < 0) {
                RT_PUSH(__dadd_rn((double)cx[j], __dmul_rn(nx[k], delta)), __dadd_rn((double)cy[j], __dmul_rn(ny[k], delta)));
                if (m >= max_out) return -1;
                out[m++] = make_int2((int)cx[j], (int)cy[j]);
                RT_PUSH(__dadd_rn((double)cx[j], __dmul_rn(nx[j], delta)), __dadd_rn((double)cy[j], __dmul_rn(ny[j], delta)));
            } else {
                const double ang = rtm::rt_atan2(sinA, __dadd_rn(__dmul_rn(nx[k], nx[j]), __dmul_rn(ny[k], ny[j])));
                long long ns = clipper_round(__dmul_rn(steps_per_rad, fabs(ang)));
                const int nsteps = ns < 1 ? 1 : (int)ns;
                double Xv = nx[k], Yv = ny[k];
                for (int i = 0; i < nsteps; ++i) {
                    RT_PUSH(__dadd_rn((double)cx[j], __dmul_rn(Xv, delta)), __dadd_rn((double)cy[j], __dmul_rn(Yv, delta)));
                    const double X2 = Xv;
                    Xv = __dsub_rn(__dmul_rn(Xv, m_cos), __dmul_rn(m_sin, Yv));
                    Yv = __dadd_rn(__dmul_rn(X2, m_sin), __dmul_rn(Yv, m_cos));
                }
                RT_PUSH(__dadd_rn((double)cx[j], __dmul_rn(nx[j], delta)), __dadd_rn((double)cy[j], __dmul_rn(ny[j], delta)));
            }
        }
        k = j;
    }
#undef RT_PUSH
    return m;
}

// The same offset computed by the whole warp: every lane evaluates the shared prologue (orientation, step count, the one
// sincos / acos — SIMT executes the big double-double routines once for the warp), then lane j < len produces the points of corner j
// (its arc is a serial rotation recurrence) at the position an exclusive prefix over the corners' point counts gives it.  Same
// operations in the same order per corner as clipper_offset_round, so the points are bit-identical; all lanes return the count.
__device__ __forceinline__ double sel4d(const double v[4], int i) { return i == 0 ? v[0] : (i == 1 ? v[1] : (i == 2 ? v[2] : v[3])); }
__device__ __forceinline__ long long sel4l(const long long v[4], int i) { return i == 0 ? v[0] : (i == 1 ? v[1] : (i == 2 ? v[2] : v[3])); }
static __device__ int warp_clipper_offset_round(const int qx[4], const int qy[4], double delta, double arc_tol, int2* out, int max_out) {
    const int lane = threadIdx.x & 31;
    long long X[4], Y[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) { X[i] = qx[(i + 1) & 3]; Y[i] = qy[(i + 1) & 3]; }  // geo-clipper: ring minus its first point
    int highI = 3;
    while (highI > 0 && X[0] == X[highI] && Y[0] == Y[highI]) --highI;
    long long cx[4], cy[4];
    int len = 0;
    cx[0] = X[0]; cy[0] = Y[0]; cx[1] = cx[2] = cx[3] = 0; cy[1] = cy[2] = cy[3] = 0;
    len = 1;
#pragma unroll
    for (int i = 1; i < 4; ++i)
        if (i <= highI) {
            const long long px = len == 1 ? cx[0] : (len == 2 ? cx[1] : cx[2]), py = len == 1 ? cy[0] : (len == 2 ? cy[1] : cy[2]);
            if (px != X[i] || py != Y[i]) {
                if (len == 1) { cx[1] = X[i]; cy[1] = Y[i]; } else if (len == 2) { cx[2] = X[i]; cy[2] = Y[i]; } else { cx[3] = X[i]; cy[3] = Y[i]; }
                ++len;
            }
        }
    if (len < 3) return 0;
    double a = 0;
    for (int i = 0, j = len - 1; i < len; ++i) {
        a = __dadd_rn(a, __dmul_rn(__dadd_rn((double)sel4l(cx, j), (double)sel4l(cx, i)), __dsub_rn((double)sel4l(cy, j), (double)sel4l(cy, i))));
        j = i;
    }
    const double area = __dmul_rn(-a, 0.5);
    if (!(area >= 0)) {   // reverse the first `len` entries
        if (len == 3) { long long t = cx[0]; cx[0] = cx[2]; cx[2] = t; t = cy[0]; cy[0] = cy[2]; cy[2] = t; }
        else { long long t = cx[0]; cx[0] = cx[3]; cx[3] = t; t = cy[0]; cy[0] = cy[3]; cy[3] = t; t = cx[1]; cx[1] = cx[2]; cx[2] = t; t = cy[1]; cy[1] = cy[2]; cy[2] = t; }
    }
    if (fabs(delta) < 1.0e-20) {
        if (lane < len) out[lane] = make_int2((int)sel4l(cx, lane), (int)sel4l(cy, lane));
        return len;
    }
    const double pi = 3.141592653589793238;
    const double two_pi = __dmul_rn(pi, 2.0);
    double yv;
    if (arc_tol <= 0.0) yv = 0.25;
    else if (arc_tol > __dmul_rn(fabs(delta), 0.25)) yv = __dmul_rn(fabs(delta), 0.25);
    else yv = arc_tol;
    double steps = __ddiv_rn(pi, rtm::rt_acos(__dsub_rn(1.0, __ddiv_rn(yv, fabs(delta)))));
    if (steps > __dmul_rn(fabs(delta), pi)) steps = __dmul_rn(fabs(delta), pi);
    double m_sin, m_cos;
    rtm::rt_sincos(__ddiv_rn(two_pi, steps), &m_sin, &m_cos);
    const double steps_per_rad = __ddiv_rn(steps, two_pi);
    if (delta < 0.0) m_sin = -m_sin;
    // lane j: its corner (j) and the previous one (k); unit normals of the edges leaving them
    const int j = lane < len ? lane : 0, k = (j + len - 1) % len;
    auto normal = [&](int e, double& nxe, double& nye) {
        const int e2 = (e + 1) % len;
        double Dx = (double)(sel4l(cx, e2) - sel4l(cx, e)), dy = (double)(sel4l(cy, e2) - sel4l(cy, e));
        if (Dx == 0 && dy == 0) { nxe = 0; nye = 0; return; }
        const double f = __ddiv_rn(1.0, __dsqrt_rn(__dadd_rn(__dmul_rn(Dx, Dx), __dmul_rn(dy, dy))));
        Dx = __dmul_rn(Dx, f); dy = __dmul_rn(dy, f);
        nxe = dy; nye = -Dx;
    };
    double nxj, nyj, nxk, nyk;
    normal(j, nxj, nyj);
    normal(k, nxk, nyk);
    const double pxj = (double)sel4l(cx, j), pyj = (double)sel4l(cy, j);
    double sinA = __dsub_rn(__dmul_rn(nxk, nyj), __dmul_rn(nxj, nyk));
    int mode = 2;   // 0: one point, 1: concave patch (3 points), 2: round join
    if (fabs(__dmul_rn(sinA, delta)) < 1.0) {
        const double cosA = __dadd_rn(__dmul_rn(nxk, nxj), __dmul_rn(nyj, nyk));
        if (cosA > 0) mode = 0;
    } else if (sinA > 1.0) sinA = 1.0;
    else if (sinA < -1.0) sinA = -1.0;
    if (mode == 2 && __dmul_rn(sinA, delta) < 0) mode = 1;
    int nsteps = 0;
    if (mode == 2) {
        const double ang = rtm::rt_atan2(sinA, __dadd_rn(__dmul_rn(nxk, nxj), __dmul_rn(nyk, nyj)));
        const long long ns = clipper_round(__dmul_rn(steps_per_rad, fabs(ang)));
        nsteps = ns < 1 ? 1 : (ns > 100000 ? 100000 : (int)ns);
    }
    int cnt = lane < len ? (mode == 0 ? 1 : (mode == 1 ? 3 : nsteps + 1)) : 0;
    int incl = cnt;
#pragma unroll
    for (int d = 1; d < 4; d <<= 1) { const int v = __shfl_up_sync(RT_FULL, incl, d); if (lane >= d) incl += v; }
    const int total = __shfl_sync(RT_FULL, incl, 3);
    if (total > max_out) return -1;
    int m = incl - cnt;
    if (lane < len) {
#define RT_PUT(xx, yy) out[m++] = make_int2((int)clipper_round(xx), (int)clipper_round(yy))
        if (mode == 0) RT_PUT(__dadd_rn(pxj, __dmul_rn(nxk, delta)), __dadd_rn(pyj, __dmul_rn(nyk, delta)));
        else if (mode == 1) {
            RT_PUT(__dadd_rn(pxj, __dmul_rn(nxk, delta)), __dadd_rn(pyj, __dmul_rn(nyk, delta)));
            out[m++] = make_int2((int)sel4l(cx, j), (int)sel4l(cy, j));
            RT_PUT(__dadd_rn(pxj, __dmul_rn(nxj, delta)), __dadd_rn(pyj, __dmul_rn(nyj, delta)));
        } else {
            double Xv = nxk, Yv = nyk;
            for (int i = 0; i < nsteps; ++i) {
                RT_PUT(__dadd_rn(pxj, __dmul_rn(Xv, delta)), __dadd_rn(pyj, __dmul_rn(Yv, delta)));
                const double X2 = Xv;
                Xv = __dsub_rn(__dmul_rn(Xv, m_cos), __dmul_rn(m_sin, Yv));
                Yv = __dadd_rn(__dmul_rn(X2, m_sin), __dmul_rn(Yv, m_cos));
            }
            RT_PUT(__dadd_rn(pxj, __dmul_rn(nxj, delta)), __dadd_rn(pyj, __dmul_rn(nyj, delta)));
        }
#undef RT_PUT
    }
    return total;
}

// points.rs:179-194
__device__ __forceinline__ float scale_clip_1(float v, double inv, double ori) {
    double x1 = round(__dmul_rn((double)v, inv));  // round(): half away from zero
    x1 = x1 < 0.0 ? 0.0 : (x1 > ori - 1.0 ? ori - 1.0 : x1);
    return (float)x1;
}
// points.rs:125-169
__device__ __forceinline__ float side_len(float ax, float ay, float bx, float by) {
    const double dx = (double)__fsub_rn(ax, bx), dy = (double)__fsub_rn(ay, by);
    return (float)__dsqrt_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));
}
