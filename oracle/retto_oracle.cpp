// retto_oracle.cpp — CPU ORACLE.  TEST INFRASTRUCTURE ONLY.
//
// A scalar C++17 restatement of the reference's (NekoImageLand/retto, retto-core) CPU image path:
// det preprocess -> DB postprocess -> rotate-crop -> cls/rec batch build -> cls flip -> CTC decode.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
// load this library; the product (libretto_b200.so) never links, imports or calls it.
//
// PARITY UNPINNED for the third-party arithmetic: the reference cannot be compiled here (no
// cargo/rustc; `image`, `imageproc`, `geo`, `geo-clipper`/Clipper, `ndarray-stats` are un-vendored
// Cargo.lock dependencies absent from /root/reference) and its own tests hold no golden vectors
// (session.rs:206-255 need network + ORT).  Functions marked RECALLED restate the published
// algorithm of the named crate@version from memory; each is isolated so it can be corrected in one
// place.  Cross-checks that ARE possible here (cv2.dilate, cv2.findContours point sets,
// cv2.convexHull, numpy normalise/argmax/CTC) live in tests/test_oracle_*.py.
//
// Every function cites the reference file:line it follows (paths relative to retto-core/src/).
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <string>
#include <vector>

namespace orc {

// ------------------------------------------------------------------------------------------
// libm.  The restatement calls the host glibc — what the reference's Rust f64::atan2/sin/cos/acos call on
// Linux — and includes NO product source.  orc_set_trig_hooks lets a test install other implementations for a
// DIAGNOSTIC second run (oracle/diag/crmath_shim.cpp wraps the CUDA path's correctly-rounded routines); every
// asserted parity comparison runs with the hooks cleared (tests/conftest.py resets them around each test).
typedef double (*trig_atan2_fn)(double, double);
typedef void (*trig_sincos_fn)(double, double*, double*);
typedef double (*trig_acos_fn)(double);
static trig_atan2_fn g_hook_atan2 = nullptr;
static trig_sincos_fn g_hook_sincos = nullptr;
static trig_acos_fn g_hook_acos = nullptr;
static inline double m_atan2(double y, double x) { return g_hook_atan2 ? g_hook_atan2(y, x) : std::atan2(y, x); }
static inline void m_sincos(double a, double* s, double* c) {
    if (g_hook_sincos) g_hook_sincos(a, s, c);
    else { *s = std::sin(a); *c = std::cos(a); }
}
static inline double m_acos(double v) { return g_hook_acos ? g_hook_acos(v) : std::acos(v); }

static inline float round_half_away_f(float v) { return roundf(v); }      // Rust f32::round
static inline double round_half_away_d(double v) { return std::round(v); }  // Rust f64::round

struct PtI { int x, y; };
struct PtD { double x, y; };

// ------------------------------------------------------------------------------------------
// RECALLED image 0.25.6 imageops::thumbnail  (call sites image_helper.rs:124,139,168,184)
// Box filter with fractional branches for upscaling.  src/dst are HWC u8 with C channels.
static inline float fractf_(float v) { return v - truncf(v); }
static inline uint32_t clampu(uint32_t a, uint32_t lo, uint32_t hi) { return a < lo ? lo : (a > hi ? hi : a); }
static inline uint8_t f32_to_u8_numcast(float v) {
    // num-traits NumCast f32->u8: Some(trunc) when -1 < v < 256 else None (reference would panic)
    if (!(v > -1.0f && v < 256.0f)) return v >= 256.0f ? 255 : 0;
    return (uint8_t)v;
}

static void thumbnail(const uint8_t* src, int h, int w, int C, uint8_t* dst, int nh, int nw) {
    if (h == 0 || w == 0 || nh == 0 || nw == 0) return;
    const float x_ratio = (float)w / (float)nw;
    const float y_ratio = (float)h / (float)nh;
    auto px = [&](uint32_t x, uint32_t y, int c) -> uint32_t { return src[((size_t)y * w + x) * C + c]; };
    for (int outy = 0; outy < nh; ++outy) {
        const float bottomf = (float)outy * y_ratio;
        const float topf = bottomf + y_ratio;
        const uint32_t bottom = clampu((uint32_t)ceilf(bottomf), 0, (uint32_t)h - 1);
        const uint32_t top = clampu((uint32_t)ceilf(topf), bottom, (uint32_t)h);
        for (int outx = 0; outx < nw; ++outx) {
            const float leftf = (float)outx * x_ratio;
            const float rightf = leftf + x_ratio;
            const uint32_t left = clampu((uint32_t)ceilf(leftf), 0, (uint32_t)w - 1);
            const uint32_t right = clampu((uint32_t)ceilf(rightf), left, (uint32_t)w);
            uint8_t* o = dst + ((size_t)outy * nw + outx) * C;
            if (bottom != top && left != right) {
                const uint32_t n = (right - left) * (top - bottom);
                for (int c = 0; c < C; ++c) {
                    uint32_t sum = 0;
                    for (uint32_t y = bottom; y < top; ++y)
                        for (uint32_t x = left; x < right; ++x) sum += px(x, y, c);
                    uint32_t v = (sum + n / 2) / n;
                    o[c] = (uint8_t)(v > 255 ? 255 : v);
                }
            } else if (bottom != top) {  // left == right: horizontal fraction
                const float fract = (fractf_(leftf) + fractf_(rightf)) / 2.0f;
                const uint32_t l = right - 1;
                const uint32_t l1 = std::min<uint32_t>(l + 1, (uint32_t)w - 1);  // clamp: flagged, see SURVEY B.1
                const float cnt = (float)(top - bottom);
                const float fact_right = fract / cnt;
                const float fact_left = (1.0f - fract) / cnt;
                for (int c = 0; c < C; ++c) {
                    uint32_t sl = 0, sr = 0;
                    for (uint32_t y = bottom; y < top; ++y) { sl += px(l, y, c); sr += px(l1, y, c); }
                    o[c] = f32_to_u8_numcast(fact_left * (float)sl + fact_right * (float)sr);
                }
            } else if (left != right) {  // bottom == top: vertical fraction
                const float fract = (fractf_(topf) + fractf_(bottomf)) / 2.0f;
                const uint32_t b = top - 1;
                const uint32_t b1 = std::min<uint32_t>(b + 1, (uint32_t)h - 1);
                const float cnt = (float)(right - left);
                const float fact_top = fract / cnt;
                const float fact_bot = (1.0f - fract) / cnt;
                for (int c = 0; c < C; ++c) {
                    uint32_t sb = 0, st = 0;
                    for (uint32_t x = left; x < right; ++x) { sb += px(x, b, c); st += px(x, b1, c); }
                    o[c] = f32_to_u8_numcast(fact_bot * (float)sb + fact_top * (float)st);
                }
            } else {  // both empty
                const float frac_v = (fractf_(topf) + fractf_(bottomf)) / 2.0f;
                const float frac_h = (fractf_(leftf) + fractf_(rightf)) / 2.0f;
                const uint32_t l = right - 1, b = top - 1;
                const uint32_t l1 = std::min<uint32_t>(l + 1, (uint32_t)w - 1);
                const uint32_t b1 = std::min<uint32_t>(b + 1, (uint32_t)h - 1);
                const float fact_tr = frac_v * frac_h;
                const float fact_tl = frac_v * (1.0f - frac_h);
                const float fact_br = (1.0f - frac_v) * frac_h;
                const float fact_bl = (1.0f - frac_v) * (1.0f - frac_h);
                for (int c = 0; c < C; ++c) {
                    const float k_bl = (float)px(l, b, c), k_tl = (float)px(l, b1, c);
                    const float k_br = (float)px(l1, b, c), k_tr = (float)px(l1, b1, c);
                    o[c] = f32_to_u8_numcast(fact_br * k_br + fact_tr * k_tr + fact_bl * k_bl + fact_tl * k_tl);
                }
            }
        }
    }
}

// image_helper.rs:106-148  ImageHelper::resize_both — output dims only (returns 1 if a resize happens)
// NOTE both branches can fire in sequence; each uses the ORIGINAL (ori_h, ori_w) for its size math.
static int resize_both_plan(int ori_h, int ori_w, int max_len, int min_len, int dims[4]) {
    int n = 0;
    const float h = (float)ori_h, w = (float)ori_w;
    if (std::max(ori_h, ori_w) > max_len) {
        const float scale = (float)max_len / std::max(h, w);
        uint32_t rh = std::max<uint32_t>((uint32_t)floorf(h * scale) / 32u, 1u) * 32u;
        uint32_t rw = std::max<uint32_t>((uint32_t)floorf(w * scale) / 32u, 1u) * 32u;
        dims[2 * n] = (int)rh; dims[2 * n + 1] = (int)rw; ++n;
    }
    if (std::min(ori_h, ori_w) < min_len) {
        const float scale = (float)min_len / std::min(h, w);
        uint32_t rh = (uint32_t)round_half_away_f(floorf(h * scale) / 32.0f) * 32u;
        uint32_t rw = (uint32_t)round_half_away_f(floorf(w * scale) / 32.0f) * 32u;
        dims[2 * n] = (int)rh; dims[2 * n + 1] = (int)rw; ++n;
    }
    return n;
}

// image_helper.rs:150-174  ImageHelper::resize_either — output dims. limit_type 0 = Min, 1 = Max
static void resize_either_plan(int h, int w, int limit_type, int limit_len, int* oh, int* ow) {
    float ratio = 1.0f;
    if (limit_type == 1) {
        if (std::max(w, h) > limit_len) ratio = (float)limit_len / (float)std::max(w, h);
    } else {
        if (std::min(w, h) < limit_len) ratio = (float)limit_len / (float)std::min(w, h);
    }
    *oh = (int)((uint32_t)round_half_away_f(floorf((float)h * ratio) / 32.0f) * 32u);
    *ow = (int)((uint32_t)round_half_away_f(floorf((float)w * ratio) / 32.0f) * 32u);
}

// det_processor.rs:256-274 preprocess: copy -> resize_either -> rgb2bgr (image_helper.rs:211-221)
// -> normalize (det_processor.rs:151-155: three separately rounded f32 ops) -> permute(2,0,1)
// (:157-160) -> batch axis; NCHW materialised by as_standard_layout (ort_worker.rs:191).
static void det_preprocess(const uint8_t* rgb, int h, int w, int limit_type, int limit_len, float scale,
                           const float mean[3], const float stdv[3], float* out, int oh, int ow) {
    std::vector<uint8_t> rs((size_t)oh * ow * 3);
    thumbnail(rgb, h, w, 3, rs.data(), oh, ow);
    (void)limit_type; (void)limit_len;
    const size_t plane = (size_t)oh * ow;
    for (size_t i = 0; i < plane; ++i) {
        for (int c = 0; c < 3; ++c) {
            const uint8_t v = rs[i * 3 + (2 - c)];  // BGR: channel 0 of the tensor is B
            float f = (float)v;
            f = f * scale;
            f = f - mean[c];
            f = f / stdv[c];
            out[c * plane + i] = f;
        }
    }
}

// det_processor.rs:284-292: threshold (strict >; NaN -> 0) then grayscale_dilate with the 2x2 mask,
// RECALLED imageproc 0.25.0 morphology::{Mask::from_image, grayscale_dilate}: offsets
// {(-1,-1),(0,-1),(-1,0),(0,0)}, out-of-image taps skipped (== cv2.dilate 2x2, checked in tests).
static void threshold_dilate(const float* pred, int h, int w, float thr, int dilate, uint8_t* out) {
    std::vector<uint8_t> m((size_t)h * w);
    for (size_t i = 0; i < (size_t)h * w; ++i) m[i] = pred[i] > thr ? 255 : 0;
    if (!dilate) { memcpy(out, m.data(), m.size()); return; }
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) {
            uint8_t v = m[(size_t)y * w + x];
            if (x > 0) v = std::max(v, m[(size_t)y * w + x - 1]);
            if (y > 0) v = std::max(v, m[(size_t)(y - 1) * w + x]);
            if (x > 0 && y > 0) v = std::max(v, m[(size_t)(y - 1) * w + x - 1]);
            out[(size_t)y * w + x] = v;
        }
}

// ------------------------------------------------------------------------------------------
// RECALLED imageproc 0.25.0 contours::find_contours::<i32>  (call site det_processor.rs:293)
// Suzuki-Abe border following, 8-connected foreground, outer AND hole borders, discovery order.
// quirk_x0: imageproc's scan only starts an OUTER border when x > 0 and a HOLE border when
// x + 1 < width (recalled; the image frame is otherwise treated as background).  With the quirk a
// blob touching column 0 is discovered at the right end of its first run as a "hole"-typed border
// with the same point set; retto ignores border_type (det_processor.rs:293-295), so only the
// discovery position changes.  quirk_x0 = 0 gives textbook Suzuki-Abe (== cv2.findContours sets).
struct Contour { std::vector<PtI> pts; int is_hole; };

static void find_contours(const uint8_t* mask, int height, int width, int quirk_x0, std::vector<Contour>& out) {
    out.clear();
    std::vector<int32_t> img((size_t)height * width, 0);
    for (size_t i = 0; i < img.size(); ++i) img[i] = mask[i] > 0 ? 1 : 0;
    auto at = [&](int x, int y) -> int32_t& { return img[(size_t)y * width + x]; };
    // ring in clockwise order (image coords): W NW N NE E SE S SW
    static const int RX[8] = {-1, -1, 0, 1, 1, 1, 0, -1};
    static const int RY[8] = {0, -1, -1, -1, 0, 1, 1, 1};
    auto dir_of = [&](int dx, int dy) { for (int i = 0; i < 8; ++i) if (RX[i] == dx && RY[i] == dy) return i; return -1; };
    auto nz = [&](int x, int y) { return x >= 0 && x < width && y >= 0 && y < height && img[(size_t)y * width + x] != 0; };
    int curr_border = 1;
    for (int y = 0; y < height; ++y) {
        for (int x = 0; x < width; ++x) {
            if (at(x, y) == 0) continue;
            int adjx = 0, adjy = 0, type = -1;
            const bool left_zero = quirk_x0 ? (x > 0 && at(x - 1, y) == 0) : (x == 0 || at(x - 1, y) == 0);
            const bool right_zero = quirk_x0 ? (x + 1 < width && at(x + 1, y) == 0) : (x + 1 == width || at(x + 1, y) == 0);
            if (at(x, y) == 1 && left_zero) { adjx = x - 1; adjy = y; type = 0; }
            else if (at(x, y) > 0 && right_zero) { adjx = x + 1; adjy = y; type = 1; }
            if (type < 0) continue;
            ++curr_border;
            Contour ct; ct.is_hole = type;
            int start = dir_of(adjx - x, adjy - y);
            // clockwise search from adj for the first non-zero neighbour
            int p1x = 0, p1y = 0; bool found = false;
            for (int k = 0; k < 8; ++k) {
                int d = (start + k) & 7;
                if (nz(x + RX[d], y + RY[d])) { p1x = x + RX[d]; p1y = y + RY[d]; found = true; break; }
            }
            if (!found) {
                ct.pts.push_back(PtI{x, y});
                at(x, y) = -curr_border;
            } else {
                int p2x = p1x, p2y = p1y, p3x = x, p3y = y;
                for (;;) {
                    ct.pts.push_back(PtI{p3x, p3y});
                    int d0 = dir_of(p2x - p3x, p2y - p3y);
                    // counter-clockwise search starting one step ccw of pos2, pos2 itself last
                    int p4x = 0, p4y = 0, d4 = -1;
                    bool right_examined = false;
                    for (int k = 1; k <= 8; ++k) {
                        int d = (d0 - k) & 7;
                        if (nz(p3x + RX[d], p3y + RY[d])) { p4x = p3x + RX[d]; p4y = p3y + RY[d]; d4 = d; break; }
                        if (d == 4) right_examined = true;  // E examined and found zero
                    }
                    (void)d4;
                    if (p3x + 1 == width || right_examined) at(p3x, p3y) = -curr_border;
                    else if (at(p3x, p3y) == 1) at(p3x, p3y) = curr_border;
                    if (p4x == x && p4y == y && p3x == p1x && p3y == p1y) break;
                    p2x = p3x; p2y = p3y; p3x = p4x; p3y = p4y;
                }
            }
            out.push_back(std::move(ct));
        }
    }
}

// ------------------------------------------------------------------------------------------
// RECALLED imageproc 0.25.0 geometry::{convex_hull, min_area_rect}  (det_processor.rs:180)
static inline long long orient_val(PtI p, PtI q, PtI r) {
    return (long long)(q.y - p.y) * (r.x - q.x) - (long long)(q.x - p.x) * (r.y - q.y);
}
// Graham scan as imageproc does it: pivot = lowest y then lowest x; others sorted by orientation
// about the pivot (CounterClockwise first; collinear: nearer first); collinear-with-pivot runs keep
// the farthest; stack pops while turn != CounterClockwise.  Result: strict hull vertices starting at
// the pivot, in the direction of negative orient_val.
static std::vector<PtI> convex_hull(const std::vector<PtI>& in) {
    std::vector<PtI> hull;
    if (in.empty()) return hull;
    std::vector<PtI> pts = in;
    size_t sp = 0;
    PtI start = pts[0];
    for (size_t i = 1; i < pts.size(); ++i)
        if (pts[i].y < start.y || (pts[i].y == start.y && pts[i].x < start.x)) { sp = i; start = pts[i]; }
    std::swap(pts[0], pts[sp]);
    pts.erase(pts.begin());
    auto dist2 = [&](PtI a) { double dx = (double)a.x - start.x, dy = (double)a.y - start.y; return dx * dx + dy * dy; };
    std::stable_sort(pts.begin(), pts.end(), [&](const PtI& a, const PtI& b) {
        long long v = orient_val(start, a, b);
        if (v == 0) return dist2(a) < dist2(b);
        return v < 0;  // CounterClockwise => Less
    });
    std::vector<PtI> rem;
    for (size_t i = 0; i < pts.size();) {
        PtI p = pts[i++];
        while (i < pts.size() && orient_val(start, p, pts[i]) == 0) p = pts[i++];
        rem.push_back(p);
    }
    hull.push_back(start);
    for (const PtI& p : rem) {
        while (hull.size() > 1 && !(orient_val(hull[hull.size() - 2], hull[hull.size() - 1], p) < 0)) hull.pop_back();
        hull.push_back(p);
    }
    return hull;
}

// rotating calipers over hull.windows(2) edges (the closing edge last->first is NOT visited:
// recalled `points.windows(2)`; switchable for the record via g_calipers_wrap).
static int g_calipers_wrap = 0;
static void rotating_calipers(const std::vector<PtI>& hull, double out_xy[8], int floor_ceil_to_int) {
    const double PI = 3.14159265358979323846;
    std::vector<double> angles;
    const size_t n = hull.size();
    const size_t ne = g_calipers_wrap ? n : n - 1;
    for (size_t i = 0; i < ne; ++i) {
        const PtI a = hull[i], b = hull[(i + 1) % n];
        const double ex = (double)b.x - (double)a.x, ey = (double)b.y - (double)a.y;
        double ang = std::fabs(std::fmod(m_atan2(ey, ex) + PI, PI / 2.0));
        if (angles.empty() || angles.back() != ang) angles.push_back(ang);  // Vec::dedup
    }
    double min_area = 1.7976931348623157e308;
    PtD res[4] = {{0, 0}, {0, 0}, {0, 0}, {0, 0}};
    for (double ang : angles) {
        double s, c;
        m_sincos(ang, &s, &c);
        double min_x = 1.7976931348623157e308, max_x = -1.7976931348623157e308;
        double min_y = 1.7976931348623157e308, max_y = -1.7976931348623157e308;
        for (const PtI& p : hull) {
            const double px = (double)p.x, py = (double)p.y;
            const double rx = px * c + py * s;   // Point::rotate
            const double ry = py * c - px * s;
            min_x = std::fmin(min_x, rx); max_x = std::fmax(max_x, rx);
            min_y = std::fmin(min_y, ry); max_y = std::fmax(max_y, ry);
        }
        const double area = (max_x - min_x) * (max_y - min_y);
        if (area < min_area) {
            min_area = area;
            auto inv = [&](double x, double y) { return PtD{x * c - y * s, y * c + x * s}; };  // invert_rotation
            res[0] = inv(max_x, min_y);
            res[1] = inv(min_x, min_y);
            res[2] = inv(min_x, max_y);
            res[3] = inv(max_x, max_y);
        }
    }
    std::stable_sort(res, res + 4, [](const PtD& a, const PtD& b) { return a.x < b.x; });
    const int i1 = res[1].y > res[0].y ? 0 : 1;
    const int i2 = res[3].y > res[2].y ? 2 : 3;
    const int i3 = res[3].y > res[2].y ? 3 : 2;
    const int i4 = res[1].y > res[0].y ? 1 : 0;
    (void)floor_ceil_to_int;
    out_xy[0] = std::floor(res[i1].x); out_xy[1] = std::floor(res[i1].y);
    out_xy[2] = std::ceil(res[i2].x);  out_xy[3] = std::floor(res[i2].y);
    out_xy[4] = std::ceil(res[i3].x);  out_xy[5] = std::ceil(res[i3].y);
    out_xy[6] = std::floor(res[i4].x); out_xy[7] = std::ceil(res[i4].y);
}

// returns hull size class: 0 => panic in the reference ("no points are defined")
static int min_area_rect(const std::vector<PtI>& pts, double out_xy[8]) {
    std::vector<PtI> hull = convex_hull(pts);
    if (hull.empty()) return 0;
    if (hull.size() == 1) { for (int i = 0; i < 4; ++i) { out_xy[2 * i] = hull[0].x; out_xy[2 * i + 1] = hull[0].y; } return 1; }
    if (hull.size() == 2) {
        const PtI q[4] = {hull[0], hull[1], hull[1], hull[0]};
        for (int i = 0; i < 4; ++i) { out_xy[2 * i] = q[i].x; out_xy[2 * i + 1] = q[i].y; }
        return 2;
    }
    rotating_calipers(hull, out_xy, 1);
    return (int)hull.size();
}

// det_processor.rs:166-186  euclid_dist (f32) and sside
static inline float euclid_dist_f32(float ax, float ay, float bx, float by) {
    const float dx = ax - bx, dy = ay - by;
    const float a = dx * dx, b = dy * dy;
    return sqrtf(a + b);
}
static inline float sside_of(const double q[8]) {
    const float s1 = euclid_dist_f32((float)q[0], (float)q[1], (float)q[2], (float)q[3]);  // tl-tr
    const float s2 = euclid_dist_f32((float)q[6], (float)q[7], (float)q[4], (float)q[5]);  // bl-br
    return std::fmin(s1, s2);
}

// ------------------------------------------------------------------------------------------
// RECALLED imageproc 0.25.0 drawing::{draw_polygon_mut, draw_line_segment_mut, BresenhamLineIter}
// returns -1 where the reference panics (poly[0] == poly[last]).
static int draw_polygon_mask(uint8_t* canvas, int cw, int ch, const PtI poly[4]) {
    if (poly[0].x == poly[3].x && poly[0].y == poly[3].y) return -1;
    int y_min = INT32_MAX, y_max = INT32_MIN;
    for (int i = 0; i < 4; ++i) { y_min = std::min(y_min, poly[i].y); y_max = std::max(y_max, poly[i].y); }
    y_min = std::max(0, std::min(y_min, ch - 1));
    y_max = std::max(0, std::min(y_max, ch - 1));
    std::vector<int> inter;
    for (int y = y_min; y <= y_max; ++y) {
        inter.clear();
        for (int e = 0; e < 4; ++e) {
            const PtI p0 = poly[e], p1 = poly[(e + 1) & 3];
            if ((p0.y <= y && p1.y >= y) || (p1.y <= y && p0.y >= y)) {
                if (p0.y == p1.y) { inter.push_back(p0.x); inter.push_back(p1.x); }
                else if (p0.y == y || p1.y == y) {
                    if (p1.y > y) inter.push_back(p0.x);
                    if (p0.y > y) inter.push_back(p1.x);
                } else {
                    const float fraction = (float)(y - p0.y) / (float)(p1.y - p0.y);
                    const float t = fraction * (float)(p1.x - p0.x);
                    const float v = (float)p0.x + t;
                    inter.push_back((int)round_half_away_f(v));
                }
            }
        }
        std::sort(inter.begin(), inter.end());
        for (size_t k = 0; k + 1 < inter.size(); k += 2) {  // odd tail would panic in the reference
            int from = std::min(inter[k], cw);
            int to = std::min(inter[k + 1], cw - 1);
            if (from < cw && to >= 0) {
                from = std::max(0, from); to = std::max(0, to);
                for (int x = from; x <= to; ++x) canvas[(size_t)y * cw + x] = 1;
            }
        }
    }
    for (int e = 0; e < 4; ++e) {
        float x0 = (float)poly[e].x, y0 = (float)poly[e].y;
        float x1 = (float)poly[(e + 1) & 3].x, y1 = (float)poly[(e + 1) & 3].y;
        const bool steep = fabsf(y1 - y0) > fabsf(x1 - x0);
        if (steep) { std::swap(x0, y0); std::swap(x1, y1); }
        if (x0 > x1) { std::swap(x0, x1); std::swap(y0, y1); }
        const float dx = x1 - x0, dy = fabsf(y1 - y0);
        int x = (int)x0, y = (int)y0;
        float error = dx / 2.0f;
        const int end_x = (int)x1, y_step = y0 < y1 ? 1 : -1;
        while (x <= end_x) {
            const int px = steep ? y : x, py = steep ? x : y;
            if (px >= 0 && px < cw && py >= 0 && py < ch) canvas[(size_t)py * cw + px] = 1;
            ++x;
            error -= dy;
            if (error < 0.0f) { y += y_step; error += dx; }
        }
    }
    return 0;
}

// det_processor.rs:188-221  box_score_fast (scores the probability map; sequential f32 fold)
static int box_score_fast(const float* pred, int h, int w, const PtI q[4], float* score) {
    int x_min = INT32_MAX, x_max = INT32_MIN, y_min = INT32_MAX, y_max = INT32_MIN;
    for (int i = 0; i < 4; ++i) {
        x_min = std::min(x_min, q[i].x); x_max = std::max(x_max, q[i].x);
        y_min = std::min(y_min, q[i].y); y_max = std::max(y_max, q[i].y);
    }
    x_min = std::min(std::max(x_min, 0), w - 1); x_max = std::min(std::max(x_max, 0), w - 1);
    y_min = std::min(std::max(y_min, 0), h - 1); y_max = std::min(std::max(y_max, 0), h - 1);
    const int bw = x_max - x_min + 1, bh = y_max - y_min + 1;
    PtI poly[4];
    for (int i = 0; i < 4; ++i) poly[i] = PtI{q[i].x - x_min, q[i].y - y_min};
    std::vector<uint8_t> mask((size_t)bw * bh, 0);
    if (draw_polygon_mask(mask.data(), bw, bh, poly) < 0) return -1;
    float sum = 0.0f;
    size_t count = 0;
    for (int y = 0; y < bh; ++y)
        for (int x = 0; x < bw; ++x) {
            const size_t m = mask[(size_t)y * bw + x];
            const float v = pred[(size_t)(y + y_min) * w + (x + x_min)];
            const float t = v * (float)m;
            sum = sum + t;
            count += m;
        }
    *score = count > 0 ? sum / (float)count : 0.0f;
    return 0;
}

// ------------------------------------------------------------------------------------------
// det_processor.rs:223-252 unclip.
// RECALLED geo 0.30.0 Area::unsigned_area (shoelace on coordinates shifted by the first vertex, f32)
// and Euclidean.length (sum of hypot per segment, f32; Rust f32::hypot -> hypotf, restated as
// (float)sqrt((double)dx*dx + (double)dy*dy) as glibc >= 2.35 computes it).
// RECALLED geo-clipper 0.9.0 + clipper-sys 0.8.0 (Clipper 6.4.2) ClipperOffset with jtRound,
// arc tolerance 0.5, etClosedPolygon, factor 1.0.  The trailing Clipper union only removes
// duplicate/collinear vertices of the (already simple, convex-source) outline and the caller only
// takes the convex hull of the returned points, so the raw offset points are returned.
static inline long long clipper_round(double v) { return v < 0 ? (long long)(v - 0.5) : (long long)(v + 0.5); }

static float unclip_distance(const float bx[4], const float by[4], float unclip_ratio) {
    // ring = p0 p1 p2 p3 p0 ; shift by p0
    float tmp = 0.0f;
    for (int i = 0; i < 4; ++i) {
        const float sx = bx[i] - bx[0], sy = by[i] - by[0];
        const float ex = bx[(i + 1) & 3] - bx[0], ey = by[(i + 1) & 3] - by[0];
        const float a = sx * ey, b = sy * ex;
        const float det = a - b;
        tmp = tmp + det;
    }
    float area = tmp / 2.0f;
    area = fabsf(area);
    float perim = 0.0f;
    for (int i = 0; i < 4; ++i) {
        const float dx = bx[i] - bx[(i + 1) & 3], dy = by[i] - by[(i + 1) & 3];
        const float l = (float)std::sqrt((double)dx * (double)dx + (double)dy * (double)dy);
        perim = perim + l;
    }
    const float t = area * unclip_ratio;
    return t / perim;
}

static void clipper_offset_round(const long long px_in[4], const long long py_in[4], double delta, double arc_tol,
                                 std::vector<PtI>& out) {
    out.clear();
    // geo-clipper feeds the closed ring minus its first point: p1 p2 p3 p0
    long long X[4], Y[4];
    for (int i = 0; i < 4; ++i) { X[i] = px_in[(i + 1) & 3]; Y[i] = py_in[(i + 1) & 3]; }
    // AddPath: strip closing duplicates and consecutive duplicates
    int highI = 3;
    while (highI > 0 && X[0] == X[highI] && Y[0] == Y[highI]) --highI;
    std::vector<long long> cx, cy;
    cx.push_back(X[0]); cy.push_back(Y[0]);
    for (int i = 1; i <= highI; ++i)
        if (cx.back() != X[i] || cy.back() != Y[i]) { cx.push_back(X[i]); cy.push_back(Y[i]); }
    int len = (int)cx.size();
    if (len < 3) return;  // etClosedPolygon && j < 2 -> path dropped -> empty result
    // FixOrientations: reverse when Area < 0   (Area = -0.5 * sum (xj + xi)(yj - yi), j = previous)
    double a = 0;
    for (int i = 0, j = len - 1; i < len; ++i) { a += ((double)cx[j] + (double)cx[i]) * ((double)cy[j] - (double)cy[i]); j = i; }
    const double area = -a * 0.5;
    if (!(area >= 0)) { std::reverse(cx.begin(), cx.end()); std::reverse(cy.begin(), cy.end()); }
    if (std::fabs(delta) < 1.0e-20) {  // NEAR_ZERO: copies the source polygon
        for (int i = 0; i < len; ++i) out.push_back(PtI{(int)cx[i], (int)cy[i]});
        return;
    }
    const double pi = 3.141592653589793238;
    const double two_pi = pi * 2;
    double y;
    if (arc_tol <= 0.0) y = 0.25;
    else if (arc_tol > std::fabs(delta) * 0.25) y = std::fabs(delta) * 0.25;
    else y = arc_tol;
    double steps = pi / m_acos(1 - y / std::fabs(delta));
    if (steps > std::fabs(delta) * pi) steps = std::fabs(delta) * pi;
    double m_sin, m_cos;
    m_sincos(two_pi / steps, &m_sin, &m_cos);
    const double steps_per_rad = steps / two_pi;
    if (delta < 0.0) m_sin = -m_sin;
    std::vector<double> nx(len), ny(len);
    for (int j = 0; j < len; ++j) {
        const int j2 = (j + 1) % len;
        double Dx = (double)(cx[j2] - cx[j]), dy = (double)(cy[j2] - cy[j]);
        if (Dx == 0 && dy == 0) { nx[j] = 0; ny[j] = 0; continue; }
        const double f = 1 * 1.0 / std::sqrt(Dx * Dx + dy * dy);
        Dx *= f; dy *= f;
        nx[j] = dy; ny[j] = -Dx;
    }
    auto push = [&](double x, double yv) { out.push_back(PtI{(int)clipper_round(x), (int)clipper_round(yv)}); };
    int k = len - 1;
    for (int j = 0; j < len; ++j) {
        double sinA = nx[k] * ny[j] - nx[j] * ny[k];
        bool done = false;
        if (std::fabs(sinA * delta) < 1.0) {
            const double cosA = nx[k] * nx[j] + ny[j] * ny[k];
            if (cosA > 0) { push((double)cx[j] + nx[k] * delta, (double)cy[j] + ny[k] * delta); done = true; }
        } else if (sinA > 1.0) sinA = 1.0;
        else if (sinA < -1.0) sinA = -1.0;
        if (!done) {
            if (sinA * delta < 0) {
                push((double)cx[j] + nx[k] * delta, (double)cy[j] + ny[k] * delta);
                out.push_back(PtI{(int)cx[j], (int)cy[j]});
                push((double)cx[j] + nx[j] * delta, (double)cy[j] + ny[j] * delta);
            } else {  // DoRound
                const double ang = m_atan2(sinA, nx[k] * nx[j] + ny[k] * ny[j]);
                const int nsteps = std::max((int)clipper_round(steps_per_rad * std::fabs(ang)), 1);
                double Xv = nx[k], Yv = ny[k], X2;
                for (int i = 0; i < nsteps; ++i) {
                    push((double)cx[j] + Xv * delta, (double)cy[j] + Yv * delta);
                    X2 = Xv;
                    Xv = Xv * m_cos - m_sin * Yv;
                    Yv = X2 * m_sin + Yv * m_cos;
                }
                push((double)cx[j] + nx[j] * delta, (double)cy[j] + ny[j] * delta);
            }
        }
        k = j;
    }
}

// points.rs:179-194 scale_and_clip (f64, round half away, clamp to [0, ori-1])
static inline float scale_clip_1(float v, double inv, double ori) {
    double x1 = round_half_away_d((double)v * inv);
    x1 = x1 < 0.0 ? 0.0 : (x1 > ori - 1.0 ? ori - 1.0 : x1);
    return (float)x1;
}
static void scale_and_clip(float box[8], double bitmap_w, double bitmap_h, double ori_w, double ori_h) {
    const double inv_w = ori_w / bitmap_w, inv_h = ori_h / bitmap_h;
    for (int i = 0; i < 4; ++i) {
        box[2 * i] = scale_clip_1(box[2 * i], inv_w, ori_w);
        box[2 * i + 1] = scale_clip_1(box[2 * i + 1], inv_h, ori_h);
    }
}
// points.rs:125-169 side lengths: T-subtraction in f32, f64 sqrt, cast back to f32
static inline float side_len(float ax, float ay, float bx, float by) {
    const double dx = (double)(ax - bx), dy = (double)(ay - by);
    return (float)std::sqrt(dx * dx + dy * dy);
}

struct DetCfg {
    float thresh, box_thresh, unclip_ratio;
    int min_mini_box_size, dilate, quirk_x0;
};
struct DetBox { float box[8]; float score; };

struct DetTrace {  // optional per-contour stage dump for golden vectors
    std::vector<int> rect1;      // 8 ints per contour
    std::vector<float> sside1;
    std::vector<float> score;    // NaN when not evaluated
    std::vector<int> status;     // 0 kept, 1 sside<min, 2 score<thresh, 3 sside2<min+2, 4 size filter, 5 empty unclip, -1 ref panic
};

// det_processor.rs:279-335 postprocess.  ret: number of boxes, or -1 where the reference panics.
static int det_postprocess(const float* pred, int h, int w, int ori_h, int ori_w, const DetCfg& cfg,
                           std::vector<DetBox>& boxes, std::vector<uint8_t>* bitmap_out, DetTrace* trace,
                           int* comparator_inconsistent) {
    std::vector<uint8_t> mask((size_t)h * w);
    threshold_dilate(pred, h, w, cfg.thresh, cfg.dilate, mask.data());
    if (bitmap_out) *bitmap_out = mask;
    std::vector<Contour> contours;
    find_contours(mask.data(), h, w, cfg.quirk_x0, contours);
    boxes.clear();
    int panic = 0;
    for (const Contour& ct : contours) {
        double q[8];
        min_area_rect(ct.pts, q);
        const float sside = sside_of(q);
        if (trace) { for (int i = 0; i < 8; ++i) trace->rect1.push_back((int)q[i]); trace->sside1.push_back(sside); }
        auto tr = [&](float sc, int st) { if (trace) { trace->score.push_back(sc); trace->status.push_back(st); } };
        if (sside < (float)cfg.min_mini_box_size) { tr(NAN, 1); continue; }
        PtI qi[4];
        for (int i = 0; i < 4; ++i) qi[i] = PtI{(int)q[2 * i], (int)q[2 * i + 1]};
        float score = 0.0f;
        if (box_score_fast(pred, h, w, qi, &score) < 0) { panic = 1; tr(NAN, -1); continue; }
        if (score < cfg.box_thresh) { tr(score, 2); continue; }
        float bx[4], by[4];
        long long ix[4], iy[4];
        for (int i = 0; i < 4; ++i) { bx[i] = (float)qi[i].x; by[i] = (float)qi[i].y; ix[i] = qi[i].x; iy[i] = qi[i].y; }
        const float distance = unclip_distance(bx, by, cfg.unclip_ratio);
        std::vector<PtI> off;
        clipper_offset_round(ix, iy, (double)distance, 0.5, off);
        if (off.empty()) { panic = 1; tr(score, 5); continue; }  // min_area_rect(&[]) panics
        double q2[8];
        min_area_rect(off, q2);
        const float sside2 = sside_of(q2);
        if (sside2 < (float)(cfg.min_mini_box_size + 2)) { tr(score, 3); continue; }
        DetBox b;
        for (int i = 0; i < 8; ++i) b.box[i] = (float)q2[i];
        scale_and_clip(b.box, (double)w, (double)h, (double)ori_w, (double)ori_h);
        const float pb_h = side_len(b.box[0], b.box[1], b.box[6], b.box[7]);  // height_tlc: tl-bl
        const float pb_w = side_len(b.box[0], b.box[1], b.box[2], b.box[3]);  // width_tlc: tl-tr
        if (pb_h <= 3.0f || pb_w <= 3.0f) { tr(score, 4); continue; }
        b.score = score;
        boxes.push_back(b);
        tr(score, 0);
    }
    // sorted_boxes det_processor.rs:324-333: stable sort, comparator not a strict weak order.
    auto less = [](const DetBox& r1, const DetBox& r2) {
        const float c1x = (r1.box[0] + r1.box[4]) / 2.0f, c1y = (r1.box[1] + r1.box[5]) / 2.0f;
        const float c2x = (r2.box[0] + r2.box[4]) / 2.0f, c2y = (r2.box[1] + r2.box[5]) / 2.0f;
        if (fabsf(c1y - c2y) < 10.0f) return c1x < c2x;
        return c1y < c2y;
    };
    // insertion sort == Rust's stable sort for n <= 20 and for any n when the comparator is consistent
    for (size_t i = 1; i < boxes.size(); ++i) {
        DetBox v = boxes[i];
        size_t j = i;
        while (j > 0 && less(v, boxes[j - 1])) { boxes[j] = boxes[j - 1]; --j; }
        boxes[j] = v;
    }
    if (comparator_inconsistent) {
        int bad = 0;
        for (size_t i = 0; i < boxes.size() && !bad; ++i)
            for (size_t j = i + 1; j < boxes.size(); ++j)
                if (less(boxes[j], boxes[i])) { bad = 1; break; }
        *comparator_inconsistent = bad;
    }
    return panic ? -1 : (int)boxes.size();
}

// ------------------------------------------------------------------------------------------
// image_helper.rs:223-249 get_crop_img.
// RECALLED imageproc 0.25.0 geometric_transformations::{Projection::from_control_points, warp_into,
// interpolate_bicubic} and image 0.25.6 imageops::rotate270.  from_control_points solves the 8x8 DLT
// system in f64 with nalgebra's SVD; that cannot be restated bit-for-bit, but the solution is cast to
// f32 immediately, so any backward-stable f64 solver gives the same f32 coefficients except when a
// coefficient sits within ~1e-15 relative of an f32 rounding boundary.  Restated here (and in the
// CUDA path, with the same operation order) as Gaussian elimination with partial pivoting.
static bool solve8(double A[8][9]) {
    for (int col = 0; col < 8; ++col) {
        int piv = col;
        double best = std::fabs(A[col][col]);
        for (int r = col + 1; r < 8; ++r) { const double v = std::fabs(A[r][col]); if (v > best) { best = v; piv = r; } }
        if (best == 0.0) return false;
        if (piv != col) for (int c = 0; c < 9; ++c) std::swap(A[piv][c], A[col][c]);
        for (int r = col + 1; r < 8; ++r) {
            const double f = A[r][col] / A[col][col];
            if (f == 0.0) continue;
            for (int c = col; c < 9; ++c) { const double t = f * A[col][c]; A[r][c] = A[r][c] - t; }
        }
    }
    for (int r = 7; r >= 0; --r) {
        double s = A[r][8];
        for (int c = r + 1; c < 8; ++c) { const double t = A[r][c] * A[c][8]; s = s - t; }
        A[r][8] = s / A[r][r];
    }
    return true;
}

struct Proj { float t[9]; int cls; };  // t maps OUTPUT (crop) -> INPUT (page); cls 0 translation 1 affine 2 projection

static bool projection_from_box(const float box[8], float W, float H, Proj* out) {
    const double fx[4] = {box[0], box[2], box[4], box[6]}, fy[4] = {box[1], box[3], box[5], box[7]};
    const double tx[4] = {0.0, (double)W, (double)W, 0.0}, ty[4] = {0.0, 0.0, (double)H, (double)H};
    double A[8][9];
    for (int i = 0; i < 4; ++i) {
        double* r0 = A[2 * i];
        double* r1 = A[2 * i + 1];
        r0[0] = 0; r0[1] = 0; r0[2] = 0; r0[3] = -fx[i]; r0[4] = -fy[i]; r0[5] = -1.0; r0[6] = ty[i] * fx[i]; r0[7] = ty[i] * fy[i]; r0[8] = -ty[i];
        r1[0] = fx[i]; r1[1] = fy[i]; r1[2] = 1.0; r1[3] = 0; r1[4] = 0; r1[5] = 0; r1[6] = -tx[i] * fx[i]; r1[7] = -tx[i] * fy[i]; r1[8] = tx[i];
    }
    if (!solve8(A)) return false;
    float t[9];
    for (int i = 0; i < 8; ++i) t[i] = (float)A[i][8];
    t[8] = 1.0f;
    // from_matrix: normalize (t[8] == 1), classify, invert by cofactors (f32)
    int cls;
    if (fabsf(t[6]) < 1e-10f && fabsf(t[7]) < 1e-10f && fabsf(t[8] - 1.0f) < 1e-10f) {
        if (fabsf(t[0] - 1.0f) < 1e-10f && fabsf(t[1]) < 1e-10f && fabsf(t[3]) < 1e-10f && fabsf(t[4] - 1.0f) < 1e-10f) cls = 0;
        else cls = 1;
    } else cls = 2;
    const float t00 = t[0], t01 = t[1], t02 = t[2], t10 = t[3], t11 = t[4], t12 = t[5], t20 = t[6], t21 = t[7], t22 = t[8];
    const float m00 = t11 * t22 - t12 * t21;
    const float m01 = t10 * t22 - t12 * t20;
    const float m02 = t10 * t21 - t11 * t20;
    const float det = t00 * m00 - t01 * m01 + t02 * m02;
    if (fabsf(det) < 1e-10f) return false;
    const float m10 = t01 * t22 - t02 * t21;
    const float m11 = t00 * t22 - t02 * t20;
    const float m12 = t00 * t21 - t01 * t20;
    const float m20 = t01 * t12 - t02 * t11;
    const float m21 = t00 * t12 - t02 * t10;
    const float m22 = t00 * t11 - t01 * t10;
    float inv[9] = {m00 / det, -m10 / det, m20 / det, -m01 / det, m11 / det, -m21 / det, m02 / det, -m12 / det, m22 / det};
    for (int i = 0; i < 8; ++i) out->t[i] = inv[i] / inv[8];
    out->t[8] = 1.0f;
    out->cls = cls;
    return true;
}

static inline uint8_t clamp_u8_trunc(float x) { return x < 255.0f ? (x > 0.0f ? (uint8_t)x : 0) : 255; }
static inline float cubic(float p0, float p1, float p2, float p3, float x) {
    // p1 + 0.5 * x * (p2 - p0 + x * (2.0 * p0 - 5.0 * p1 + 4.0 * p2 - p3 + x * (3.0 * (p1 - p2) + p3 - p0)))
    const float a = 3.0f * (p1 - p2) + p3 - p0;
    const float b = 2.0f * p0 - 5.0f * p1 + 4.0f * p2 - p3 + x * a;
    const float c = p2 - p0 + x * b;
    return p1 + 0.5f * x * c;
}

// out dims: crop_dims(); returns 0 ok / -1 where the reference panics (degenerate projection)
static void crop_dims(const float box[8], int* cw, int* ch, int* rotated, float* Wf, float* Hf) {
    const float w_brc = side_len(box[6], box[7], box[4], box[5]);  // bl-br
    const float w_tlc = side_len(box[0], box[1], box[2], box[3]);  // tl-tr
    const float h_brc = side_len(box[2], box[3], box[4], box[5]);  // tr-br
    const float h_tlc = side_len(box[0], box[1], box[6], box[7]);  // tl-bl
    const float W = std::max(w_brc, w_tlc), H = std::max(h_brc, h_tlc);
    const uint32_t w = (uint32_t)W, h = (uint32_t)H;
    *Wf = W; *Hf = H;
    const int rot = (w > 0) && ((float)h / (float)w >= 1.5f);
    *rotated = rot;
    *cw = rot ? (int)h : (int)w;
    *ch = rot ? (int)w : (int)h;
}

static int get_crop_img(const uint8_t* img, int ih, int iw, const float box[8], uint8_t* out) {
    int cw, ch, rot; float W, H;
    crop_dims(box, &cw, &ch, &rot, &W, &H);
    const int w = rot ? ch : cw, h = rot ? cw : ch;  // un-rotated warp size
    Proj pr;
    if (!projection_from_box(box, W, H, &pr)) return -1;
    const float* t = pr.t;
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) {
            float px, py;
            const float xf = (float)x, yf = (float)y;
            if (pr.cls == 2) {
                const float d = t[6] * xf + t[7] * yf + t[8];
                px = (t[0] * xf + t[1] * yf + t[2]) / d;
                py = (t[3] * xf + t[4] * yf + t[5]) / d;
            } else if (pr.cls == 1) {
                px = t[0] * xf + t[1] * yf + t[2];
                py = t[3] * xf + t[4] * yf + t[5];
            } else { px = xf + t[2]; py = yf + t[5]; }
            uint8_t rgb[3] = {255, 255, 255};
            const float left = floorf(px) - 1.0f, right = left + 4.0f;
            const float top = floorf(py) - 1.0f, bottom = top + 4.0f;
            const float xw = px - (left + 1.0f), yw = py - (top + 1.0f);
            if (!(left < 0.0f || right >= (float)iw || top < 0.0f || bottom >= (float)ih) && std::isfinite(px) && std::isfinite(py)) {
                const uint32_t l = (uint32_t)left, tp = (uint32_t)top;
                for (int c = 0; c < 3; ++c) {
                    float col[4];
                    for (int r = 0; r < 4; ++r) {
                        const uint8_t* row = img + ((size_t)(tp + r) * iw + l) * 3 + c;
                        col[r] = (float)clamp_u8_trunc(cubic((float)row[0], (float)row[3], (float)row[6], (float)row[9], xw));
                    }
                    rgb[c] = clamp_u8_trunc(cubic(col[0], col[1], col[2], col[3], yw));
                }
            }
            size_t o;
            if (rot) o = ((size_t)(w - 1 - x) * h + y) * 3;  // rotate270: out(y, w-1-x) = in(x, y); out dims (h wide, w tall)
            else o = ((size_t)y * w + x) * 3;
            out[o] = rgb[0]; out[o + 1] = rgb[1]; out[o + 2] = rgb[2];
        }
    return 0;
}

// image_helper.rs:176-209 resize_norm_image  (img_c == 3 path)
static int resize_norm_plan(int ori_h, int ori_w, int img_h, int img_w_cfg, int has_ratio, float ratio, int* img_w, int* resized_w) {
    int iw = has_ratio ? (int)(size_t)((float)img_h * ratio) : img_w_cfg;
    const double rw = std::ceil((double)img_h * (double)ori_w / (double)ori_h);
    const size_t rwi = (size_t)rw;
    *img_w = iw;
    *resized_w = (int)std::min<size_t>((size_t)iw, rwi);
    return 0;
}
static void resize_norm_image(const uint8_t* crop, int ch, int cw, int flip180, int img_h, int img_w, int resized_w, float* out) {
    std::vector<uint8_t> src;
    if (flip180) {  // image 0.25.6 rotate180_in_place: pixel i <-> n-1-i  (image_helper.rs:268-286)
        src.resize((size_t)ch * cw * 3);
        const size_t n = (size_t)ch * cw;
        for (size_t i = 0; i < n; ++i) memcpy(&src[i * 3], crop + (n - 1 - i) * 3, 3);
        crop = src.data();
    }
    std::vector<uint8_t> rs((size_t)img_h * resized_w * 3);
    thumbnail(crop, ch, cw, 3, rs.data(), img_h, resized_w);
    const size_t plane = (size_t)img_h * img_w;
    for (size_t i = 0; i < 3 * plane; ++i) out[i] = 0.0f;
    for (int c = 0; c < 3; ++c)
        for (int y = 0; y < img_h; ++y)
            for (int x = 0; x < resized_w; ++x) {
                float v = (float)rs[((size_t)y * resized_w + x) * 3 + c] / 255.0f;
                v = v - 0.5f;
                v = v / 0.5f;
                out[c * plane + (size_t)y * img_w + x] = v;
            }
}

// RECALLED ndarray-stats 0.6.0 QuantileExt::argmax: first maximum wins (strict >), NaN -> Err -> panic
static int argmax_first(const float* v, int n, int* idx) {
    int best = 0;
    if (v[0] != v[0]) return -1;
    for (int i = 1; i < n; ++i) {
        if (v[i] != v[i]) return -1;
        if (v[i] > v[best]) best = i;
    }
    *idx = best;
    return 0;
}

}  // namespace orc

// ==========================================================================================
// C ABI for ctypes (tests) — prefix orc_
extern "C" {

void orc_set_trig_hooks(orc::trig_atan2_fn a, orc::trig_sincos_fn sc, orc::trig_acos_fn ac) {
    orc::g_hook_atan2 = a; orc::g_hook_sincos = sc; orc::g_hook_acos = ac;
}
int orc_trig_hooked() { return orc::g_hook_atan2 != nullptr; }
void orc_set_calipers_wrap(int v) { orc::g_calipers_wrap = v; }

void orc_thumbnail(const uint8_t* src, int h, int w, int c, uint8_t* dst, int nh, int nw) { orc::thumbnail(src, h, w, c, dst, nh, nw); }
int orc_resize_both_plan(int h, int w, int max_len, int min_len, int* dims) { return orc::resize_both_plan(h, w, max_len, min_len, dims); }
void orc_resize_either_plan(int h, int w, int limit_type, int limit_len, int* oh, int* ow) { orc::resize_either_plan(h, w, limit_type, limit_len, oh, ow); }
void orc_det_preprocess(const uint8_t* rgb, int h, int w, float scale, const float* mean, const float* stdv, float* out, int oh, int ow) {
    orc::det_preprocess(rgb, h, w, 0, 0, scale, mean, stdv, out, oh, ow);
}
void orc_threshold_dilate(const float* pred, int h, int w, float thr, int dilate, uint8_t* out) { orc::threshold_dilate(pred, h, w, thr, dilate, out); }

// contours flattened: pts (x,y) int32 pairs, offsets[n+1], is_hole[n]. returns n contours (or -needed if buffers too small)
int orc_find_contours(const uint8_t* mask, int h, int w, int quirk_x0, int* pts, long long max_pts, int* offsets, int* is_hole, int max_contours) {
    std::vector<orc::Contour> cs;
    orc::find_contours(mask, h, w, quirk_x0, cs);
    long long total = 0;
    for (auto& c : cs) total += (long long)c.pts.size();
    if ((int)cs.size() > max_contours || total > max_pts) return -1;
    long long o = 0;
    for (size_t i = 0; i < cs.size(); ++i) {
        offsets[i] = (int)o;
        is_hole[i] = cs[i].is_hole;
        for (auto& p : cs[i].pts) { pts[2 * o] = p.x; pts[2 * o + 1] = p.y; ++o; }
    }
    offsets[cs.size()] = (int)o;
    return (int)cs.size();
}

int orc_convex_hull(const int* pts, int n, int* out) {
    std::vector<orc::PtI> v(n);
    for (int i = 0; i < n; ++i) v[i] = orc::PtI{pts[2 * i], pts[2 * i + 1]};
    auto h = orc::convex_hull(v);
    for (size_t i = 0; i < h.size(); ++i) { out[2 * i] = h[i].x; out[2 * i + 1] = h[i].y; }
    return (int)h.size();
}
int orc_min_area_rect(const int* pts, int n, double* out8) {
    std::vector<orc::PtI> v(n);
    for (int i = 0; i < n; ++i) v[i] = orc::PtI{pts[2 * i], pts[2 * i + 1]};
    return orc::min_area_rect(v, out8);
}
int orc_box_score_fast(const float* pred, int h, int w, const int* q8, float* score) {
    orc::PtI q[4];
    for (int i = 0; i < 4; ++i) q[i] = orc::PtI{q8[2 * i], q8[2 * i + 1]};
    return orc::box_score_fast(pred, h, w, q, score);
}
int orc_polygon_mask(uint8_t* canvas, int cw, int ch, const int* q8) {
    orc::PtI q[4];
    for (int i = 0; i < 4; ++i) q[i] = orc::PtI{q8[2 * i], q8[2 * i + 1]};
    return orc::draw_polygon_mask(canvas, cw, ch, q);
}
int orc_unclip(const int* q8, float unclip_ratio, int* out_pts, int max_pts, float* distance) {
    float bx[4], by[4]; long long ix[4], iy[4];
    for (int i = 0; i < 4; ++i) { bx[i] = (float)q8[2 * i]; by[i] = (float)q8[2 * i + 1]; ix[i] = q8[2 * i]; iy[i] = q8[2 * i + 1]; }
    const float d = orc::unclip_distance(bx, by, unclip_ratio);
    if (distance) *distance = d;
    std::vector<orc::PtI> off;
    orc::clipper_offset_round(ix, iy, (double)d, 0.5, off);
    if ((int)off.size() > max_pts) return -1;
    for (size_t i = 0; i < off.size(); ++i) { out_pts[2 * i] = off[i].x; out_pts[2 * i + 1] = off[i].y; }
    return (int)off.size();
}

struct orc_det_cfg { float thresh, box_thresh, unclip_ratio; int min_mini_box_size, dilate, quirk_x0; };

// boxes_out: n x 8 f32, scores_out: n f32.  bitmap_out optional (h*w).  returns n (>=0), -1 reference-panic, -2 overflow
int orc_det_postprocess(const float* pred, int h, int w, int ori_h, int ori_w, const orc_det_cfg* cfg, float* boxes_out,
                        float* scores_out, int max_boxes, uint8_t* bitmap_out, int* comparator_inconsistent) {
    orc::DetCfg c{cfg->thresh, cfg->box_thresh, cfg->unclip_ratio, cfg->min_mini_box_size, cfg->dilate, cfg->quirk_x0};
    std::vector<orc::DetBox> boxes;
    std::vector<uint8_t> bm;
    int r = orc::det_postprocess(pred, h, w, ori_h, ori_w, c, boxes, bitmap_out ? &bm : nullptr, nullptr, comparator_inconsistent);
    if (bitmap_out) memcpy(bitmap_out, bm.data(), bm.size());
    if ((int)boxes.size() > max_boxes) return -2;
    for (size_t i = 0; i < boxes.size(); ++i) { memcpy(boxes_out + 8 * i, boxes[i].box, 32); scores_out[i] = boxes[i].score; }
    return r < 0 ? r : (int)boxes.size();
}

// per-contour trace (golden vectors): rect1 [n,8] int, sside1 [n], score [n], status [n]; returns n contours
int orc_det_trace(const float* pred, int h, int w, int ori_h, int ori_w, const orc_det_cfg* cfg, int* rect1, float* sside1,
                  float* score, int* status, int max_contours) {
    orc::DetCfg c{cfg->thresh, cfg->box_thresh, cfg->unclip_ratio, cfg->min_mini_box_size, cfg->dilate, cfg->quirk_x0};
    std::vector<orc::DetBox> boxes;
    orc::DetTrace tr;
    orc::det_postprocess(pred, h, w, ori_h, ori_w, c, boxes, nullptr, &tr, nullptr);
    int n = (int)tr.status.size();
    if (n > max_contours) return -2;
    memcpy(rect1, tr.rect1.data(), sizeof(int) * 8 * n);
    memcpy(sside1, tr.sside1.data(), sizeof(float) * n);
    memcpy(score, tr.score.data(), sizeof(float) * n);
    memcpy(status, tr.status.data(), sizeof(int) * n);
    return n;
}

void orc_scale_and_clip(float* box8, double bw, double bh, double ow, double oh) { orc::scale_and_clip(box8, bw, bh, ow, oh); }

void orc_crop_dims(const float* box8, int* cw, int* ch, int* rotated) { float W, H; orc::crop_dims(box8, cw, ch, rotated, &W, &H); }
int orc_get_crop_img(const uint8_t* img, int ih, int iw, const float* box8, uint8_t* out) { return orc::get_crop_img(img, ih, iw, box8, out); }
int orc_projection(const float* box8, float* t9, int* cls) {
    int cw, ch, rot; float W, H;
    orc::crop_dims(box8, &cw, &ch, &rot, &W, &H);
    orc::Proj p;
    if (!orc::projection_from_box(box8, W, H, &p)) return -1;
    memcpy(t9, p.t, 36); *cls = p.cls;
    return 0;
}

void orc_resize_norm_plan(int ori_h, int ori_w, int img_h, int img_w_cfg, int has_ratio, float ratio, int* img_w, int* resized_w) {
    orc::resize_norm_plan(ori_h, ori_w, img_h, img_w_cfg, has_ratio, ratio, img_w, resized_w);
}
void orc_resize_norm_image(const uint8_t* crop, int ch, int cw, int flip180, int img_h, int img_w, int resized_w, float* out) {
    orc::resize_norm_image(crop, ch, cw, flip180, img_h, img_w, resized_w, out);
}

// cls_processor.rs:108-121: per-row first-max argmax; label = labels[idx]; score = row[idx].  -1 on NaN (reference panics)
int orc_cls_postprocess(const float* logits, int n, int ncls, int* idx_out, float* score_out) {
    for (int i = 0; i < n; ++i) {
        int k;
        if (orc::argmax_first(logits + (size_t)i * ncls, ncls, &k) < 0) return -1;
        idx_out[i] = k; score_out[i] = logits[(size_t)i * ncls + k];
    }
    return 0;
}

// rec_processor.rs:190-208 + 48-97: argmax/max over classes, CTC greedy collapse (remove duplicates,
// drop index 0 = blank / ignored token 0), score = sum(p)/count sequential f32 (NaN when count == 0).
// tokens_out [n,T] (-1 padded kept ids), counts_out [n], score_out [n]; idx_out/prob_out optional [n,T].
int orc_ctc_decode(const float* logits, int n, int T, int C, int* idx_out, float* prob_out, int* tokens_out, int* counts_out, float* score_out) {
    std::vector<int> idx(T);
    std::vector<float> prob(T);
    for (int i = 0; i < n; ++i) {
        for (int t = 0; t < T; ++t) {
            int k;
            if (orc::argmax_first(logits + ((size_t)i * T + t) * C, C, &k) < 0) return -1;
            idx[t] = k; prob[t] = logits[((size_t)i * T + t) * C + k];
            if (idx_out) idx_out[(size_t)i * T + t] = k;
            if (prob_out) prob_out[(size_t)i * T + t] = prob[t];
        }
        float acc = 0.0f; unsigned cnt = 0;
        for (int t = 0; t < T; ++t) {
            bool sel = idx[t] != 0;
            if (t > 0) sel = sel && idx[t] != idx[t - 1];
            sel = sel && idx[t] != 0;  // ignored_tokens = [0]
            if (tokens_out) tokens_out[(size_t)i * T + t] = -1;
            if (sel) { if (tokens_out) tokens_out[(size_t)i * T + cnt] = idx[t]; acc = acc + prob[t]; ++cnt; }
        }
        counts_out[i] = (int)cnt;
        score_out[i] = acc / (float)cnt;
    }
    return 0;
}

}  // extern "C"
