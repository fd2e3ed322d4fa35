// db_post.cu — K2..K6: DetProcessor::postprocess (det_processor.rs:279-335) on the device, batched
// over pages:
//   A  bitmap_runs     threshold (strict >) + 2x2 dilate (det_processor.rs:284-292) -> u8 bitmap,
//                      run-start labels, per-tile foreground flags            [HBM: 4 B/px in, 5 B/px out]
//   B  ccl_merge       8-connected union-find merge of runs (atomicMin on the label image), Euler
//                      number of the bitmap (#components - #holes)            [foreground tiles only]
//   C  ccl_flatten     path compression; label = min linear index of the component == the pixel at
//                      which imageproc::contours::find_contours (det_processor.rs:293) discovers
//                      the component's outer border; collects the roots
//   D  comp_sort       roots ascending -> dense component ids in discovery order
//   E  comp_extent     per-component ymax/xmin/xmax from run ends
//   F  row_alloc       per-component row table allocation (prefix sum)
//   G  row_extreme     per-row min/max x of each component  (all hull vertices are among them)
//   H  box_geometry    one warp per component: hull -> min_area_rect -> sside filter -> box_score_fast
//                      -> unclip -> min_area_rect -> scale_and_clip -> filters   (db_geom.cuh)
//   I  page_finalize   per page: keep valid boxes in discovery order, sorted_boxes (stable insertion)
//   J  pack            prefix over pages, dense box array
// Hole borders (find_contours also returns them, retto does not filter by border_type) are detected
// through the Euler number; pages with holes are flagged (hole boxes: see DESIGN.md).
#include "common.cuh"
#include "db_geom.cuh"

#define TILE_W 128
#define TILE_H 16
#define ROWCAP 131072        // row-table entries per page
#define MAX_OFFSET_PTS 256   // points of one unclipped polygon
#define P0_ROWS 192          // rows of a component whose first hull / rect is computed out of shared memory

#define RUN_CAP 6144                       // runs per page held in shared memory by ccl_runs_kernel (88 KB: two blocks per SM)
struct RunRec { int key; int x1; };        // key = y * W + x0 (raster index of the first pixel), x1 = last pixel
// non-finite probe of the probability map (PageCounters::nonfinite): a running NaN-propagating maximum of |v| — one FMNMX3.NAN per
// two pixels (the |.| is an operand modifier) — that ends up NaN or +Inf iff some probability is NaN / +-Inf
__device__ __forceinline__ float nf_max3(float m, float a, float b) {
    float r;
    asm("max.NaN.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(m), "f"(fabsf(a)), "f"(fabsf(b)));
    return r;
}
__device__ __forceinline__ bool nf_bad(float m) { return !(m <= 3.402823466e+38f); }
struct BoxCand {
    float xy[8];
    float score;
    int valid;
    int key;     // discovery position of the contour (find_contours order)
    int status;  // parity tap: 0 kept, 1 sside<min, 2 score<box_thresh, 3 sside2<min+2, 4 size filter, 5 empty unclip,
                 //             6 never discovered by find_contours, -1 reference panic
    int rect1[8];
    float sside1;
    int rect2[8];   // second min_area_rect (after unclip), before scale_and_clip
    float dist;     // unclip distance
    int n_off;      // hull size of the unclipped polygon
    int st0;        // status after the first rect (immutable once PHASE 0 has run)
    int st2, err2;  // PHASE 3 (unclip running beside the score kernel): its status / page error, merged by geom_merge_kernel
    int big;        // PHASE 0: score list the box went to (0 huge, 1 big, 2 small)
};

__device__ __forceinline__ int find_root(const int* __restrict__ L, int a) {
    int p = L[a];
    while (p != a) { a = p; p = L[a]; }
    return a;
}
__device__ __forceinline__ void union_labels(int* L, int a, int b) {
    bool done;
    do {
        a = find_root(L, a);
        b = find_root(L, b);
        if (a < b) { const int old = atomicMin(&L[b], a); done = (old == b); b = old; }
        else if (b < a) { const int old = atomicMin(&L[a], b); done = (old == a); a = old; }
        else done = true;
    } while (!done);
}

__device__ __forceinline__ bool tile_lookup(const DetPostPage* __restrict__ pages, const int* __restrict__ tile_prefix, int n_pages,
                                            int total_tiles, int warps_per_block, DetPostPage& pg, int& page, int& s, int& rb, int& tile) {
    tile = blockIdx.x * warps_per_block + (threadIdx.x >> 5);
    if (tile >= total_tiles) return false;
    page = rt_find_segment(tile_prefix, n_pages, tile);
    pg = pages[page];
    const int lt = tile - tile_prefix[page];
    rb = lt / pg.strips;
    s = lt - rb * pg.strips;
    return true;
}

// ---- A: threshold + dilate + run labels ---------------------------------------------------------
template <bool VEC>
__global__ void __launch_bounds__(128) bitmap_runs_kernel(const DetPostPage* __restrict__ pages, const int* __restrict__ tile_prefix,
                                                           int n_pages, int total_tiles, float thr, int dilate,
                                                           unsigned char* __restrict__ bitmap, int* __restrict__ labels,
                                                           unsigned char* __restrict__ tileflags, int* __restrict__ cid_at,
                                                           int* __restrict__ key_at, PageCounters* __restrict__ counters) {
    DetPostPage pg; int page, s, rb, tile;
    if (!tile_lookup(pages, tile_prefix, n_pages, total_tiles, 4, pg, page, s, rb, tile)) return;
    const int lane = threadIdx.x & 31;
    const int W = pg.w, H = pg.h;
    const int x0 = s * TILE_W, x = x0 + lane * 4;
    const int y0 = rb * TILE_H, y1 = min(y0 + TILE_H, H);
    unsigned char* bm = bitmap + pg.px_base;
    float nf = 0.0f;   // running NaN-propagating max of |v| over the loaded probabilities (see nf_max3 / PageCounters::nonfinite)
    int* lab = labels + pg.px_base;
    int* ymax_at = cid_at + pg.px_base;   // root-indexed slots, initialised at every run start (a root is always one)
    int* keyp = key_at + pg.px_base;
    const float* __restrict__ prob = pg.prob;

    auto load_raw = [&](int y, unsigned& bits5) {  // bit0 = t(x-1), bits1..4 = t(x..x+3)
        unsigned t = 0;
        const float* row = prob + (size_t)y * W;
        if (VEC) {
            if (x < W) {
                const float4 v = __ldg(reinterpret_cast<const float4*>(row + x));
                t = (v.x > thr ? 1u : 0u) | (v.y > thr ? 2u : 0u) | (v.z > thr ? 4u : 0u) | (v.w > thr ? 8u : 0u);
                nf = nf_max3(nf_max3(nf, v.x, v.y), v.z, v.w);
            }
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (x + j < W) { const float v = __ldg(row + x + j); nf = nf_max3(nf, v, v); if (v > thr) t |= 1u << j; }
        }
        unsigned left = __shfl_up_sync(RT_FULL, (t >> 3) & 1u, 1);
        if (lane == 0) left = (x0 > 0 && __ldg(row + x0 - 1) > thr) ? 1u : 0u;
        bits5 = (t << 1) | left;
    };

    unsigned prev = 0, any = 0;
    if (dilate && y0 > 0) load_raw(y0 - 1, prev);
    for (int y = y0; y < y1; ++y) {
        unsigned cur;
        load_raw(y, cur);
        unsigned nib;
        if (dilate) {
            const unsigned m = cur | prev;          // vertical OR
            nib = ((m >> 1) | m) & 0xfu;            // bit j = m[j+1] | m[j]  (pixel j sits at bit j+1)
        } else nib = (cur >> 1) & 0xfu;
        prev = cur;
        if (VEC) { if (x >= W) nib = 0; }
        else {
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (x + j >= W) nib &= ~(1u << j);
        }
        any |= nib;
        // run starts
        const unsigned notfull = __ballot_sync(RT_FULL, nib != 0xfu);
        const unsigned below = notfull & ((1u << lane) - 1u);
        const int l = below ? 31 - __clz(below) : 0;
        const unsigned nibl = __shfl_sync(RT_FULL, nib, l);
        int carry;  // run start if the run enters this lane from the left
        if (!below) carry = x0;
        else {
            const unsigned z = (~nibl) & 0xfu;      // non-zero since lane l is not full
            carry = x0 + 4 * l + (31 - __clz(z)) + 1;
        }
        int lb[4];
        int start = carry;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (nib & (1u << j)) lb[j] = y * W + start;
            else { lb[j] = -1; start = x + j + 1; }
            if (lb[j] == y * W + x + j) { ymax_at[lb[j]] = -1; keyp[lb[j]] = 0x7fffffff; }   // run start
        }
        if (VEC) {
            if (x < W) {
                const unsigned packed = ((nib & 1u) ? 0xffu : 0u) | ((nib & 2u) ? 0xff00u : 0u) | ((nib & 4u) ? 0xff0000u : 0u) |
                                        ((nib & 8u) ? 0xff000000u : 0u);
                *reinterpret_cast<unsigned*>(bm + (size_t)y * W + x) = packed;
                *reinterpret_cast<int4*>(lab + (size_t)y * W + x) = make_int4(lb[0], lb[1], lb[2], lb[3]);
            }
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (x + j < W) { bm[(size_t)y * W + x + j] = (nib & (1u << j)) ? 255 : 0; lab[(size_t)y * W + x + j] = lb[j]; }
        }
    }
    const unsigned anyw = __ballot_sync(RT_FULL, any != 0);
    if (lane == 0) tileflags[tile] = anyw ? 1 : 0;
    if (__any_sync(RT_FULL, nf_bad(nf)) && lane == 0) atomicOr(&counters[page].nonfinite, 1);
}

// neighbourhood bits of a 4-pixel group for rows y (c) and y-1 (u): bit k = pixel x-1+k, k = 0..5
__device__ __forceinline__ void load_bits6(const unsigned char* __restrict__ bm, int W, int y, int x, int x0, int lane, bool valid_row,
                                           unsigned& bits6) {
    unsigned t = 0;
    if (valid_row && x < W) {
        if ((W & 3) == 0) {
            const unsigned v = *reinterpret_cast<const unsigned*>(bm + (size_t)y * W + x);
            t = ((v & 0xffu) ? 1u : 0u) | ((v & 0xff00u) ? 2u : 0u) | ((v & 0xff0000u) ? 4u : 0u) | ((v & 0xff000000u) ? 8u : 0u);
        } else {
            for (int j = 0; j < 4; ++j)
                if (x + j < W && bm[(size_t)y * W + x + j]) t |= 1u << j;
        }
    }
    unsigned left = __shfl_up_sync(RT_FULL, (t >> 3) & 1u, 1);
    unsigned right = __shfl_down_sync(RT_FULL, t & 1u, 1);
    if (lane == 0) left = (valid_row && x0 > 0 && bm[(size_t)y * W + x0 - 1]) ? 1u : 0u;
    if (lane == 31) right = (valid_row && x0 + TILE_W < W && bm[(size_t)y * W + x0 + TILE_W]) ? 1u : 0u;
    bits6 = left | (t << 1) | (right << 5);
}

// load_bits6 in two halves, so that the loads of row y+1 can be issued before row y is processed (the tile passes below
// are chains of dependent global accesses per row; this takes the bitmap load out of the chain)
struct RawRow { unsigned v, e; };   // v: the 4 bitmap bytes of this lane; e: neighbour byte of lane 0 / lane 31
__device__ __forceinline__ RawRow load_raw_row(const unsigned char* __restrict__ bm, int W, int y, int x, int x0, int lane, bool valid_row) {
    RawRow r; r.v = 0; r.e = 0;
    if (!valid_row) return r;
    if (x < W) {
        if ((W & 3) == 0) r.v = *reinterpret_cast<const unsigned*>(bm + (size_t)y * W + x);
        else {
            unsigned t = 0;
            for (int j = 0; j < 4; ++j)
                if (x + j < W && bm[(size_t)y * W + x + j]) t |= 0xffu << (8 * j);
            r.v = t;
        }
    }
    if (lane == 0 && x0 > 0) r.e = bm[(size_t)y * W + x0 - 1];
    if (lane == 31 && x0 + TILE_W < W) r.e = bm[(size_t)y * W + x0 + TILE_W];
    return r;
}
__device__ __forceinline__ unsigned bits6_of(const RawRow r, int lane) {
    const unsigned v = r.v;
    const unsigned t = ((v & 0xffu) ? 1u : 0u) | ((v & 0xff00u) ? 2u : 0u) | ((v & 0xff0000u) ? 4u : 0u) | ((v & 0xff000000u) ? 8u : 0u);
    unsigned left = __shfl_up_sync(RT_FULL, (t >> 3) & 1u, 1);
    unsigned right = __shfl_down_sync(RT_FULL, t & 1u, 1);
    if (lane == 0) left = r.e ? 1u : 0u;
    if (lane == 31) right = r.e ? 1u : 0u;
    return left | (t << 1) | (right << 5);
}

// ---- B: merge runs (8-connectivity) + Euler number ----------------------------------------------
__global__ void __launch_bounds__(128) ccl_merge_kernel(const DetPostPage* __restrict__ pages, const int* __restrict__ tile_prefix, int n_pages,
                                                         int total_tiles, const unsigned char* __restrict__ bitmap, int* __restrict__ labels,
                                                         const unsigned char* __restrict__ tileflags, PageCounters* __restrict__ counters) {
    DetPostPage pg; int page, s, rb, tile;
    if (!tile_lookup(pages, tile_prefix, n_pages, total_tiles, 4, pg, page, s, rb, tile)) return;
    if (!tileflags[tile]) return;
    const int lane = threadIdx.x & 31;
    const int W = pg.w, H = pg.h;
    const int x0 = s * TILE_W, x = x0 + lane * 4;
    const int y0 = rb * TILE_H, y1 = min(y0 + TILE_H, H);
    const unsigned char* bm = bitmap + pg.px_base;
    int* L = labels + pg.px_base;
    unsigned up;
    load_bits6(bm, W, y0 - 1, x, x0, lane, y0 > 0, up);
    int euler = 0;
    RawRow nxt = load_raw_row(bm, W, y0, x, x0, lane, true);
    for (int y = y0; y < y1; ++y) {
        const RawRow raw = nxt;
        nxt = load_raw_row(bm, W, y + 1, x, x0, lane, y + 1 < y1);
        const unsigned cur = bits6_of(raw, lane);
        if (cur & 0x1eu) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const unsigned cb = cur >> j, ub = up >> j;      // bit0 = W/NW, bit1 = self/N, bit2 = E/NE
                if (!(cb & 2u)) continue;
                const int p = y * W + x + j;
                const bool Wn = cb & 1u, En = cb & 4u, NW = ub & 1u, N = ub & 2u, NE = ub & 4u;
                if (N) { if (!(Wn && NW)) union_labels(L, p, p - W); }
                else {
                    if (NW && !Wn) union_labels(L, p, p - W - 1);
                    if (NE && !En) union_labels(L, p, p - W + 1);
                }
                if (j == 0 && lane == 0 && Wn) union_labels(L, p, p - 1);  // runs are labelled per strip
                // Euler characteristic of the 8-connected clique complex (V - E + F - T); every edge /
                // triangle / tetrahedron is counted at its member with the largest raster index, so only
                // foreground pixels contribute and empty tiles can be skipped.
                const int w_ = Wn ? 1 : 0, nw_ = NW ? 1 : 0, n_ = N ? 1 : 0, ne_ = NE ? 1 : 0;
                const int k3 = w_ + nw_ + n_;
                euler += 1 - (w_ + nw_ + n_ + ne_) + (k3 * (k3 - 1)) / 2 + (n_ & ne_) - (k3 == 3 ? 1 : 0);
            }
        }
        up = cur;
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) euler += __shfl_xor_sync(RT_FULL, euler, off);
    if (lane == 0 && euler) atomicAdd(&counters[page].euler, euler);
}

// ---- C: flatten + collect roots ---------------------------------------------------------------------
__global__ void __launch_bounds__(128) ccl_flatten_kernel(const DetPostPage* __restrict__ pages, const int* __restrict__ tile_prefix, int n_pages,
                                                           int total_tiles, const unsigned char* __restrict__ bitmap, int* __restrict__ labels,
                                                           const unsigned char* __restrict__ tileflags, PageCounters* __restrict__ counters,
                                                           int* __restrict__ roots, int max_comps, int* __restrict__ cid_at,
                                                           int* __restrict__ key_at) {
    DetPostPage pg; int page, s, rb, tile;
    if (!tile_lookup(pages, tile_prefix, n_pages, total_tiles, 4, pg, page, s, rb, tile)) return;
    if (!tileflags[tile]) return;
    const int lane = threadIdx.x & 31;
    const int W = pg.w, H = pg.h;
    const int x0 = s * TILE_W, x = x0 + lane * 4;
    const int y0 = rb * TILE_H, y1 = min(y0 + TILE_H, H);
    const unsigned char* bm = bitmap + pg.px_base;
    int* L = labels + pg.px_base;
    int* ymax_at = cid_at + pg.px_base;
    int* keyp = key_at + pg.px_base;
    RawRow nxt = load_raw_row(bm, W, y0, x, x0, lane, true);
    for (int y = y0; y < y1; ++y) {
        const RawRow raw = nxt;
        nxt = load_raw_row(bm, W, y + 1, x, x0, lane, y + 1 < y1);
        const unsigned cur = bits6_of(raw, lane);
        if (!(cur & 0x1eu)) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const unsigned cb = cur >> j;
            if (!(cb & 2u)) continue;
            // Only the two ends of a run are resolved to the component root here: nothing downstream reads the label of a
            // run-interior pixel (run_end_kernel looks at run ends only), and such a pixel still points at its run start,
            // whose label IS the root after this pass — labels_finalize_kernel resolves that one hop when the label plane
            // is asked for (retto_b200_det_post_fetch_labels).
            // (runs are labelled per 128-px strip: the first pixel of a strip is the representative of a run entering from the left)
            if ((cb & 5u) == 5u && !(j == 0 && lane == 0)) continue;
            const int p = y * W + x + j;
            const int l = L[p];
            const int r = find_root(L, l);
            if (r != l) L[p] = r;
            if (r == p) {
                const int slot = atomicAdd(&counters[page].n_roots, 1);
                if (slot < max_comps) roots[(size_t)page * max_comps + slot] = p;
            }
            // component extents from run ends only (root-indexed slots initialised by bitmap_runs_kernel):
            //   last row, and the position at which imageproc's scan (RECALLED, contours.rs) discovers the border: an
            //   outer border starts only at x > 0 with a zero to the left, or (as a "hole"-typed start tracing the same
            //   border) at x + 1 < width with a zero to the right; the frame itself never starts a border.
            const bool is_start = !(cb & 1u), is_end = !(cb & 4u);
            if (is_start) atomicMax(&ymax_at[r], y);
            if ((is_start && x + j > 0) || (is_end && x + j + 1 < W)) atomicMin(&keyp[r], p);
        }
    }
}

// labels of run-interior pixels: one hop through their run start (see ccl_flatten_kernel)
__global__ void labels_finalize_kernel(const unsigned char* __restrict__ bitmap, int* __restrict__ labels, long long base, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || !bitmap[base + i]) return;
    const int l = labels[base + i];
    if (l >= 0) labels[base + i] = labels[base + l];
}

// ascending bitonic sort of r[0..n) by the whole block (r must have room for the next power of two)
__device__ void block_sort_ints(int* r, int n) {
    int np2 = 1;
    while (np2 < n) np2 <<= 1;
    for (int i = n + threadIdx.x; i < np2; i += blockDim.x) r[i] = 0x7fffffff;
    __syncthreads();
    for (int k = 2; k <= np2; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < np2; i += blockDim.x) {
                const int ixj = i ^ j;
                if (ixj > i) {
                    const int a = r[i], b = r[ixj];
                    const bool up = ((i & k) == 0);
                    if ((a > b) == up) { r[i] = b; r[ixj] = a; }
                }
            }
            __syncthreads();
        }
}

// ---- D: sort roots, assign dense ids ----------------------------------------------------------------
__global__ void __launch_bounds__(1024) comp_sort_kernel(const DetPostPage* __restrict__ pages, PageCounters* __restrict__ counters,
                                                          int* __restrict__ roots, CompRec* __restrict__ comps, int* __restrict__ cid_at,
                                                          const int* __restrict__ key_at, int2* __restrict__ rowtab, int max_comps) {
    const int page = blockIdx.x;
    const DetPostPage pg = pages[page];
    int n = counters[page].n_roots;
    if (n > max_comps) {
        if (threadIdx.x == 0) { counters[page].status = RETTO_B200_ERR_CAPACITY; counters[page].n_roots = 0; }
        return;
    }
    int* r = roots + (size_t)page * max_comps;
    block_sort_ints(r, n);
    CompRec* c = comps + (size_t)page * max_comps;
    int* cid = cid_at + pg.px_base;
    const int* keyp = key_at + pg.px_base;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const int root = r[i];
        CompRec cr;
        cr.root = root; cr.ymax = cid[root]; cr.xmin = 0; cr.xmax = 0; cr.row_off = 0;
        cr.key = keyp[root]; cr.ymin = root / pg.w; cr.pad = 0;
        c[i] = cr;
        cid[root] = i;   // the slot now holds the dense component id (discovery order of the first pixel)
    }
    __syncthreads();
    // row-table allocation (one row per component row), exclusive prefix over the components
    __shared__ int s_part[1024];
    __shared__ int s_total;
    const int per = (n + blockDim.x - 1) / blockDim.x;
    const int b = threadIdx.x * per, e = min(n, b + per);
    int sum = 0;
    for (int i = b; i < e; ++i) sum += c[i].ymax - c[i].ymin + 1;
    s_part[threadIdx.x] = sum;
    __syncthreads();
    if (threadIdx.x == 0) {
        int run = 0;
        for (int i = 0; i < (int)blockDim.x; ++i) { const int v = s_part[i]; s_part[i] = run; run += v; }
        s_total = run;
        counters[page].row_total = run;
        if (run > ROWCAP) counters[page].status = RETTO_B200_ERR_CAPACITY;
    }
    __syncthreads();
    int off = s_part[threadIdx.x];
    for (int i = b; i < e; ++i) { c[i].row_off = off; off += c[i].ymax - c[i].ymin + 1; }
    const int total = min(s_total, ROWCAP);
    int2* rt = rowtab + (size_t)page * ROWCAP;
    for (int i = threadIdx.x; i < total; i += blockDim.x) rt[i] = make_int2(0x7fffffff, -1);
}

// ---- E / G: run-end passes ------------------------------------------------------------------------------
template <int PASS>  // 0: component extents, 1: per-row extremes
__global__ void __launch_bounds__(128) run_end_kernel(const DetPostPage* __restrict__ pages, const int* __restrict__ tile_prefix, int n_pages,
                                                       int total_tiles, const unsigned char* __restrict__ bitmap, const int* __restrict__ labels,
                                                       const unsigned char* __restrict__ tileflags, const int* __restrict__ cid_at,
                                                       CompRec* __restrict__ comps, int2* __restrict__ rowtab,
                                                       const PageCounters* __restrict__ counters, int max_comps) {
    DetPostPage pg; int page, s, rb, tile;
    if (!tile_lookup(pages, tile_prefix, n_pages, total_tiles, 4, pg, page, s, rb, tile)) return;
    if (!tileflags[tile]) return;
    if (counters[page].status != RETTO_B200_OK) return;
    const int lane = threadIdx.x & 31;
    const int W = pg.w, H = pg.h;
    const int x0 = s * TILE_W, x = x0 + lane * 4;
    const int y0 = rb * TILE_H, y1 = min(y0 + TILE_H, H);
    const unsigned char* bm = bitmap + pg.px_base;
    const int* L = labels + pg.px_base;
    const int* cid = cid_at + pg.px_base;
    CompRec* c = comps + (size_t)page * max_comps;
    int2* rt = rowtab + (size_t)page * ROWCAP;
    RawRow nxt = load_raw_row(bm, W, y0, x, x0, lane, true);
    for (int y = y0; y < y1; ++y) {
        const RawRow raw = nxt;
        nxt = load_raw_row(bm, W, y + 1, x, x0, lane, y + 1 < y1);
        const unsigned cur = bits6_of(raw, lane);
        if (!(cur & 0x1eu)) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const unsigned cb = cur >> j;
            if (!(cb & 2u)) continue;
            const bool is_start = !(cb & 1u), is_end = !(cb & 4u);
            if (!is_start && !is_end) continue;
            const int p = y * W + x + j;
            const int id = cid[L[p]];
            if (PASS == 0) {
                if (is_start) { atomicMin(&c[id].xmin, x + j); atomicMax(&c[id].ymax, y); }
                if (is_end) atomicMax(&c[id].xmax, x + j);
                // imageproc's scan (RECALLED, contours.rs) starts an outer border only at x > 0 with a zero to
                // the left, or (as a "hole"-typed start tracing the same border) at x + 1 < width with a zero to
                // the right; the frame itself never starts a border.
                if ((is_start && x + j > 0) || (is_end && x + j + 1 < W)) atomicMin(&c[id].key, p);
            } else {
                const CompRec cr = c[id];
                const int ridx = cr.row_off + (y - cr.ymin);
                if (ridx < ROWCAP) {
                    if (is_start) atomicMin(&rt[ridx].x, x + j);
                    if (is_end) atomicMax(&rt[ridx].y, x + j);
                }
            }
        }
    }
}

// ---- F: row table allocation -------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) row_alloc_kernel(const DetPostPage* __restrict__ pages, PageCounters* __restrict__ counters,
                                                          CompRec* __restrict__ comps, int2* __restrict__ rowtab, int max_comps) {
    const int page = blockIdx.x;
    const DetPostPage pg = pages[page];
    if (counters[page].status != RETTO_B200_OK) return;
    const int n = counters[page].n_roots;
    CompRec* c = comps + (size_t)page * max_comps;
    __shared__ int s_part[1024];
    __shared__ int s_total;
    // each thread owns a contiguous chunk of components
    const int per = (n + blockDim.x - 1) / blockDim.x;
    const int b = threadIdx.x * per, e = min(n, b + per);
    int sum = 0;
    for (int i = b; i < e; ++i) sum += c[i].ymax - c[i].ymin + 1;
    s_part[threadIdx.x] = sum;
    __syncthreads();
    if (threadIdx.x == 0) {
        int run = 0;
        for (int i = 0; i < (int)blockDim.x; ++i) { const int v = s_part[i]; s_part[i] = run; run += v; }
        s_total = run;
        counters[page].row_total = run;
        if (run > ROWCAP) counters[page].status = RETTO_B200_ERR_CAPACITY;
    }
    __syncthreads();
    int off = s_part[threadIdx.x];
    for (int i = b; i < e; ++i) { c[i].row_off = off; off += c[i].ymax - c[i].ymin + 1; }
    const int total = min(s_total, ROWCAP);
    int2* rt = rowtab + (size_t)page * ROWCAP;
    for (int i = threadIdx.x; i < total; i += blockDim.x) rt[i] = make_int2(0x7fffffff, -1);
}

// ---- H: per-component geometry ---------------------------------------------------------------------------------
struct GeomParams {
    float box_thresh, unclip_ratio;
    int min_mini_box_size;
};

// Three phases, one warp per contour each (same grid mapping), communicating through BoxCand:
//   PHASE 0  hull -> first min_area_rect -> sside filter                     (f64 trig: register heavy)
//   PHASE 1  box_score_fast                                                  (long sequential chain: light kernel, so
//                                                                             every box of an SM is resident at once)
//   PHASE 2  unclip -> second min_area_rect -> scale_and_clip -> filters     (f64 trig: register heavy)
//   PHASE 3  = PHASE 2 for every box that reached the score stage, launched on a second stream BESIDE the score kernel
//            (the score kernel is a set of long dependent add chains — its tail leaves the SMs nearly empty, and the
//            unclip of a box does not need its score, only the decision).  It writes st2 / err2 instead of status /
//            valid / the page status; geom_merge_kernel applies them to the boxes whose score passed.
#define ST_NEED_SCORE 7
#define SCORE_HUGE_AREA 49152    // bounding-box pixels above which a box is scored ahead of the others (one warp per box)
#define SCORE_BIG_AREA 16384     // ... above which a box comes before the small ones
#define SCORE_HUGE_BLOCKS 74     // leading blocks of box_score_wpb_kernel that work through the huge list
#define ST_NEED_UNCLIP 8
template <int PHASE>
__global__ void __launch_bounds__(128, PHASE == 1 ? 10 : 4) box_geometry_kernel(const DetPostPage* __restrict__ pages,
                                                            int n_pages, PageCounters* __restrict__ counters, const CompRec* __restrict__ comps,
                                                            const int2* __restrict__ rowtab, int2* __restrict__ hullbuf, BoxCand* __restrict__ cand,
                                                            int max_comps, GeomParams gp, const int* __restrict__ hole_pages, int* __restrict__ biglist = nullptr, int list_cap = 0) {
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    // hole_pages == nullptr: ids [0, n_roots) of every page (outer borders); else ids [n_roots, n_roots + n_holes)
    // of the listed pages (hole borders)
    const int page = hole_pages ? hole_pages[blockIdx.y] : blockIdx.y;
    if (counters[page].status == RETTO_B200_ERR_CAPACITY) return;
    const int id = (hole_pages ? counters[page].n_roots : 0) + blockIdx.x * 4 + wib;
    const int n_comp = counters[page].n_roots + (hole_pages ? counters[page].n_holes : 0);
    if (id >= n_comp) return;
    const DetPostPage pg = pages[page];
    BoxCand* out = cand + (size_t)page * max_comps + id;

    if (PHASE == 0) {
        __shared__ int s_n0[4];
        __shared__ int2 s_rows[4][P0_ROWS];        // the component's row extremes, staged by the whole warp (coalesced)
        __shared__ int2 s_h0[4][2 * P0_ROWS];      // hull stack of the serial monotone chains + input of the calipers
        const CompRec cr = comps[(size_t)page * max_comps + id];
        const int ymin = cr.ymin;
        const int R = cr.ymax - ymin + 1;
        const int2* rt = rowtab + (size_t)page * ROWCAP + cr.row_off;
        int2* hull = hullbuf + ((size_t)page * ROWCAP + cr.row_off) * 2;
        // hull of the component == hull of its border.  The two monotone chains are serial (lane 0) and pop / push a stack at every
        // row: with the rows and the stack in global memory every step was a dependent global access (0.55 ms on config 2's rotated
        // rectangles); components of up to P0_ROWS rows now run entirely out of shared memory, taller ones keep the global path.
        const bool in_smem = R <= P0_ROWS;
        if (in_smem) {
            for (int i = lane; i < R; i += 32) s_rows[wib][i] = rt[i];
            __syncwarp();
            hull = s_h0[wib];
        }
        if (lane == 0) {
            if (in_smem) {
                const int2* sr = s_rows[wib];
                s_n0[wib] = hull_from_rows(R, [&](int i, int& y, int& a, int& b) { const int2 v = sr[i]; y = ymin + i; a = v.x; b = v.y; }, hull);
            } else
                s_n0[wib] = hull_from_rows(R, [&](int i, int& y, int& a, int& b) { const int2 v = rt[i]; y = ymin + i; a = v.x; b = v.y; }, hull);
        }
        __syncwarp();
        const int nh = s_n0[wib];
        double q[8];
        warp_min_area_rect(hull, nh, q);
        const float sside = sside_of(q);
        if (lane == 0) {
            out->valid = 0; out->key = cr.key; out->sside1 = sside; out->score = CUDART_NAN_F;
#pragma unroll
            for (int i = 0; i < 8; ++i) out->rect1[i] = (int)q[i];
            int st = ST_NEED_SCORE;
            if (cr.key == 0x7fffffff) st = 6;                          // every run spans the full width: never discovered
            else if (sside < (float)gp.min_mini_box_size) st = 1;
            out->status = st; out->st0 = st; out->st2 = -2; out->err2 = 0;
            int big = 0;
            if (biglist && st == ST_NEED_SCORE) {
                // class by bounding-box area
                int xa = (int)q[0], xb = xa, ya = (int)q[1], yb = ya;
#pragma unroll
                for (int i = 1; i < 4; ++i) { xa = min(xa, (int)q[2 * i]); xb = max(xb, (int)q[2 * i]); ya = min(ya, (int)q[2 * i + 1]); yb = max(yb, (int)q[2 * i + 1]); }
                xa = min(max(xa, 0), pg.w - 1); xb = min(max(xb, 0), pg.w - 1); ya = min(max(ya, 0), pg.h - 1); yb = min(max(yb, 0), pg.h - 1);
                const long long area = (long long)(xb - xa + 1) * (long long)(yb - ya + 1);
                big = area > SCORE_HUGE_AREA ? 0 : area > SCORE_BIG_AREA ? 1 : 2;
                const int slot = atomicAdd(&biglist[big], 1);
                if (slot < list_cap) reinterpret_cast<int2*>(biglist + 4)[(size_t)big * list_cap + slot] = make_int2(page, id);
            }
            out->big = big;
        }
        return;
    }
    if (PHASE == 1) {
        if (out->status != ST_NEED_SCORE) return;
        int qx[4], qy[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) { qx[i] = out->rect1[2 * i]; qy[i] = out->rect1[2 * i + 1]; }
        __syncwarp();
        __shared__ __align__(16) float s_buf[4][128];
        __shared__ int s_cov[4][32 * 17];
        float score;
        if (!warp_box_score(pg.prob, pg.h, pg.w, qx, qy, &score, s_buf[wib], s_cov[wib])) {
            if (lane == 0) { atomicMax(&counters[page].status, RETTO_B200_ERR_DEGENERATE_QUAD); out->status = -1; }
            return;
        }
        if (lane == 0) { out->score = score; out->status = (score < gp.box_thresh) ? 2 : ST_NEED_UNCLIP; }
        return;
    }
    if (PHASE == 4) {
        // Pages whose map holds NaN / +-Inf (PageCounters::nonfinite, set by the bitmap kernel; the host launches this phase
        // only when its early counter read-back shows such a page): the reference folds v * m over the WHOLE bounding box
        // (det_processor.rs:212-219), so a non-finite value outside the polygon (m = 0) still poisons the sum — redo the
        // scored boxes of these pages with the exact classification.  NaN is not < box_thresh: the box is kept, as in the reference.
        if (!counters[page].nonfinite || out->st0 != ST_NEED_SCORE || out->status == -1) return;
        int qx[4], qy[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) { qx[i] = out->rect1[2 * i]; qy[i] = out->rect1[2 * i + 1]; }
        const float prev = out->score;
        __syncwarp();
        const float score = warp_box_score_nonfinite(pg.prob, pg.h, pg.w, qx, qy, prev);
        if (lane == 0) { out->score = score; out->status = (score < gp.box_thresh) ? 2 : ST_NEED_UNCLIP; }
        return;
    }
    if (PHASE == 2 || PHASE == 3) {
        __shared__ int2 s_pts[4][MAX_OFFSET_PTS];
        __shared__ int2 s_hull[4][2 * MAX_OFFSET_PTS];
        __shared__ int s_n[4];
        if (PHASE == 2 ? out->status != ST_NEED_UNCLIP : out->st0 != ST_NEED_SCORE) return;
        int qx[4], qy[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) { qx[i] = out->rect1[2 * i]; qy[i] = out->rect1[2 * i + 1]; }
        __syncwarp();
        {
            // unclip by the whole warp: the Clipper offset with one corner per lane, a rank sort of the offset points by (y, x), then
            // lane 0 compresses the sorted points to per-row extremes and runs the two monotone chains (was: everything on lane 0)
            const float dist = unclip_distance(qx, qy, gp.unclip_ratio);
            if (lane == 0) out->dist = dist;
            int2* p = s_pts[wib];
            int m = warp_clipper_offset_round(qx, qy, (double)dist, 0.5, p, MAX_OFFSET_PTS);
            __syncwarp();
            if (m > 0) {
                int2* sorted = s_hull[wib];   // scratch: sorted copy, then copied back
                for (int i = lane; i < m; i += 32) {
                    const int2 v = p[i];
                    int rank = 0;
                    for (int j = 0; j < m; ++j) {
                        const int2 u = p[j];
                        rank += (u.y < v.y || (u.y == v.y && (u.x < v.x || (u.x == v.x && j < i)))) ? 1 : 0;
                    }
                    sorted[rank] = v;
                }
                __syncwarp();
                for (int i = lane; i < m; i += 32) p[i] = sorted[i];
                __syncwarp();
            }
            if (lane == 0) {
                if (m > 0) {
                    // compress to rows in place: (y, xmin, xmax) stored as pairs in s_hull's upper half
                    int2* rows_y = s_hull[wib] + MAX_OFFSET_PTS;            // .x = y, .y unused
                    int2* rows_x = s_hull[wib] + MAX_OFFSET_PTS + MAX_OFFSET_PTS / 2;  // .x = xmin, .y = xmax
                    int nr = 0;
                    for (int i = 0; i < m;) {
                        int j = i;
                        while (j + 1 < m && p[j + 1].y == p[i].y) ++j;
                        if (nr < MAX_OFFSET_PTS / 2) { rows_y[nr] = make_int2(p[i].y, 0); rows_x[nr] = make_int2(p[i].x, p[j].x); ++nr; }
                        i = j + 1;
                    }
                    m = hull_from_rows(nr, [&](int i, int& y, int& a, int& b) { y = rows_y[i].x; a = rows_x[i].x; b = rows_x[i].y; }, s_hull[wib]);
                }
                s_n[wib] = m;
            }
        }
        __syncwarp();
        const int nh = s_n[wib];
        if (nh <= 0) {
            // Clipper returned nothing (or overflow): the reference would panic in min_area_rect(&[])
            if (lane == 0) {
                const int code = nh < 0 ? RETTO_B200_ERR_CAPACITY : RETTO_B200_ERR_DEGENERATE_QUAD;
                if (PHASE == 2) { atomicMax(&counters[page].status, code); out->status = 5; }
                else { out->err2 = code; out->st2 = 5; }
            }
            return;
        }
        double q2[8];
        warp_min_area_rect(s_hull[wib], nh, q2);
        const float sside2 = sside_of(q2);
        if (lane == 0) { out->n_off = nh; for (int i = 0; i < 8; ++i) out->rect2[i] = (int)q2[i]; }
        if (sside2 < (float)(gp.min_mini_box_size + 2)) { if (lane == 0) { if (PHASE == 2) out->status = 3; else out->st2 = 3; } return; }
        if (lane == 0) {
            if (PHASE == 2) out->status = 4; else out->st2 = 4;
            const double inv_w = __ddiv_rn((double)pg.ori_w, (double)pg.w), inv_h = __ddiv_rn((double)pg.ori_h, (double)pg.h);
            float b[8];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                b[2 * i] = scale_clip_1((float)q2[2 * i], inv_w, (double)pg.ori_w);
                b[2 * i + 1] = scale_clip_1((float)q2[2 * i + 1], inv_h, (double)pg.ori_h);
            }
            const float pb_h = side_len(b[0], b[1], b[6], b[7]);
            const float pb_w = side_len(b[0], b[1], b[2], b[3]);
            if (!(pb_h <= 3.0f || pb_w <= 3.0f)) {
#pragma unroll
                for (int i = 0; i < 8; ++i) out->xy[i] = b[i];
                if (PHASE == 2) { out->valid = 1; out->status = 0; }   // out->score is already the box score (PHASE 1)
                else out->st2 = 0;
            }
        }
    }
}

// box_score_fast for the outer-border boxes of a batch (PHASE 1 of box_geometry_kernel, re-cut): one warp per box
// (warp_box_score), but in an order that suits the add chains instead of component order.  PHASE 0 sorts the boxes that need a
// score into three lists by bounding-box area (huge / big / small); the first SCORE_HUGE_BLOCKS blocks work through the huge
// list — block order is dispatch order, so the longest chains of the batch (they ARE the kernel's duration when they start late)
// start first — and the other blocks take the big list, then the small one, grid-stride (the counts live on the device).
// Measured and dropped this round (all bit-exact, `profiles/r02_score_variants.md`): eight boxes per warp with the pieces staged
// through shared memory (half the instructions, but 16 warps per SM and a serial piece generator: no faster), one box per lane
// (a tenth of the instructions, but every 8-pixel step waits for an L2 round trip: 3x slower).
__device__ __forceinline__ const int2* score_list(const int* lists, int cap, int which) { return reinterpret_cast<const int2*>(lists + 4) + (size_t)which * cap; }

__global__ void __launch_bounds__(128, 10) box_score_kernel(const DetPostPage* __restrict__ pages, PageCounters* __restrict__ counters,
                                                            BoxCand* __restrict__ cand, int max_comps, GeomParams gp,
                                                            const int* __restrict__ lists, int cap) {
    __shared__ __align__(16) float s_buf[4][128];
    __shared__ int s_cov[4][32 * 17];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const bool huge = blockIdx.x < SCORE_HUGE_BLOCKS;
    const int n_huge = min(lists[0], cap), n_big = min(lists[1], cap), n_small = min(lists[2], cap);
    const int cnt = huge ? n_huge : n_big + n_small;
    const int first = (huge ? blockIdx.x : blockIdx.x - SCORE_HUGE_BLOCKS) * 4 + wib;
    const int stride = (huge ? SCORE_HUGE_BLOCKS : (int)gridDim.x - SCORE_HUGE_BLOCKS) * 4;
    for (int i = first; i < cnt; i += stride) {
        const int2 e = huge ? score_list(lists, cap, 0)[i] : i < n_big ? score_list(lists, cap, 1)[i] : score_list(lists, cap, 2)[i - n_big];
        const DetPostPage pg = pages[e.x];
        BoxCand* out = cand + (size_t)e.x * max_comps + e.y;
        int qx[4], qy[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) { qx[k] = out->rect1[2 * k]; qy[k] = out->rect1[2 * k + 1]; }
        float score;
        const bool ok = warp_box_score(pg.prob, pg.h, pg.w, qx, qy, &score, s_buf[wib], s_cov[wib]);
        if (lane == 0) {
            if (!ok) { atomicMax(&counters[e.x].status, RETTO_B200_ERR_DEGENERATE_QUAD); out->status = -1; }
            else { out->score = score; out->status = (score < gp.box_thresh) ? 2 : ST_NEED_UNCLIP; }
        }
        __syncwarp();
    }
}

// after the join of box_geometry_kernel<1> (main stream) and <3> (auxiliary stream): boxes whose score passed take the
// outcome of their unclip, everything else keeps the status the score stage gave it
__global__ void __launch_bounds__(128) geom_merge_kernel(int n_pages, PageCounters* __restrict__ counters, BoxCand* __restrict__ cand, int max_comps) {
    const int page = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (page >= n_pages || counters[page].status == RETTO_B200_ERR_CAPACITY || i >= counters[page].n_roots || i >= max_comps) return;
    BoxCand& b = cand[(size_t)page * max_comps + i];
    if (b.st0 != ST_NEED_SCORE || b.status != ST_NEED_UNCLIP) return;
    const int st2 = b.st2;
    b.status = st2;
    b.valid = st2 == 0;
    if (st2 == 5) atomicMax(&counters[page].status, b.err2);
}

// ---- hole borders ---------------------------------------------------------------------------------------------
// find_contours (det_processor.rs:293) also returns hole borders and retto does not filter by border_type, so every
// background region enclosed by text pixels yields one more contour: the foreground pixels 4-adjacent to it.
// Pages whose Euler number says holes exist (rare after the 2x2 dilation) take this extra path: 4-connected CCL of
// the background (labels reuse the cid_at array, which is free by now), regions not connected to the frame are
// holes; each becomes one more CompRec (discovery key = the pixel left of the hole's first pixel) with per-row
// extremes of its border pixels, and goes through the same box_geometry kernel.
#define MAX_HOLES 4096
struct HoleArgs {
    const DetPostPage* pages; const int* flag_pages; const int* flag_prefix; int nf; int total;
};
__device__ __forceinline__ bool hole_px(const HoleArgs& h, int& page, DetPostPage& pg, int& p) {
    const int u = blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= h.total) return false;
    const int k = rt_find_segment(h.flag_prefix, h.nf, u);
    page = h.flag_pages[k];
    pg = h.pages[page];
    p = u - h.flag_prefix[k];
    return true;
}
__global__ void bg_init_kernel(HoleArgs h, const unsigned char* __restrict__ bitmap, int* __restrict__ bgl, int* __restrict__ labels_init) {
    int page, p; DetPostPage pg;
    if (!hole_px(h, page, pg, p)) return;
    bgl[pg.px_base + p] = bitmap[pg.px_base + p] ? -1 : p;
    // run-table path: the label plane has not been written; the hole kernels use its background entries as scratch
    if (labels_init) labels_init[pg.px_base + p] = -1;
}
__global__ void bg_merge_kernel(HoleArgs h, const unsigned char* __restrict__ bitmap, int* __restrict__ bgl) {
    int page, p; DetPostPage pg;
    if (!hole_px(h, page, pg, p)) return;
    const unsigned char* bm = bitmap + pg.px_base;
    if (bm[p]) return;
    int* L = bgl + pg.px_base;
    const int y = p / pg.w, x = p - y * pg.w;
    if (x > 0 && !bm[p - 1]) union_labels(L, p, p - 1);
    if (y > 0 && !bm[p - pg.w]) union_labels(L, p, p - pg.w);
}
__global__ void bg_flatten_kernel(HoleArgs h, const unsigned char* __restrict__ bitmap, int* __restrict__ bgl, int* __restrict__ labels) {
    int page, p; DetPostPage pg;
    if (!hole_px(h, page, pg, p)) return;
    if (bitmap[pg.px_base + p]) return;
    int* L = bgl + pg.px_base;
    const int r = find_root(L, p);
    L[p] = r;
    const int y = p / pg.w, x = p - y * pg.w;
    if (x == 0 || y == 0 || x == pg.w - 1 || y == pg.h - 1) labels[pg.px_base + r] = -2;  // connected to the frame: outer background
}
__global__ void hole_collect_kernel(HoleArgs h, const unsigned char* __restrict__ bitmap, const int* __restrict__ bgl,
                                    const int* __restrict__ labels, PageCounters* __restrict__ counters, int* __restrict__ holes) {
    int page, p; DetPostPage pg;
    if (!hole_px(h, page, pg, p)) return;
    if (bitmap[pg.px_base + p] || bgl[pg.px_base + p] != p || labels[pg.px_base + p] == -2) return;
    const int slot = atomicAdd(&counters[page].n_holes, 1);
    if (slot < MAX_HOLES) holes[(size_t)page * MAX_HOLES + slot] = p;
}
__global__ void __launch_bounds__(1024) hole_sort_kernel(HoleArgs h, PageCounters* __restrict__ counters, int* __restrict__ holes,
                                                          int* __restrict__ labels, CompRec* __restrict__ comps, int max_comps) {
    const int page = h.flag_pages[blockIdx.x];
    const DetPostPage pg = h.pages[page];
    const int n = counters[page].n_holes, nr = counters[page].n_roots;
    if (n > MAX_HOLES || nr + n > max_comps) {
        if (threadIdx.x == 0) { counters[page].status = RETTO_B200_ERR_CAPACITY; counters[page].n_holes = 0; }
        return;
    }
    int* r = holes + (size_t)page * MAX_HOLES;
    block_sort_ints(r, n);
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const int root = r[i];
        labels[pg.px_base + root] = -(3 + i);       // root -> hole id (restored to -1 by hole_restore_kernel)
        CompRec cr;
        cr.root = root; cr.ymax = -1; cr.xmin = 0; cr.xmax = 0; cr.row_off = 0;
        cr.key = root - 1;                           // discovered at the foreground pixel left of the hole's first pixel
        cr.ymin = root / pg.w; cr.pad = 0;
        comps[(size_t)page * max_comps + nr + i] = cr;
    }
}
__global__ void hole_extent_kernel(HoleArgs h, const unsigned char* __restrict__ bitmap, const int* __restrict__ bgl,
                                   const int* __restrict__ labels, const PageCounters* __restrict__ counters, CompRec* __restrict__ comps,
                                   int max_comps) {
    int page, p; DetPostPage pg;
    if (!hole_px(h, page, pg, p)) return;
    if (bitmap[pg.px_base + p] || counters[page].status != RETTO_B200_OK) return;
    const int v = labels[pg.px_base + bgl[pg.px_base + p]];
    if (v > -3) return;
    atomicMax(&comps[(size_t)page * max_comps + counters[page].n_roots + (-v - 3)].ymax, p / pg.w);
}
__global__ void __launch_bounds__(256) hole_row_alloc_kernel(HoleArgs h, PageCounters* __restrict__ counters, CompRec* __restrict__ comps,
                                                              int2* __restrict__ rowtab, int max_comps) {
    const int page = h.flag_pages[blockIdx.x];
    if (counters[page].status != RETTO_B200_OK) return;
    __shared__ int s_first, s_total;
    const int n = counters[page].n_holes, nr = counters[page].n_roots;
    CompRec* c = comps + (size_t)page * max_comps + nr;
    if (threadIdx.x == 0) {
        int off = counters[page].row_total;
        s_first = off;
        for (int i = 0; i < n; ++i) {   // border rows span [ymin - 1, ymax + 1]
            c[i].ymin -= 1; c[i].ymax += 1;
            c[i].row_off = off;
            off += c[i].ymax - c[i].ymin + 1;
        }
        s_total = off;
        counters[page].row_total = off;
        if (off > ROWCAP) counters[page].status = RETTO_B200_ERR_CAPACITY;
    }
    __syncthreads();
    int2* rt = rowtab + (size_t)page * ROWCAP;
    for (int i = s_first + threadIdx.x; i < min(s_total, ROWCAP); i += blockDim.x) rt[i] = make_int2(0x7fffffff, -1);
}
__global__ void hole_rows_kernel(HoleArgs h, const unsigned char* __restrict__ bitmap, const int* __restrict__ bgl,
                                 const int* __restrict__ labels, const PageCounters* __restrict__ counters, const CompRec* __restrict__ comps,
                                 int2* __restrict__ rowtab, int max_comps) {
    int page, p; DetPostPage pg;
    if (!hole_px(h, page, pg, p)) return;
    const unsigned char* bm = bitmap + pg.px_base;
    if (bm[p] || counters[page].status != RETTO_B200_OK) return;
    const int v = labels[pg.px_base + bgl[pg.px_base + p]];
    if (v > -3) return;
    const CompRec cr = comps[(size_t)page * max_comps + counters[page].n_roots + (-v - 3)];
    int2* rt = rowtab + (size_t)page * ROWCAP + cr.row_off;
    const int y = p / pg.w, x = p - y * pg.w;   // a hole never touches the frame: all 4 neighbours are inside the page
    if (bm[p - 1]) { atomicMin(&rt[y - cr.ymin].x, x - 1); atomicMax(&rt[y - cr.ymin].y, x - 1); }
    if (bm[p + 1]) { atomicMin(&rt[y - cr.ymin].x, x + 1); atomicMax(&rt[y - cr.ymin].y, x + 1); }
    if (bm[p - pg.w]) { atomicMin(&rt[y - 1 - cr.ymin].x, x); atomicMax(&rt[y - 1 - cr.ymin].y, x); }
    if (bm[p + pg.w]) { atomicMin(&rt[y + 1 - cr.ymin].x, x); atomicMax(&rt[y + 1 - cr.ymin].y, x); }
}
// Outer borders on pages with holes.  imageproc's scan starts an outer border at a pixel with a zero to its LEFT (x > 0), or — typed
// "hole" but still the outer border — at a pixel with a zero to its RIGHT (x + 1 < width), provided the pixel has not been marked by an
// earlier trace.  A run start / end that faces a HOLE of its own component never qualifies: the hole's border was traced (and its pixels
// marked) from the pixel left of the hole's first pixel, which precedes every other pixel around the hole in raster order.  So the
// discovery key of a component is the minimum over run starts / ends that face background the component does not enclose (frame-
// connected background, or the hole of ANOTHER component in which this one is an island); a component that
// touches x = 0 in its first row and otherwise only faces its own holes (a page-filling background with letters as holes) is never
// discovered at all.  ccl_runs / ccl_flatten computed the key over all starts / ends (exact when the page has no holes); these three
// kernels recompute it for the pages the hole path runs on, using the background labelling that path has just built.
__device__ __forceinline__ int root_of_fg_pixel(bool run_path, const RunRec* __restrict__ psorted, const int* __restrict__ plabel, int n_runs,
                                                const int* __restrict__ lab, int p) {
    if (!run_path) return lab[p];      // run starts / ends are resolved to the root by ccl_flatten
    int lo = 0, hi = n_runs;           // last run with key <= p
    while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (psorted[mid].key <= p) lo = mid; else hi = mid; }
    return plabel[lo];
}
__global__ void outer_key_init_kernel(HoleArgs h, const PageCounters* __restrict__ counters, const CompRec* __restrict__ comps, int max_comps,
                                      int* __restrict__ key_at) {
    const int page = h.flag_pages[blockIdx.y];
    const int id = blockIdx.x * blockDim.x + threadIdx.x;
    if (counters[page].status == RETTO_B200_ERR_CAPACITY || id >= counters[page].n_roots) return;
    key_at[h.pages[page].px_base + comps[(size_t)page * max_comps + id].root] = 0x7fffffff;
}
__global__ void outer_key_scan_kernel(HoleArgs h, const unsigned char* __restrict__ bitmap, const int* __restrict__ bgl, const int* __restrict__ labels,
                                      const PageCounters* __restrict__ counters, int run_path, const RunRec* __restrict__ runs,
                                      int* __restrict__ key_at) {
    int page, p; DetPostPage pg;
    if (!hole_px(h, page, pg, p)) return;
    const unsigned char* bm = bitmap + pg.px_base;
    if (!bm[p] || counters[page].status == RETTO_B200_ERR_CAPACITY) return;
    const int y = p / pg.w, x = p - y * pg.w;
    const bool left0 = x > 0 && !bm[p - 1], right0 = x + 1 < pg.w && !bm[p + 1];
    if (!left0 && !right0) return;
    const RunRec* prun = runs ? runs + (size_t)page * 3 * RUN_CAP : nullptr;
    const RunRec* psorted = prun ? prun + RUN_CAP : nullptr;
    const int* plabel = prun ? reinterpret_cast<const int*>(prun + 2 * RUN_CAP) : nullptr;
    const int n_runs = counters[page].n_runs;
    const int* lab = labels + pg.px_base;
    const int root = root_of_fg_pixel(run_path != 0, psorted, plabel, n_runs, lab, p);
    // the background region a start / end faces: frame-connected -> the pixel starts the border; a hole -> only if this component does
    // NOT enclose it (an island inside another component's hole).  The encloser of a hole is the component of the foreground pixel left
    // of the hole's first pixel (= its root: labels are minimum raster indices).
    auto faces_outside = [&](int q) {
        const int br = bgl[pg.px_base + q];
        if (lab[br] == -2) return true;
        return root_of_fg_pixel(run_path != 0, psorted, plabel, n_runs, lab, br - 1) != root;
    };
    bool cand = false;
    if (left0 && faces_outside(p - 1)) cand = true;
    if (!cand && right0 && faces_outside(p + 1)) cand = true;
    if (!cand) return;
    atomicMin(&key_at[pg.px_base + root], p);
}
__global__ void outer_key_apply_kernel(HoleArgs h, const PageCounters* __restrict__ counters, const CompRec* __restrict__ comps, int max_comps,
                                       const int* __restrict__ key_at, BoxCand* __restrict__ cand) {
    const int page = h.flag_pages[blockIdx.y];
    const int id = blockIdx.x * blockDim.x + threadIdx.x;
    if (counters[page].status == RETTO_B200_ERR_CAPACITY || id >= counters[page].n_roots) return;
    const int nk = key_at[h.pages[page].px_base + comps[(size_t)page * max_comps + id].root];
    BoxCand& b = cand[(size_t)page * max_comps + id];
    if (nk == b.key) return;
    b.key = nk;
    if (nk == 0x7fffffff) { b.valid = 0; b.status = 6; b.st0 = 6; }   // only ever faces its own holes: find_contours never starts its outer border
}
__global__ void hole_restore_kernel(HoleArgs h, int* __restrict__ labels) {
    int page, p; DetPostPage pg;
    if (!hole_px(h, page, pg, p)) return;
    if (labels[pg.px_base + p] < -1) labels[pg.px_base + p] = -1;
}

// ---- I: per-page compaction + sorted_boxes (det_processor.rs:324-333) ----------------------------------------------
__device__ __forceinline__ bool box_less(const BoxCand& r1, const BoxCand& r2) {
    const float c1x = __fdiv_rn(__fadd_rn(r1.xy[0], r1.xy[4]), 2.0f), c1y = __fdiv_rn(__fadd_rn(r1.xy[1], r1.xy[5]), 2.0f);
    const float c2x = __fdiv_rn(__fadd_rn(r2.xy[0], r2.xy[4]), 2.0f), c2y = __fdiv_rn(__fadd_rn(r2.xy[1], r2.xy[5]), 2.0f);
    if (fabsf(__fsub_rn(c1y, c2y)) < 10.0f) return c1x < c2x;
    return c1y < c2y;
}

struct TraceRec { int key, status; int rect1[8]; float sside1, score; int rect2[8]; float dist; int n_off; };
__global__ void trace_copy_kernel(int n_pages, const PageCounters* __restrict__ counters, const BoxCand* __restrict__ cand, int max_comps,
                                  TraceRec* __restrict__ trace) {
    const int page = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (page >= n_pages || i >= counters[page].n_roots + counters[page].n_holes || i >= max_comps) return;
    const BoxCand& b = cand[(size_t)page * max_comps + i];
    TraceRec t;
    t.key = b.key; t.status = b.status; t.sside1 = b.sside1; t.score = b.score;
    for (int k = 0; k < 8; ++k) { t.rect1[k] = b.rect1[k]; t.rect2[k] = b.rect2[k]; }
    t.dist = b.dist; t.n_off = b.n_off;
    trace[(size_t)page * max_comps + i] = t;
}

// One warp per page.  The candidates are compacted by ballot into shared memory as (discovery key, centre x, centre y, index)
// tuples, lane 0 runs the two stable insertion sorts on those 16-byte tuples (the reference's comparator is not a strict weak
// order, so it is the ALGORITHM that has to be reproduced, not just an ordering), and the result is a permutation that
// pack_boxes_kernel reads through — no 140-byte BoxCand records are moved in global memory (0.033 -> see DESIGN.md).
#define PS_CAP 1024
__global__ void __launch_bounds__(32) page_sort_kernel(int n_pages, PageCounters* __restrict__ counters, BoxCand* __restrict__ cand, int max_comps,
                                                       int* __restrict__ order) {
    __shared__ int s_key[PS_CAP], s_idx[PS_CAP];
    __shared__ float s_cx[PS_CAP], s_cy[PS_CAP];
    const int page = blockIdx.x, lane = threadIdx.x;
    if (page >= n_pages) return;
    if (counters[page].status == RETTO_B200_ERR_CAPACITY) { if (lane == 0) counters[page].n_boxes = 0; return; }
    const int n = counters[page].n_roots + counters[page].n_holes;
    BoxCand* c = cand + (size_t)page * max_comps;
    int* ord = order + (size_t)page * max_comps;
    int m = 0;
    bool fits = true;
    for (int base = 0; base < n; base += 32) {
        const int i = base + lane;
        const bool v = i < n && c[i].valid;
        const unsigned bal = __ballot_sync(RT_FULL, v);
        const int pos = m + __popc(bal & ((1u << lane) - 1u));
        if (v && pos < PS_CAP) {
            s_idx[pos] = i; s_key[pos] = c[i].key;
            s_cx[pos] = __fdiv_rn(__fadd_rn(c[i].xy[0], c[i].xy[4]), 2.0f);
            s_cy[pos] = __fdiv_rn(__fadd_rn(c[i].xy[1], c[i].xy[5]), 2.0f);
        }
        m += __popc(bal);
    }
    if (m > PS_CAP) fits = false;
    __syncwarp();
    if (!fits) {   // more valid boxes than the shared-memory table holds: the record-moving path (one thread)
        if (lane == 0) {
            int mm = 0;
            for (int i = 0; i < n; ++i)
                if (c[i].valid) { if (mm != i) c[mm] = c[i]; ++mm; }
            for (int i = 1; i < mm; ++i) {
                const BoxCand v = c[i];
                int j = i;
                while (j > 0 && v.key < c[j - 1].key) { c[j] = c[j - 1]; --j; }
                c[j] = v;
            }
            for (int i = 1; i < mm; ++i) {
                const BoxCand v = c[i];
                int j = i;
                while (j > 0 && box_less(v, c[j - 1])) { c[j] = c[j - 1]; --j; }
                c[j] = v;
            }
            for (int i = 0; i < mm; ++i) ord[i] = i;
            counters[page].n_boxes = mm;
        }
        return;
    }
    if (lane == 0) {
        // contours come in find_contours discovery order
        for (int i = 1; i < m; ++i) {
            const int k = s_key[i], id = s_idx[i];
            const float x = s_cx[i], y = s_cy[i];
            int j = i;
            while (j > 0 && k < s_key[j - 1]) { s_key[j] = s_key[j - 1]; s_idx[j] = s_idx[j - 1]; s_cx[j] = s_cx[j - 1]; s_cy[j] = s_cy[j - 1]; --j; }
            s_key[j] = k; s_idx[j] = id; s_cx[j] = x; s_cy[j] = y;
        }
        // stable insertion sort == Rust's sort_by for n <= 20, and for any n when the comparator is consistent
        for (int i = 1; i < m; ++i) {
            const int id = s_idx[i];
            const float x = s_cx[i], y = s_cy[i];
            int j = i;
            while (j > 0 && (fabsf(__fsub_rn(y, s_cy[j - 1])) < 10.0f ? x < s_cx[j - 1] : y < s_cy[j - 1])) {
                s_idx[j] = s_idx[j - 1]; s_cx[j] = s_cx[j - 1]; s_cy[j] = s_cy[j - 1]; --j;
            }
            s_idx[j] = id; s_cx[j] = x; s_cy[j] = y;
        }
        counters[page].n_boxes = m;
    }
    __syncwarp();
    for (int i = lane; i < m; i += 32) ord[i] = s_idx[i];
}

// ---- J: dense packing ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) pack_offsets_kernel(int n_pages, const PageCounters* __restrict__ counters, int* __restrict__ offsets) {
    // exclusive prefix of the per-page box counts: one block, 256 pages per trip (warp scans + a scan of the warp totals)
    __shared__ int s_w[8];
    __shared__ int s_carry;
    if (blockIdx.x != 0) return;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (int base = 0; base < n_pages; base += 256) {
        const int p = base + (int)threadIdx.x;
        const int v = p < n_pages ? counters[p].n_boxes : 0;
        int x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(RT_FULL, x, o); if (lane >= o) x += t; }
        if (lane == 31) s_w[w] = x;
        __syncthreads();
        int before = 0, total = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) { const int t = s_w[k]; if (k < w) before += t; total += t; }
        const int carry = s_carry;
        if (p < n_pages) offsets[p] = carry + before + x - v;
        __syncthreads();
        if (threadIdx.x == 0) s_carry = carry + total;
        __syncthreads();
    }
    if (threadIdx.x == 0) offsets[n_pages] = s_carry;
}
__global__ void __launch_bounds__(128) pack_boxes_kernel(int n_pages, const PageCounters* __restrict__ counters, const int* __restrict__ offsets,
                                                          const BoxCand* __restrict__ cand, int max_comps, retto_b200_box* __restrict__ dense,
                                                          int cap, const int* __restrict__ order) {
    const int page = blockIdx.x;
    const int n = counters[page].n_boxes, base = offsets[page];
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        if (base + i >= cap) break;
        const BoxCand b = cand[(size_t)page * max_comps + order[(size_t)page * max_comps + i]];
        retto_b200_box o;
#pragma unroll
        for (int k = 0; k < 8; ++k) o.xy[k] = b.xy[k];
        o.score = b.score;
        dense[base + i] = o;
    }
}


// =====================================================================================================================
// Run-table CCL (default path).  Text probability maps are sparse and blobby: a 1280x1280 page has ~10^3 horizontal
// foreground runs against 1.6 * 10^6 pixels, so connected components are resolved on the RUN table, on chip, instead
// of with three more passes of dependent global accesses over the pixel label plane:
//   A' bitmap_runs2   threshold + 2x2 dilate -> bitmap, appends one record per run (per 128-px strip) to the page's run
//                     table, and counts the Euler number of the bitmap on the fly           [HBM: 4 B/px in, 1 B/px out]
//   B' ccl_runs       one block per page, everything in shared memory: sort the runs into raster order, union runs
//                     that touch (8-connectivity: previous row, x ranges within 1; same row for strip-split runs),
//                     components in raster order of their first pixel (== find_contours discovery order), per-component
//                     last row / discovery key, row-table allocation and per-row extremes
// The pixel label plane is not needed downstream; retto_b200_det_post_fetch_labels materialises it from the runs.
// A page with more runs than the on-chip table holds (noise) sends the batch to the pixel path (kernels A-G above).
#define RUN_ROW_MAX 256                    // runs in one image row handled on chip (insertion sort + linear neighbour scan)
#define RUN_MAX_H 4096                     // rows of the per-row index (== the cap on max_det_side)
#define RUN_X_MASK ((1 << 29) - 1)

// Same row walk as bitmap_runs_kernel (threshold, 2x2 dilate, bitmap store), but instead of a label plane it appends one
// record per horizontal run of the 128-px strip row, emitted by the lane that holds the run's last pixel (run start
// from the same ballot "carry" as the pixel path).  Runs are cut at strip boundaries; ccl_runs_kernel re-joins them.
template <bool VEC>
__global__ void __launch_bounds__(128) bitmap_runs2_kernel(const DetPostPage* __restrict__ pages, const int* __restrict__ tile_prefix,
                                                            int n_pages, int total_tiles, float thr, int dilate,
                                                            unsigned char* __restrict__ bitmap, RunRec* __restrict__ runs,
                                                            PageCounters* __restrict__ counters) {
    DetPostPage pg; int page, s, rb, tile;
    if (!tile_lookup(pages, tile_prefix, n_pages, total_tiles, 4, pg, page, s, rb, tile)) return;
    const int lane = threadIdx.x & 31;
    const int W = pg.w, H = pg.h;
    const int x0 = s * TILE_W, x = x0 + lane * 4;
    const int y0 = rb * TILE_H, y1 = min(y0 + TILE_H, H);
    unsigned char* bm = bitmap + pg.px_base;
    const float* __restrict__ prob = pg.prob;
    RunRec* prun = runs + (size_t)page * 3 * RUN_CAP;
    float nf = 0.0f;   // running NaN-propagating max of |v| over the loaded probabilities (see nf_max3 / PageCounters::nonfinite)

    // raw (float4, left-edge scalar) of one row; the threshold / shuffles are applied when the row is consumed, so the
    // loads of row y+1 are in flight while row y is processed
    struct Raw { float4 v; float l; };
    auto fetch = [&](int y) -> Raw {
        Raw r; r.v = make_float4(0.f, 0.f, 0.f, 0.f); r.l = 0.f;
        if (y < 0 || y >= y1) return r;
        const float* row = prob + (size_t)y * W;
        if (VEC) { if (x < W) r.v = __ldg(reinterpret_cast<const float4*>(row + x)); }
        else {
            if (x < W) r.v.x = __ldg(row + x);
            if (x + 1 < W) r.v.y = __ldg(row + x + 1);
            if (x + 2 < W) r.v.z = __ldg(row + x + 2);
            if (x + 3 < W) r.v.w = __ldg(row + x + 3);
        }
        if (lane == 0 && x0 > 0) r.l = __ldg(row + x0 - 1);
        return r;
    };
    auto bits5 = [&](const Raw& r, bool valid) -> unsigned {   // bit0 = t(x-1), bits1..4 = t(x..x+3)
        unsigned t = 0;
        if (valid) t = (r.v.x > thr ? 1u : 0u) | (r.v.y > thr ? 2u : 0u) | (r.v.z > thr ? 4u : 0u) | (r.v.w > thr ? 8u : 0u);
        if (VEC) { if (x >= W) t = 0; }
        else {
#pragma unroll
            for (int j = 0; j < 4; ++j) if (x + j >= W) t &= ~(1u << j);
        }
        unsigned left = __shfl_up_sync(RT_FULL, (t >> 3) & 1u, 1);
        if (lane == 0) left = (valid && x0 > 0 && r.l > thr) ? 1u : 0u;
        return (t << 1) | left;
    };
    unsigned prev = 0;
    if (dilate && y0 > 0) { const Raw r = fetch(y0 - 1); prev = bits5(r, true); }
    Raw nxt = fetch(y0);
    for (int y = y0; y < y1; ++y) {
        const Raw rawc = nxt;
        nxt = fetch(y + 1);
        nf = nf_max3(nf_max3(nf, rawc.v.x, rawc.v.y), rawc.v.z, rawc.v.w);   // lanes beyond the page hold zeros
        const unsigned cur = bits5(rawc, true);
        unsigned nib;
        if (dilate) {
            const unsigned m = cur | prev;          // vertical OR
            nib = ((m >> 1) | m) & 0xfu;            // bit j = m[j+1] | m[j]  (pixel j sits at bit j+1)
        } else nib = (cur >> 1) & 0xfu;
        prev = cur;
        if (VEC) { if (x >= W) nib = 0; }
        else {
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (x + j >= W) nib &= ~(1u << j);
        }
        if (VEC) {
            if (x < W) *reinterpret_cast<unsigned*>(bm + (size_t)y * W + x) =
                ((nib & 1u) ? 0xffu : 0u) | ((nib & 2u) ? 0xff00u : 0u) | ((nib & 4u) ? 0xff0000u : 0u) | ((nib & 8u) ? 0xff000000u : 0u);
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (x + j < W) bm[(size_t)y * W + x + j] = (nib & (1u << j)) ? 255 : 0;
        }
        if (!__ballot_sync(RT_FULL, nib != 0)) continue;
        // ---- runs of this strip row
        const unsigned notfull = __ballot_sync(RT_FULL, nib != 0xfu);
        const unsigned below = notfull & ((1u << lane) - 1u);
        const int l = below ? 31 - __clz(below) : 0;
        const unsigned nibl = __shfl_sync(RT_FULL, nib, l);
        unsigned nb = __shfl_down_sync(RT_FULL, nib & 1u, 1);            // pixel x+4; the strip ends after lane 31
        if (lane == 31) nb = 0;
        int start = below ? x0 + 4 * l + (31 - __clz((~nibl) & 0xfu)) + 1 : x0;
        const unsigned ext = nib | (nb << 4);
        int ends[2], starts[2], cnt = 0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (nib & (1u << j)) {
                if (!((ext >> (j + 1)) & 1u)) { starts[cnt] = start; ends[cnt] = x + j; ++cnt; }   // at most two runs end in 4 pixels
            } else start = x + j + 1;
        }
        int incl = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(RT_FULL, incl, o); if (lane >= o) incl += v; }
        const int total = __shfl_sync(RT_FULL, incl, 31);
        int base = 0;
        if (lane == 31) base = atomicAdd(&counters[page].n_runs, total);
        base = __shfl_sync(RT_FULL, base, 31) + incl - cnt;
        for (int k = 0; k < cnt; ++k)
            if (base + k < RUN_CAP) prun[base + k] = RunRec{y * W + starts[k], ends[k]};
    }
    if (__any_sync(RT_FULL, nf_bad(nf)) && lane == 0) atomicOr(&counters[page].nonfinite, 1);
}

// A'' (default): the same stage with one warp per 256-px strip and the row held as EIGHT warp-uniform 32-bit words —
// lane l loads pixels x0 + 32 j + l (eight fully coalesced 128-B loads per row), and a ballot per load turns the
// threshold test straight into the bit word of 32 consecutive pixels.  Dilation, run starts and run ends are then a
// handful of funnel shifts on uniform words instead of per-lane nibble logic + shuffles (bitmap_runs2 issues ~115
// instructions per 128-px row and is issue-bound at 73 % issue utilisation; this form needs about half per pixel).
// Runs are buffered per warp in shared memory and appended to the page's run table with ONE atomic per tile.
#define BR3_BUF 96
template <bool ALIGNED, int NW, int MINB, bool PROBE = true>   // NW words of 32 pixels per strip row (4: 128-px strips on the tile grid of the pixel path, 8: 256-px strips)
__global__ void __launch_bounds__(128, MINB) bitmap_runs3_kernel(const DetPostPage* __restrict__ pages, const int* __restrict__ tile_prefix,
                                                            int n_pages, int total_tiles, float thr, int dilate,
                                                            unsigned char* __restrict__ bitmap, RunRec* __restrict__ runs,
                                                            PageCounters* __restrict__ counters) {
    constexpr int SW = 32 * NW;
    __shared__ RunRec s_buf[4][BR3_BUF];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int tile = blockIdx.x * 4 + wib;
    if (tile >= total_tiles) return;
    const int page = rt_find_segment(tile_prefix, n_pages, tile);
    const DetPostPage pg = pages[page];
    const int W = pg.w, H = pg.h;
    const int strips = (W + SW - 1) / SW;
    const int lt = tile - tile_prefix[page];
    const int rb = lt / strips, s = lt - rb * strips;
    const int x0 = s * SW;
    const int y0 = rb * TILE_H, y1 = min(y0 + TILE_H, H);
    unsigned char* bm = bitmap + pg.px_base;
    const float* __restrict__ prob = pg.prob;
    RunRec* prun = runs + (size_t)page * 3 * RUN_CAP;
    RunRec* buf = s_buf[wib];
    int nbuf = 0;
    auto flush = [&]() {
        __syncwarp();
        int base = 0;
        if (lane == 0) base = atomicAdd(&counters[page].n_runs, nbuf);
        base = __shfl_sync(RT_FULL, base, 0);
        for (int i = lane; i < nbuf; i += 32)
            if (base + i < RUN_CAP) prun[base + i] = buf[i];
        __syncwarp();
        nbuf = 0;
    };
    const int nv = W - x0;                  // pixels of the strip inside the page (only the last strip of a row is partial)
    const bool partial = nv < SW;
    float nf = 0.0f;   // running NaN-propagating max of |v| over the loaded probabilities (see nf_max3 / PageCounters::nonfinite)
    struct Raw { float v[NW]; float l; };
    auto fetch = [&](int y) -> Raw {
        Raw r;
#pragma unroll
        for (int j = 0; j < NW; ++j) r.v[j] = -3.0e38f;   // outside the page: below every threshold (finite: the non-finite probe sees these too)
        r.l = -3.0e38f;
        if (y < 0 || y >= y1) return r;
        const float* row = prob + (size_t)y * W + x0 + lane;
        if (!partial) {   // full strip (all but the last of a page row): unpredicated loads
#pragma unroll
            for (int j = 0; j < NW; ++j) r.v[j] = __ldg(row + 32 * j);
        } else {
#pragma unroll
            for (int j = 0; j < NW; ++j) if (32 * j + lane < nv) r.v[j] = __ldg(row + 32 * j);
        }
        if (lane == 0 && x0 > 0) r.l = __ldg(row - 1);
        return r;
    };
    unsigned prev[NW], prevL = 0;
#pragma unroll
    for (int j = 0; j < NW; ++j) prev[j] = 0;
    if (dilate && y0 > 0) {
        const Raw r = fetch(y0 - 1);
#pragma unroll
        for (int j = 0; j < NW; ++j) prev[j] = __ballot_sync(RT_FULL, r.v[j] > thr);
        prevL = __ballot_sync(RT_FULL, r.l > thr) & 1u;
    }
    Raw nxt = fetch(y0);
    for (int y = y0; y < y1; ++y) {
        const Raw rc = nxt;
        nxt = fetch(y + 1);
        if (PROBE) {   // probe the row being CONSUMED (its loads have landed): probing inside fetch() made the warp wait for the loads it had
                       // just issued and cost 0.09 ms per 256 pages (the next row's loads are meant to fly during this row's work)
#pragma unroll
            for (int j = 0; j < NW; j += 2) nf = nf_max3(nf, rc.v[j], rc.v[j + 1]);
        }
        unsigned d[NW];
        {
            const unsigned curL = __ballot_sync(RT_FULL, rc.l > thr) & 1u;
            unsigned carry = dilate ? (curL | prevL) << 31 : 0u;
            prevL = curL;
#pragma unroll
            for (int j = 0; j < NW; ++j) {
                const unsigned c = __ballot_sync(RT_FULL, rc.v[j] > thr);
                if (dilate) {
                    const unsigned m = c | prev[j];
                    d[j] = m | __funnelshift_l(carry, m, 1);   // out(x) = m(x) | m(x - 1)
                    carry = m;
                } else d[j] = c;
                prev[j] = c;
            }
            if (partial && dilate) {   // the dilation may reach one pixel past the page
#pragma unroll
                for (int j = 0; j < NW; ++j) d[j] &= nv >= 32 * (j + 1) ? 0xFFFFFFFFu : (nv <= 32 * j ? 0u : (1u << (nv - 32 * j)) - 1u);
            }
        }
        // bitmap: lane l stores pixels x0 + NW l .. + NW - 1
        {
            unsigned wsel;
            if (NW == 8) {
                const unsigned a0 = (lane & 4) ? d[1] : d[0], a1 = (lane & 4) ? d[3] : d[2], a2 = (lane & 4) ? d[5 % NW] : d[4 % NW], a3 = (lane & 4) ? d[7 % NW] : d[6 % NW];
                const unsigned b0 = (lane & 8) ? a1 : a0, b1 = (lane & 8) ? a3 : a2;
                wsel = (lane & 16) ? b1 : b0;
            } else {
                const unsigned a0 = (lane & 8) ? d[1] : d[0], a1 = (lane & 8) ? d[3] : d[2];
                wsel = (lane & 16) ? a1 : a0;
            }
            const unsigned bits = wsel >> ((NW * lane) & 31);
            const unsigned lo = (((bits & 0xFu) * 0x00204081u) & 0x01010101u) * 0xFFu;
            const int x = x0 + NW * lane;
            if (NW == 8) {
                const unsigned hi = ((((bits >> 4) & 0xFu) * 0x00204081u) & 0x01010101u) * 0xFFu;
                if (ALIGNED) { if (x < W) *reinterpret_cast<uint2*>(bm + (size_t)y * W + x) = make_uint2(lo, hi); }
                else {
#pragma unroll
                    for (int k = 0; k < 8; ++k)
                        if (x + k < W) bm[(size_t)y * W + x + k] = (unsigned char)((k < 4 ? lo >> (8 * k) : hi >> (8 * (k - 4))) & 0xFFu);
                }
            } else {
                if (ALIGNED) { if (x < W) *reinterpret_cast<unsigned*>(bm + (size_t)y * W + x) = lo; }
                else {
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        if (x + k < W) bm[(size_t)y * W + x + k] = (unsigned char)((lo >> (8 * k)) & 0xFFu);
                }
            }
        }
        unsigned any = 0;
#pragma unroll
        for (int j = 0; j < NW; ++j) any |= d[j];
        if (!any) continue;
        // runs of this strip row: starts / ends alternate, so the k-th start pairs with the k-th end (all values warp-uniform)
        int open_start = -1;
#pragma unroll
        for (int j = 0; j < NW; ++j) {
            if (!d[j]) continue;   // no start and no end in an empty word (warp-uniform)
            unsigned st = d[j] & ~__funnelshift_l(j ? d[j ? j - 1 : 0] : 0u, d[j], 1);
            unsigned en = d[j] & ~__funnelshift_r(d[j], j < NW - 1 ? d[j < NW - 1 ? j + 1 : j] : 0u, 1);
            const int xb = x0 + 32 * j;
            while (en) {
                int sx;
                if (open_start >= 0) { sx = open_start; open_start = -1; }
                else { sx = xb + __ffs(st) - 1; st &= st - 1; }
                const int ex = xb + __ffs(en) - 1;
                en &= en - 1;
                if (nbuf == BR3_BUF) flush();
                if (lane == 0) buf[nbuf] = RunRec{y * W + sx, ex};
                ++nbuf;
            }
            if (st) open_start = xb + __ffs(st) - 1;
        }
    }
    if (nbuf) flush();
    if (__any_sync(RT_FULL, nf_bad(nf)) && lane == 0) atomicOr(&counters[page].nonfinite, 1);
}

// B': one block per page.  Shared memory: sorted runs (key, x1) | parent | per-row index.
// find with path halving: every visited node is re-pointed at its grandparent (atomicMin keeps the parent pointers
// monotonically decreasing under concurrent unions, so a shortcut can never undo a link)
__device__ __forceinline__ int sm_find(int* par, int a) {
    int p = par[a];
    while (p != a) {
        const int g = par[p];
        if (g != p) atomicMin(&par[a], g);
        a = p; p = g;
    }
    return a;
}
__device__ __forceinline__ void sm_union(int* par, int a, int b) {
    bool done;
    do {
        a = sm_find(par, a);
        b = sm_find(par, b);
        if (a < b) { const int old = atomicMin(&par[b], a); done = (old == b); b = old; }
        else if (b < a) { const int old = atomicMin(&par[a], b); done = (old == a); a = old; }
        else done = true;
    } while (!done);
}
__device__ __forceinline__ int block_excl_scan(int v, int* s_scan, int* total) {   // exclusive prefix over the block's threads
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    int x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(RT_FULL, x, o); if (lane >= o) x += t; }
    if (lane == 31) s_scan[w] = x;
    __syncthreads();
    if (w == 0) {
        int t = lane < nw ? s_scan[lane] : 0;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(RT_FULL, t, o); if (lane >= o) t += u; }
        s_scan[lane] = t;
    }
    __syncthreads();
    const int incl = x + (w ? s_scan[w - 1] : 0);
    *total = s_scan[nw - 1];
    __syncthreads();
    return incl - v;
}
// (Measured and dropped in round 2: a first tier of 256-thread / 30-KB blocks so that a 1024-page batch is resident at once instead
// of in 3.5 waves of these blocks — 0.54 ms against 0.41 ms on config 2: the kernel is bound by the work per page, not by the waves.)
__global__ void __launch_bounds__(1024, 2) ccl_runs_kernel(const DetPostPage* __restrict__ pages, PageCounters* __restrict__ counters,
                                                            RunRec* __restrict__ runs, CompRec* __restrict__ comps, int2* __restrict__ rowtab,
                                                            int max_comps) {
    constexpr int CAP = RUN_CAP;
    constexpr int MAXH = RUN_MAX_H;
    extern __shared__ int s_mem[];
    __shared__ int s_scan[32];
    __shared__ int s_flag;
    const int page = blockIdx.x;
    const DetPostPage pg = pages[page];
    const int W = pg.w, H = pg.h;
    const int n = counters[page].n_runs;
    if (n > CAP || H > MAXH) {
        if (threadIdx.x == 0) counters[page].fallback = 1;
        return;
    }
    int2* srt = reinterpret_cast<int2*>(s_mem);          // CAP x (key, x1), raster order
    int* par = s_mem + 2 * CAP;                          // CAP
    int* rowp = par + CAP;                               // MAXH + 1: first run of every row (rowp[H] = n)
    RunRec* praw = runs + (size_t)page * 3 * RUN_CAP;
    if (threadIdx.x == 0) s_flag = 0;
    for (int y = threadIdx.x; y <= H; y += blockDim.x) rowp[y] = 0;
    __syncthreads();
    // counting sort by row: histogram -> inclusive scan (row ends) -> scatter from the back (leaves the row starts)
    for (int i = threadIdx.x; i < n; i += blockDim.x) atomicAdd(&rowp[praw[i].key / W], 1);
    __syncthreads();
    {
        const int per = (H + (int)blockDim.x - 1) / (int)blockDim.x;
        const int r0 = min(H, (int)threadIdx.x * per), r1 = min(H, r0 + per);
        int sum = 0;
        for (int y = r0; y < r1; ++y) { sum += rowp[y]; if (rowp[y] > RUN_ROW_MAX) s_flag = 1; }
        int tot;
        int run = block_excl_scan(sum, s_scan, &tot);
        for (int y = r0; y < r1; ++y) { run += rowp[y]; rowp[y] = run; }   // inclusive: end of row y
        if (threadIdx.x == 0) rowp[H] = n;
    }
    __syncthreads();
    if (s_flag) {   // a row with hundreds of runs (noise): the pixel path handles it
        if (threadIdx.x == 0) counters[page].fallback = 1;
        return;
    }
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const RunRec r = praw[i];
        const int pos = atomicSub(&rowp[r.key / W], 1) - 1;
        srt[pos] = make_int2(r.key, r.x1);
    }
    __syncthreads();
    // rows are short: one thread sorts one row by x0 (insertion sort)
    for (int y = threadIdx.x; y < H; y += blockDim.x) {
        const int a = rowp[y], b = rowp[y + 1];
        for (int i = a + 1; i < b; ++i) {
            const int2 v = srt[i];
            int j = i;
            while (j > a && srt[j - 1].x > v.x) { srt[j] = srt[j - 1]; --j; }
            srt[j] = v;
        }
    }
    for (int i = threadIdx.x; i < n; i += blockDim.x) par[i] = i;
    __syncthreads();
    // unions over touching pieces (8-connectivity: previous row, x ranges within one pixel; same row: the pieces of a run
    // cut at strip boundaries).  The Euler number of the bitmap (#components - #holes) is #runs - #touching pairs of
    // WHOLE runs in consecutive rows: every independent cycle of the run adjacency graph encloses exactly one background
    // region (counted on whole runs — the pieces of a cut run would add cycles that enclose nothing).
    auto is_head = [&](int i, int row_first) { return !(i > row_first && srt[i - 1].y + 1 == srt[i].x - (srt[i].x / W) * W); };
    auto chain_end = [&](int i, int row_end) {   // last pixel of the whole run that piece i starts
        int e = srt[i].y;
        for (int k = i + 1; k < row_end && srt[k].x - (srt[k].x / W) * W == e + 1; ++k) e = srt[k].y;
        return e;
    };
    int pairs = 0, heads = 0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const int2 r = srt[i];
        const int y = r.x / W, xa = r.x - y * W, xb = r.y;
        const bool head = !(i > rowp[y] && srt[i - 1].y + 1 == xa);
        if (!head) sm_union(par, i, i - 1);
        if (y > 0) {
            const int lim = (y - 1) * W + min(xb + 1, W - 1);
            for (int j = rowp[y - 1]; j < rowp[y]; ++j) {
                const int2 q = srt[j];
                if (q.x > lim) break;
                if (q.y >= xa - 1) sm_union(par, i, j);
            }
        }
        if (head) {
            ++heads;
            if (y > 0) {
                const int b = chain_end(i, rowp[y + 1]);
                for (int j = rowp[y - 1]; j < rowp[y]; ++j) {
                    const int c = srt[j].x - (y - 1) * W;
                    if (c > b + 1) break;
                    if (!is_head(j, rowp[y - 1])) continue;
                    if (chain_end(j, rowp[y]) >= xa - 1) ++pairs;
                }
            }
        }
    }
    int tot_pairs, tot_heads;
    block_excl_scan(pairs, s_scan, &tot_pairs);
    block_excl_scan(heads, s_scan, &tot_heads);
    {   // flatten: every run points at its root.  Roots are collected first and written after a barrier, so no plain store races with
        // another thread's find (compute-sanitizer racecheck is clean; a concurrent store would be benign — every value on the way
        // leads to the same root — but it is cheaper to keep the tool quiet than to explain 712 hazards)
        constexpr int PER = RUN_CAP / 1024;
        int roots[PER];
#pragma unroll
        for (int q = 0; q < PER; ++q) { const int i = (int)threadIdx.x + q * (int)blockDim.x; roots[q] = i < n ? sm_find(par, i) : 0; }
        __syncthreads();
#pragma unroll
        for (int q = 0; q < PER; ++q) { const int i = (int)threadIdx.x + q * (int)blockDim.x; if (i < n) par[i] = roots[q]; }
    }
    __syncthreads();
    // dense component ids in raster order of the root run
    const int per = (n + (int)blockDim.x - 1) / (int)blockDim.x;
    const int b0 = min(n, (int)threadIdx.x * per), e0 = min(n, b0 + per);
    int cntr = 0;
    for (int i = b0; i < e0; ++i) cntr += (par[i] == i);
    int n_roots;
    int id = block_excl_scan(cntr, s_scan, &n_roots);
    if (n_roots > max_comps) {
        if (threadIdx.x == 0) { counters[page].status = RETTO_B200_ERR_CAPACITY; counters[page].n_roots = 0; counters[page].euler = tot_heads - tot_pairs; }
        return;
    }
    CompRec* c = comps + (size_t)page * max_comps;
    for (int i = b0; i < e0; ++i)
        if (par[i] == i) {
            const int key = srt[i].x;
            CompRec cr;
            cr.root = key; cr.ymax = key / W; cr.xmin = 0; cr.xmax = 0; cr.row_off = 0; cr.key = 0x7fffffff; cr.ymin = key / W; cr.pad = 0;
            c[id] = cr;
            par[i] = -1 - id;      // roots now carry their component id
            ++id;
        }
    __syncthreads();
    auto comp_of = [&](int i) { const int p = par[i]; return p < 0 ? -1 - p : -1 - par[p]; };
    // per-component last row and discovery key (imageproc's scan, RECALLED: a border starts at x > 0 with a zero to the
    // left, or at x + 1 < width with a zero to the right — see ccl_flatten_kernel); a run cut at a strip boundary has no
    // true start / end there
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const int2 r = srt[i];
        const int y = r.x / W, xa = r.x - y * W, xb = r.y;
        const bool tstart = !(i > rowp[y] && srt[i - 1].y + 1 == xa);
        const bool tend = !(i + 1 < rowp[y + 1] && srt[i + 1].x == r.x + (xb - xa) + 1);
        CompRec* cr = c + comp_of(i);
        atomicMax(&cr->ymax, y);
        if (tstart && xa > 0) atomicMin(&cr->key, r.x);
        if (tend && xb + 1 < W) atomicMin(&cr->key, y * W + xb);
    }
    __syncthreads();
    // row-table allocation: exclusive prefix over the components of (ymax - ymin + 1)
    const int perc = (n_roots + (int)blockDim.x - 1) / (int)blockDim.x;
    const int cb0 = min(n_roots, (int)threadIdx.x * perc), ce0 = min(n_roots, cb0 + perc);
    int sum = 0;
    for (int i = cb0; i < ce0; ++i) sum += c[i].ymax - c[i].ymin + 1;
    int row_total;
    int off = block_excl_scan(sum, s_scan, &row_total);
    for (int i = cb0; i < ce0; ++i) { c[i].row_off = off; off += c[i].ymax - c[i].ymin + 1; }
    if (threadIdx.x == 0) {
        counters[page].n_roots = n_roots;
        counters[page].euler = tot_heads - tot_pairs;
        counters[page].row_total = row_total;
        if (row_total > ROWCAP) counters[page].status = RETTO_B200_ERR_CAPACITY;
    }
    const int total_rows = min(row_total, ROWCAP);
    int2* rt = rowtab + (size_t)page * ROWCAP;
    for (int i = threadIdx.x; i < total_rows; i += blockDim.x) rt[i] = make_int2(0x7fffffff, -1);
    __syncthreads();
    if (row_total > ROWCAP) return;
    // per-row extremes (true run starts / ends only, like run_end_kernel<1>) + the sorted table and the labels for the tap
    RunRec* psorted = praw + RUN_CAP;
    int* plabel = reinterpret_cast<int*>(praw + 2 * RUN_CAP);
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const int2 r = srt[i];
        const int y = r.x / W, xa = r.x - y * W, xb = r.y;
        const bool tstart = !(i > rowp[y] && srt[i - 1].y + 1 == xa);
        const bool tend = !(i + 1 < rowp[y + 1] && srt[i + 1].x == r.x + (xb - xa) + 1);
        const CompRec cr = c[comp_of(i)];
        const int ridx = cr.row_off + (y - cr.ymin);
        if (tstart) atomicMin(&rt[ridx].x, xa);
        if (tend) atomicMax(&rt[ridx].y, xb);
        psorted[i] = RunRec{r.x, xb};
        plabel[i] = cr.root;
    }
}

// label plane from the run table (parity tap): background -1, every run pixel = raster index of its component's first pixel
__global__ void labels_from_runs_kernel(const RunRec* __restrict__ sorted, const int* __restrict__ label, int n, int W, int* __restrict__ labels) {
    const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (i >= n) return;
    const int key = sorted[i].key, xb = sorted[i].x1, lab = label[i];
    const int y = key / W, xa = key - y * W;
    for (int x = xa + lane; x <= xb; x += 32) labels[(size_t)y * W + x] = lab;
}

__global__ void zero_counters_kernel(PageCounters* c, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { PageCounters z; memset(&z, 0, sizeof(z)); c[i] = z; }
}

// pixel-plane CCL chain (kernels A-G): the fallback of the run-table path, and the path under RETTO_B200_PIXEL_CCL=1.
// Ends with the early counter read-back (event) enqueued before the row-extreme kernel.
static retto_b200_status dp_pixel_ccl(retto_b200_ctx* ctx) {
    retto_b200_ctx::DpRun& R = ctx->dp;
    const int n = R.n, total_tiles = R.total_tiles, max_comps = ctx->cfg.max_components_per_page;
    cudaStream_t st = ctx->stream;
    const DetPostPage* d_pages = ctx->d_dp_pages.as<DetPostPage>();
    const int* d_tile_prefix = reinterpret_cast<const int*>(ctx->d_dp_pages.as<char>() + sizeof(DetPostPage) * n);
    PageCounters* d_cnt = ctx->d_dp_counters.as<PageCounters>();
    unsigned char* d_bm = ctx->d_bitmap.as<unsigned char>();
    int* d_lab = ctx->d_labels.as<int>();
    int* d_cid = ctx->d_cid_at.as<int>();
    int* d_key = ctx->d_key_at.as<int>();
    unsigned char* d_tf = ctx->d_tileflags.as<unsigned char>();
    int* d_roots = ctx->d_roots.as<int>();
    CompRec* d_comps = ctx->d_comps.as<CompRec>();
    int2* d_rowtab = ctx->d_rowtab.as<int2>();
    ctx->dp_run_path = false;
    RT_LAUNCH_BEGIN(ctx, "zero_counters_kernel");
    zero_counters_kernel<<<(n + 127) / 128, 128, 0, st>>>(d_cnt, n);
    RT_LAUNCH_CHECK(ctx);
    const int tgrid = (total_tiles + 3) / 4;
    RT_LAUNCH_BEGIN(ctx, "bitmap_runs_kernel");
    if (R.vec)
        bitmap_runs_kernel<true><<<tgrid, 128, 0, st>>>(d_pages, d_tile_prefix, n, total_tiles, ctx->cfg.det_thresh, ctx->cfg.det_dilation_2x2, d_bm, d_lab, d_tf, d_cid, d_key, d_cnt);
    else
        bitmap_runs_kernel<false><<<tgrid, 128, 0, st>>>(d_pages, d_tile_prefix, n, total_tiles, ctx->cfg.det_thresh, ctx->cfg.det_dilation_2x2, d_bm, d_lab, d_tf, d_cid, d_key, d_cnt);
    RT_LAUNCH_CHECK(ctx);
    RT_LAUNCH_BEGIN(ctx, "ccl_merge_kernel");
    ccl_merge_kernel<<<tgrid, 128, 0, st>>>(d_pages, d_tile_prefix, n, total_tiles, d_bm, d_lab, d_tf, d_cnt);
    RT_LAUNCH_CHECK(ctx);
    RT_LAUNCH_BEGIN(ctx, "ccl_flatten_kernel");
    ccl_flatten_kernel<<<tgrid, 128, 0, st>>>(d_pages, d_tile_prefix, n, total_tiles, d_bm, d_lab, d_tf, d_cnt, d_roots, max_comps, d_cid, d_key);
    RT_LAUNCH_CHECK(ctx);
    RT_LAUNCH_BEGIN(ctx, "comp_sort_kernel");
    comp_sort_kernel<<<n, 1024, 0, st>>>(d_pages, d_cnt, d_roots, d_comps, d_cid, d_key, d_rowtab, max_comps);
    RT_LAUNCH_CHECK(ctx);
    // The geometry grid and the hole-border decision need the per-page component counts (n_roots, euler, status), which
    // are final after comp_sort: their read-back is enqueued BEFORE the row-extreme kernel and the host waits on an
    // event behind the copy only, so the round trip hides under run_end_kernel instead of idling the GPU.
    RT_CUDA_OK(ctx, cudaMemcpyAsync(ctx->h_dp.as<PageCounters>(), d_cnt, sizeof(PageCounters) * n, cudaMemcpyDeviceToHost, st));
    RT_CUDA_OK(ctx, cudaEventRecord(ctx->ev_dp, st));
    RT_LAUNCH_BEGIN(ctx, "run_end_kernel<1>");
    run_end_kernel<1><<<tgrid, 128, 0, st>>>(d_pages, d_tile_prefix, n, total_tiles, d_bm, d_lab, d_tf, d_cid, d_comps, d_rowtab, d_cnt, max_comps);
    RT_LAUNCH_CHECK(ctx);
    return RETTO_B200_OK;
}

// ---- host ------------------------------------------------------------------------------------------------------------------
// det_postprocess runs in three host steps so that a caller with several page batches in flight (session.cu lanes) can
// do other work at the two points where the host needs numbers from the device:
//   begin : threshold/dilate, CCL, component table, row extremes enqueued; early counter read-back behind an event
//   mid   : waits for that event only (it hides under run_end_kernel), then geometry, hole borders, sort, pack and
//           the result read-back are enqueued
//   end   : one stream sync, results to the caller's arrays
retto_b200_status rt_det_post_begin(retto_b200_ctx* ctx, const retto_b200_det_post_desc* h_descs, int32_t n, int32_t max_boxes_total) {
    retto_b200_ctx::DpRun& R = ctx->dp;
    R = retto_b200_ctx::DpRun{};
    R.n = n;
    ctx->dp_pages.clear();
    if (n == 0) return RETTO_B200_OK;
    const int max_comps = ctx->cfg.max_components_per_page;
    std::vector<int> tile_prefix(n + 1, 0), tile2_prefix(n + 1, 0);   // 128-px strips (pixel path, bitmap_runs2) / 256-px strips (bitmap_runs3)
    long long px = 0;
    bool vec = true, a8 = true;
    int w_or = 0;
    for (int i = 0; i < n; ++i) {
        const retto_b200_det_post_desc& d = h_descs[i];
        if (!d.d_prob || d.h <= 0 || d.w <= 0 || d.ori_h <= 0 || d.ori_w <= 0 || (long long)d.h * d.w > 0x7fffffffLL) {
            ctx->set_error("det_postprocess: bad descriptor " + std::to_string(i));
            return RETTO_B200_ERR_INVALID_ARG;
        }
        DetPostPage pg;
        pg.prob = d.d_prob; pg.h = d.h; pg.w = d.w; pg.ori_h = d.ori_h; pg.ori_w = d.ori_w;
        pg.strips = (d.w + TILE_W - 1) / TILE_W;
        pg.rowblocks = (d.h + TILE_H - 1) / TILE_H;
        pg.tile_base = tile_prefix[i];
        pg.px_base = px;
        pg.comp_base = i * max_comps;
        pg.box_base = 0;
        tile_prefix[i + 1] = tile_prefix[i] + pg.strips * pg.rowblocks;
        tile2_prefix[i + 1] = tile2_prefix[i] + ((d.w + 255) / 256) * pg.rowblocks;
        px += ((long long)d.h * d.w + 15) & ~15LL;
        if ((d.w & 3) || ((uintptr_t)d.d_prob & 15)) vec = false;
        if (d.w & 7) a8 = false;
        w_or |= d.w;
        ctx->dp_pages.push_back(pg);
    }
    const int total_tiles = tile_prefix[n];
    ctx->dp_labels_final.assign(n, 0);
    cudaStream_t st = ctx->stream;
    // device state
    {
        std::vector<char> blob(sizeof(DetPostPage) * n + 2 * sizeof(int) * (n + 1));
        memcpy(blob.data(), ctx->dp_pages.data(), sizeof(DetPostPage) * n);
        memcpy(blob.data() + sizeof(DetPostPage) * n, tile_prefix.data(), sizeof(int) * (n + 1));
        memcpy(blob.data() + sizeof(DetPostPage) * n + sizeof(int) * (n + 1), tile2_prefix.data(), sizeof(int) * (n + 1));
        RT_TRY(rt_upload(ctx, ctx->d_dp_pages, blob.data(), blob.size()));
    }
    const DetPostPage* d_pages = ctx->d_dp_pages.as<DetPostPage>();
    const int* d_tile_prefix = reinterpret_cast<const int*>(ctx->d_dp_pages.as<char>() + sizeof(DetPostPage) * n);
    const int* d_tile2_prefix = d_tile_prefix + (n + 1);
    RT_CUDA_OK(ctx, ctx->d_dp_counters.ensure(sizeof(PageCounters) * n + sizeof(int) * (n + 1), st));
    RT_CUDA_OK(ctx, ctx->d_bitmap.ensure((size_t)px, st));
    RT_CUDA_OK(ctx, ctx->d_labels.ensure((size_t)px * 4, st));
    RT_CUDA_OK(ctx, ctx->d_cid_at.ensure((size_t)px * 4, st));
    RT_CUDA_OK(ctx, ctx->d_key_at.ensure((size_t)px * 4, st));
    RT_CUDA_OK(ctx, ctx->d_tileflags.ensure((size_t)total_tiles, st));
    RT_CUDA_OK(ctx, ctx->d_roots.ensure(sizeof(int) * (size_t)n * max_comps, st));
    RT_CUDA_OK(ctx, ctx->d_comps.ensure(sizeof(CompRec) * (size_t)n * max_comps, st));
    RT_CUDA_OK(ctx, ctx->d_rowtab.ensure(sizeof(int2) * (size_t)n * ROWCAP * 3, st));  // row table + 2x hull scratch
    RT_CUDA_OK(ctx, ctx->d_cand.ensure(sizeof(BoxCand) * (size_t)n * max_comps, st));
    PageCounters* d_cnt = ctx->d_dp_counters.as<PageCounters>();
    int* d_offsets = reinterpret_cast<int*>(ctx->d_dp_counters.as<char>() + sizeof(PageCounters) * n);
    unsigned char* d_bm = ctx->d_bitmap.as<unsigned char>();
    int* d_lab = ctx->d_labels.as<int>();
    int* d_cid = ctx->d_cid_at.as<int>();
    int* d_key = ctx->d_key_at.as<int>();
    unsigned char* d_tf = ctx->d_tileflags.as<unsigned char>();
    int* d_roots = ctx->d_roots.as<int>();
    CompRec* d_comps = ctx->d_comps.as<CompRec>();
    int2* d_rowtab = ctx->d_rowtab.as<int2>();
    int2* d_hull = d_rowtab + (size_t)n * ROWCAP;
    BoxCand* d_cand = ctx->d_cand.as<BoxCand>();

    (void)d_lab; (void)d_cid; (void)d_key; (void)d_tf; (void)d_roots; (void)d_rowtab; (void)d_hull; (void)d_cand; (void)d_offsets; (void)d_bm; (void)d_comps;
    const int cap = std::max(max_boxes_total, 0);
    const size_t hdr_bytes = (sizeof(PageCounters) * n + sizeof(int) * (n + 1) + 63) & ~size_t(63);
    RT_CUDA_OK(ctx, ctx->h_dp.ensure(2 * hdr_bytes + sizeof(retto_b200_box) * (size_t)std::max(cap, 1)));
    if (!ctx->ev_dp) RT_CUDA_OK(ctx, cudaEventCreateWithFlags(&ctx->ev_dp, cudaEventDisableTiming));
    R.cap = cap; R.total_tiles = total_tiles; R.hdr_bytes = hdr_bytes; R.vec = vec; R.total_tiles2 = tile2_prefix[n]; R.a8 = a8; R.w_or = w_or;
    // default: run-table CCL; RETTO_B200_PIXEL_CCL=1 (tests) or a page with too many runs: the pixel-plane passes
    ctx->dp_run_path = getenv("RETTO_B200_PIXEL_CCL") == nullptr;
    if (ctx->dp_run_path) {
        bool& attr_set = ctx->ccl_runs_attr_set;   // per context = per device: a process may drive several GPUs (retto_b200/cli.py --gpus N)
        const int smem = RUN_CAP * 12 + (RUN_MAX_H + 1) * 4;   // sorted runs (8 B) + parent per run, per-row index
        if (!attr_set) { RT_CUDA_OK(ctx, cudaFuncSetAttribute(ccl_runs_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); attr_set = true; }
        RT_CUDA_OK(ctx, ctx->d_runs.ensure(sizeof(RunRec) * 3 * RUN_CAP * (size_t)n, st));
        RT_LAUNCH_BEGIN(ctx, "zero_counters_kernel");
        zero_counters_kernel<<<(n + 127) / 128, 128, 0, st>>>(d_cnt, n);
        RT_LAUNCH_CHECK(ctx);
        static const bool use_br2 = getenv("RETTO_B200_BR2") != nullptr;   // A/B: the 4-px-per-lane kernel
        if (use_br2) {
            const int tgrid = (total_tiles + 3) / 4;
            RT_LAUNCH_BEGIN(ctx, "bitmap_runs2_kernel");
            if (vec)
                bitmap_runs2_kernel<true><<<tgrid, 128, 0, st>>>(d_pages, d_tile_prefix, n, total_tiles, ctx->cfg.det_thresh, ctx->cfg.det_dilation_2x2, d_bm, ctx->d_runs.as<RunRec>(), d_cnt);
            else
                bitmap_runs2_kernel<false><<<tgrid, 128, 0, st>>>(d_pages, d_tile_prefix, n, total_tiles, ctx->cfg.det_thresh, ctx->cfg.det_dilation_2x2, d_bm, ctx->d_runs.as<RunRec>(), d_cnt);
            RT_LAUNCH_CHECK(ctx);
        } else {
            // NW = 4: 128-px strips (the tile grid of the pixel path), NW = 8: 256-px strips; rows are stored with 4- / 8-byte words
            // when every page width allows it
            static const int nw = getenv("RETTO_B200_BR3_NW") ? atoi(getenv("RETTO_B200_BR3_NW")) : 8;   // measured: 0.43 ms (8) / 0.58 ms (4) / 0.49 ms (bitmap_runs2) per 256 pages 1280x1280
            const bool a4 = vec || !(R.w_or & 3);
#define BR3_ARGS(pre, tot) d_pages, pre, n, tot, ctx->cfg.det_thresh, ctx->cfg.det_dilation_2x2, d_bm, ctx->d_runs.as<RunRec>(), d_cnt
            RT_LAUNCH_BEGIN(ctx, "bitmap_runs3_kernel");
            if (nw == 8) {   // 62 registers, 8 blocks per SM; capping the registers for 10 / 12 blocks spills and is slower (0.59 / 0.92 ms)
                const int tgrid = (R.total_tiles2 + 3) / 4;
                if (getenv("RETTO_B200_NO_NF_PROBE") && R.a8) bitmap_runs3_kernel<true, 8, 8, false><<<tgrid, 128, 0, st>>>(BR3_ARGS(d_tile2_prefix, R.total_tiles2));   // A/B only
                else if (R.a8) bitmap_runs3_kernel<true, 8, 8><<<tgrid, 128, 0, st>>>(BR3_ARGS(d_tile2_prefix, R.total_tiles2));
                else bitmap_runs3_kernel<false, 8, 8><<<tgrid, 128, 0, st>>>(BR3_ARGS(d_tile2_prefix, R.total_tiles2));
            } else {
                const int tgrid = (total_tiles + 3) / 4;
                if (a4) bitmap_runs3_kernel<true, 4, 12><<<tgrid, 128, 0, st>>>(BR3_ARGS(d_tile_prefix, total_tiles));
                else bitmap_runs3_kernel<false, 4, 12><<<tgrid, 128, 0, st>>>(BR3_ARGS(d_tile_prefix, total_tiles));
            }
#undef BR3_ARGS
            RT_LAUNCH_CHECK(ctx);
        }
        RT_LAUNCH_BEGIN(ctx, "ccl_runs_kernel");
        ccl_runs_kernel<<<n, 1024, smem, st>>>(d_pages, d_cnt, ctx->d_runs.as<RunRec>(), d_comps, d_rowtab, max_comps);
        RT_LAUNCH_CHECK(ctx);
        RT_CUDA_OK(ctx, cudaMemcpyAsync(ctx->h_dp.as<PageCounters>(), d_cnt, sizeof(PageCounters) * n, cudaMemcpyDeviceToHost, st));
        RT_CUDA_OK(ctx, cudaEventRecord(ctx->ev_dp, st));
        return RETTO_B200_OK;
    }
    return dp_pixel_ccl(ctx);
}

retto_b200_status rt_det_post_mid(retto_b200_ctx* ctx) {
    retto_b200_ctx::DpRun& R = ctx->dp;
    const int n = R.n;
    if (n == 0) return RETTO_B200_OK;
    const int max_comps = ctx->cfg.max_components_per_page;
    const int cap = R.cap;
    cudaStream_t st = ctx->stream;
    const DetPostPage* d_pages = ctx->d_dp_pages.as<DetPostPage>();
    PageCounters* d_cnt = ctx->d_dp_counters.as<PageCounters>();
    int* d_offsets = reinterpret_cast<int*>(ctx->d_dp_counters.as<char>() + sizeof(PageCounters) * n);
    unsigned char* d_bm = ctx->d_bitmap.as<unsigned char>();
    int* d_lab = ctx->d_labels.as<int>();
    int* d_cid = ctx->d_cid_at.as<int>();
    CompRec* d_comps = ctx->d_comps.as<CompRec>();
    int2* d_rowtab = ctx->d_rowtab.as<int2>();
    int2* d_hull = d_rowtab + (size_t)n * ROWCAP;
    BoxCand* d_cand = ctx->d_cand.as<BoxCand>();
    PageCounters* h_cnt0 = ctx->h_dp.as<PageCounters>();
    PageCounters* h_cnt = reinterpret_cast<PageCounters*>(ctx->h_dp.as<char>() + R.hdr_bytes);
    retto_b200_box* h_stage_boxes = reinterpret_cast<retto_b200_box*>(ctx->h_dp.as<char>() + 2 * R.hdr_bytes);
    (void)d_bm; (void)d_lab; (void)d_cid; (void)d_comps; (void)d_rowtab; (void)d_hull; (void)d_cand; (void)h_cnt0; (void)h_cnt; (void)h_stage_boxes; (void)d_offsets; (void)max_comps; (void)cap; (void)st; (void)d_pages; (void)d_cnt;
    RT_CUDA_OK(ctx, cudaEventSynchronize(ctx->ev_dp));
    if (ctx->dp_run_path) {
        bool fb = false;
        for (int i = 0; i < n; ++i) fb |= h_cnt0[i].fallback != 0;
        if (fb) {   // a page has more runs than ccl_runs_kernel holds on chip: redo the batch on the pixel planes
            RT_TRY(dp_pixel_ccl(ctx));
            RT_CUDA_OK(ctx, cudaEventSynchronize(ctx->ev_dp));
        }
    }
    int max_n = 0;
    long long box_bound = 0;   // every box comes from one outer or one hole border
    for (int i = 0; i < n; ++i) {
        const int nr = std::min(h_cnt0[i].n_roots, max_comps);
        max_n = std::max(max_n, nr);
        box_bound += nr + std::min(std::max(h_cnt0[i].n_roots - h_cnt0[i].euler, 0), MAX_HOLES);
    }
    bool any_nonfinite = false;
    for (int i = 0; i < n; ++i) any_nonfinite |= h_cnt0[i].nonfinite != 0;
    if (max_n > 0) {
        GeomParams gp{ctx->cfg.det_box_thresh, ctx->cfg.det_unclip_ratio, ctx->cfg.det_min_mini_box_size};
        dim3 grid((max_n + 3) / 4, n);
        // score lists: every component can need a score, so their capacity is the batch's component count
        long long n_comp_total = 0;
        for (int i = 0; i < n; ++i) n_comp_total += std::min(h_cnt0[i].n_roots, max_comps);
        const int list_cap = (int)std::min<long long>(n_comp_total, 0x3fffffff);
        RT_CUDA_OK(ctx, ctx->d_biglist.ensure(sizeof(int) * (4 + 6 * (size_t)list_cap), st));
        int* d_lists = ctx->d_biglist.as<int>();
        RT_CUDA_OK(ctx, cudaMemsetAsync(d_lists, 0, 16, st));
        RT_LAUNCH_BEGIN(ctx, "box_geometry_kernel<0:rect1>");
        box_geometry_kernel<0><<<grid, 128, 0, st>>>(d_pages, n, d_cnt, d_comps, d_rowtab, d_hull, d_cand, max_comps, gp, nullptr, d_lists, list_cap);
        RT_LAUNCH_CHECK(ctx);
        // score: one warp per box, the huge boxes first (RETTO_B200_SCORE_WPB=1: component order, the previous kernel — A/B)
        static const bool score_wpb = getenv("RETTO_B200_SCORE_WPB") != nullptr;
        auto launch_score = [&]() -> retto_b200_status {
            if (score_wpb) {
                RT_LAUNCH_BEGIN(ctx, "box_geometry_kernel<1:score>");
                box_geometry_kernel<1><<<grid, 128, 0, st>>>(d_pages, n, d_cnt, d_comps, d_rowtab, d_hull, d_cand, max_comps, gp, nullptr);
                RT_LAUNCH_CHECK(ctx);
                return RETTO_B200_OK;
            }
            RT_LAUNCH_BEGIN(ctx, "box_score_kernel");
            box_score_kernel<<<SCORE_HUGE_BLOCKS + std::min((list_cap + 3) / 4, 148 * 10 * 2), 128, 0, st>>>(d_pages, d_cnt, d_cand, max_comps, gp, d_lists, list_cap);
            RT_LAUNCH_CHECK(ctx);
            return RETTO_B200_OK;
        };
        const bool concurrent = getenv("RETTO_B200_GEOM_CONCURRENT") != nullptr;   // opt-in: no gain on 256 text pages (4.89 vs 4.90 ms), -5 % on config 2 (40 k boxes)
        if (!concurrent) {
            RT_TRY(launch_score());
            if (any_nonfinite) {
                RT_LAUNCH_BEGIN(ctx, "box_geometry_kernel<4:score,nonfinite>");
                box_geometry_kernel<4><<<grid, 128, 0, st>>>(d_pages, n, d_cnt, d_comps, d_rowtab, d_hull, d_cand, max_comps, gp, nullptr);
                RT_LAUNCH_CHECK(ctx);
            }
            RT_LAUNCH_BEGIN(ctx, "box_geometry_kernel<2:unclip>");
            box_geometry_kernel<2><<<grid, 128, 0, st>>>(d_pages, n, d_cnt, d_comps, d_rowtab, d_hull, d_cand, max_comps, gp, nullptr);
            RT_LAUNCH_CHECK(ctx);
        } else {
            // score (main stream) || unclip of every box that reached the score stage (auxiliary stream), then the merge
            if (!ctx->aux_stream) {
                RT_CUDA_OK(ctx, cudaStreamCreateWithFlags(&ctx->aux_stream, cudaStreamNonBlocking));
                RT_CUDA_OK(ctx, cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming));
                RT_CUDA_OK(ctx, cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming));
            }
            RT_CUDA_OK(ctx, cudaEventRecord(ctx->ev_fork, st));
            RT_CUDA_OK(ctx, cudaStreamWaitEvent(ctx->aux_stream, ctx->ev_fork, 0));
            RT_TRY(launch_score());
            if (any_nonfinite) {
                RT_LAUNCH_BEGIN(ctx, "box_geometry_kernel<4:score,nonfinite>");
                box_geometry_kernel<4><<<grid, 128, 0, st>>>(d_pages, n, d_cnt, d_comps, d_rowtab, d_hull, d_cand, max_comps, gp, nullptr);
                RT_LAUNCH_CHECK(ctx);
            }
            ctx->timer_stream = ctx->aux_stream;
            RT_LAUNCH_BEGIN(ctx, "box_geometry_kernel<3:unclip||score>");
            box_geometry_kernel<3><<<grid, 128, 0, ctx->aux_stream>>>(d_pages, n, d_cnt, d_comps, d_rowtab, d_hull, d_cand, max_comps, gp, nullptr);
            ctx->launches++;
            if (ctx->timing_enabled) ctx->timer_end();
            ctx->timer_stream = nullptr;
            if (cudaGetLastError() != cudaSuccess) { ctx->set_error("kernel launch: box_geometry_kernel<3>"); return RETTO_B200_ERR_CUDA; }
            RT_CUDA_OK(ctx, cudaEventRecord(ctx->ev_join, ctx->aux_stream));
            RT_CUDA_OK(ctx, cudaStreamWaitEvent(st, ctx->ev_join, 0));
            RT_LAUNCH_BEGIN(ctx, "geom_merge_kernel");
            geom_merge_kernel<<<dim3((max_n + 127) / 128, n), 128, 0, st>>>(n, d_cnt, d_cand, max_comps);
            RT_LAUNCH_CHECK(ctx);
        }
    }
    // hole borders: pages with #components - Euler number > 0
    {
        std::vector<int> fp, fpre{0};
        for (int i = 0; i < n; ++i)
            if (h_cnt0[i].status == RETTO_B200_OK && h_cnt0[i].n_roots <= max_comps && h_cnt0[i].n_roots - h_cnt0[i].euler > 0) {
                fp.push_back(i);
                fpre.push_back(fpre.back() + ctx->dp_pages[i].h * ctx->dp_pages[i].w);
            }
        if (!fp.empty()) {
            const int nf = (int)fp.size(), total = fpre.back();
            std::vector<int> blob(fp);
            blob.insert(blob.end(), fpre.begin(), fpre.end());
            RT_TRY(rt_upload(ctx, ctx->d_hole_pages, blob.data(), blob.size() * sizeof(int)));
            RT_CUDA_OK(ctx, ctx->d_holes.ensure(sizeof(int) * (size_t)n * MAX_HOLES, st));
            HoleArgs ha{d_pages, ctx->d_hole_pages.as<int>(), ctx->d_hole_pages.as<int>() + nf, nf, total};
            int* d_holes = ctx->d_holes.as<int>();
            const int g = (total + 255) / 256;
            RT_LAUNCH_BEGIN(ctx, "bg_init_kernel");
            bg_init_kernel<<<g, 256, 0, st>>>(ha, d_bm, d_cid, ctx->dp_run_path ? d_lab : nullptr);
            RT_LAUNCH_CHECK(ctx);
            RT_LAUNCH_BEGIN(ctx, "bg_merge_kernel");
            bg_merge_kernel<<<g, 256, 0, st>>>(ha, d_bm, d_cid);
            RT_LAUNCH_CHECK(ctx);
            RT_LAUNCH_BEGIN(ctx, "bg_flatten_kernel");
            bg_flatten_kernel<<<g, 256, 0, st>>>(ha, d_bm, d_cid, d_lab);
            RT_LAUNCH_CHECK(ctx);
            {   // discovery keys of the outer borders of these pages: only run starts / ends that face frame-connected background count
                int max_r = 1;
                for (int i : fp) max_r = std::max(max_r, std::min(h_cnt0[i].n_roots, max_comps));
                const dim3 kg((max_r + 255) / 256, nf);
                int* d_key = ctx->d_key_at.as<int>();
                RT_LAUNCH_BEGIN(ctx, "outer_key_init_kernel");
                outer_key_init_kernel<<<kg, 256, 0, st>>>(ha, d_cnt, d_comps, max_comps, d_key);
                RT_LAUNCH_CHECK(ctx);
                RT_LAUNCH_BEGIN(ctx, "outer_key_scan_kernel");
                outer_key_scan_kernel<<<g, 256, 0, st>>>(ha, d_bm, d_cid, d_lab, d_cnt, ctx->dp_run_path ? 1 : 0,
                                                        ctx->dp_run_path ? ctx->d_runs.as<RunRec>() : nullptr, d_key);
                RT_LAUNCH_CHECK(ctx);
                RT_LAUNCH_BEGIN(ctx, "outer_key_apply_kernel");
                outer_key_apply_kernel<<<kg, 256, 0, st>>>(ha, d_cnt, d_comps, max_comps, d_key, d_cand);
                RT_LAUNCH_CHECK(ctx);
            }
            RT_LAUNCH_BEGIN(ctx, "hole_collect_kernel");
            hole_collect_kernel<<<g, 256, 0, st>>>(ha, d_bm, d_cid, d_lab, d_cnt, d_holes);
            RT_LAUNCH_CHECK(ctx);
            RT_LAUNCH_BEGIN(ctx, "hole_sort_kernel");
            hole_sort_kernel<<<nf, 1024, 0, st>>>(ha, d_cnt, d_holes, d_lab, d_comps, max_comps);
            RT_LAUNCH_CHECK(ctx);
            RT_LAUNCH_BEGIN(ctx, "hole_extent_kernel");
            hole_extent_kernel<<<g, 256, 0, st>>>(ha, d_bm, d_cid, d_lab, d_cnt, d_comps, max_comps);
            RT_LAUNCH_CHECK(ctx);
            RT_LAUNCH_BEGIN(ctx, "hole_row_alloc_kernel");
            hole_row_alloc_kernel<<<nf, 256, 0, st>>>(ha, d_cnt, d_comps, d_rowtab, max_comps);
            RT_LAUNCH_CHECK(ctx);
            RT_LAUNCH_BEGIN(ctx, "hole_rows_kernel");
            hole_rows_kernel<<<g, 256, 0, st>>>(ha, d_bm, d_cid, d_lab, d_cnt, d_comps, d_rowtab, max_comps);
            RT_LAUNCH_CHECK(ctx);
            int max_h = 0;
            for (int i : fp) max_h = std::max(max_h, std::min(h_cnt0[i].n_roots - h_cnt0[i].euler, MAX_HOLES));
            GeomParams gp{ctx->cfg.det_box_thresh, ctx->cfg.det_unclip_ratio, ctx->cfg.det_min_mini_box_size};
            const dim3 hg((max_h + 3) / 4, nf);
            RT_LAUNCH_BEGIN(ctx, "box_geometry_kernel(holes)");
            box_geometry_kernel<0><<<hg, 128, 0, st>>>(d_pages, n, d_cnt, d_comps, d_rowtab, d_hull, d_cand, max_comps, gp, ha.flag_pages);
            RT_LAUNCH_CHECK(ctx);
            RT_LAUNCH_BEGIN(ctx, "box_geometry_kernel(holes)");
            box_geometry_kernel<1><<<hg, 128, 0, st>>>(d_pages, n, d_cnt, d_comps, d_rowtab, d_hull, d_cand, max_comps, gp, ha.flag_pages);
            RT_LAUNCH_CHECK(ctx);
            if (any_nonfinite) {
                RT_LAUNCH_BEGIN(ctx, "box_geometry_kernel(holes)");
                box_geometry_kernel<4><<<hg, 128, 0, st>>>(d_pages, n, d_cnt, d_comps, d_rowtab, d_hull, d_cand, max_comps, gp, ha.flag_pages);
                RT_LAUNCH_CHECK(ctx);
            }
            RT_LAUNCH_BEGIN(ctx, "box_geometry_kernel(holes)");
            box_geometry_kernel<2><<<hg, 128, 0, st>>>(d_pages, n, d_cnt, d_comps, d_rowtab, d_hull, d_cand, max_comps, gp, ha.flag_pages);
            RT_LAUNCH_CHECK(ctx);
            RT_LAUNCH_BEGIN(ctx, "hole_restore_kernel");
            hole_restore_kernel<<<g, 256, 0, st>>>(ha, d_lab);
            RT_LAUNCH_CHECK(ctx);
            max_n = std::max(max_n, std::min(max_n + max_h, max_comps));
        }
    }
    ctx->dp_trace_valid = false;
    if (ctx->dp_trace_enabled && max_n > 0) {
        RT_CUDA_OK(ctx, ctx->d_trace.ensure(sizeof(TraceRec) * (size_t)n * max_comps, st));
        RT_LAUNCH_BEGIN(ctx, "trace_copy_kernel");
        trace_copy_kernel<<<dim3((max_n + 127) / 128, n), 128, 0, st>>>(n, d_cnt, d_cand, max_comps, ctx->d_trace.as<TraceRec>());
        RT_LAUNCH_CHECK(ctx);
        ctx->dp_trace_valid = true;
    }
    RT_LAUNCH_BEGIN(ctx, "page_sort_kernel");
    RT_CUDA_OK(ctx, ctx->d_order.ensure(sizeof(int) * (size_t)n * max_comps, st));
    page_sort_kernel<<<n, 32, 0, st>>>(n, d_cnt, d_cand, max_comps, ctx->d_order.as<int>());
    RT_LAUNCH_CHECK(ctx);
    RT_LAUNCH_BEGIN(ctx, "pack_offsets_kernel");
    pack_offsets_kernel<<<1, 256, 0, st>>>(n, d_cnt, d_offsets);
    RT_LAUNCH_CHECK(ctx);
    RT_CUDA_OK(ctx, ctx->d_boxes_out.ensure(sizeof(retto_b200_box) * (size_t)std::max(cap, 1), st));
    RT_LAUNCH_BEGIN(ctx, "pack_boxes_kernel");
    pack_boxes_kernel<<<n, 128, 0, st>>>(n, d_cnt, d_offsets, d_cand, max_comps, ctx->d_boxes_out.as<retto_b200_box>(), cap, ctx->d_order.as<int>());
    RT_LAUNCH_CHECK(ctx);
    // one read-back, one sync: counters, offsets and the boxes (as many as the early counts allow for) into pinned memory
    int* h_off = reinterpret_cast<int*>(reinterpret_cast<char*>(h_cnt) + sizeof(PageCounters) * n);
    const int nspec = (int)std::min<long long>(box_bound, cap);
    RT_CUDA_OK(ctx, cudaMemcpyAsync(h_cnt, d_cnt, sizeof(PageCounters) * n + sizeof(int) * (n + 1), cudaMemcpyDeviceToHost, st));
    if (nspec > 0) RT_CUDA_OK(ctx, cudaMemcpyAsync(h_stage_boxes, ctx->d_boxes_out.p, sizeof(retto_b200_box) * (size_t)nspec, cudaMemcpyDeviceToHost, st));
    R.nspec = nspec;
    // rt_det_post_end waits for THIS point, not for the stream: the session enqueues the crop kernels behind it
    if (!ctx->ev_dp2) RT_CUDA_OK(ctx, cudaEventCreateWithFlags(&ctx->ev_dp2, cudaEventDisableTiming));
    RT_CUDA_OK(ctx, cudaEventRecord(ctx->ev_dp2, st));
    return RETTO_B200_OK;
}

retto_b200_status rt_det_post_end(retto_b200_ctx* ctx, int32_t* h_page_status, int32_t* h_box_offsets, retto_b200_box* h_boxes) {
    h_box_offsets[0] = 0;
    retto_b200_ctx::DpRun& R = ctx->dp;
    const int n = R.n;
    if (n == 0) return RETTO_B200_OK;
    const int max_comps = ctx->cfg.max_components_per_page;
    const int cap = R.cap;
    cudaStream_t st = ctx->stream;
    const DetPostPage* d_pages = ctx->d_dp_pages.as<DetPostPage>();
    PageCounters* d_cnt = ctx->d_dp_counters.as<PageCounters>();
    int* d_offsets = reinterpret_cast<int*>(ctx->d_dp_counters.as<char>() + sizeof(PageCounters) * n);
    unsigned char* d_bm = ctx->d_bitmap.as<unsigned char>();
    int* d_lab = ctx->d_labels.as<int>();
    int* d_cid = ctx->d_cid_at.as<int>();
    CompRec* d_comps = ctx->d_comps.as<CompRec>();
    int2* d_rowtab = ctx->d_rowtab.as<int2>();
    int2* d_hull = d_rowtab + (size_t)n * ROWCAP;
    BoxCand* d_cand = ctx->d_cand.as<BoxCand>();
    PageCounters* h_cnt0 = ctx->h_dp.as<PageCounters>();
    PageCounters* h_cnt = reinterpret_cast<PageCounters*>(ctx->h_dp.as<char>() + R.hdr_bytes);
    retto_b200_box* h_stage_boxes = reinterpret_cast<retto_b200_box*>(ctx->h_dp.as<char>() + 2 * R.hdr_bytes);
    (void)d_bm; (void)d_lab; (void)d_cid; (void)d_comps; (void)d_rowtab; (void)d_hull; (void)d_cand; (void)h_cnt0; (void)h_cnt; (void)h_stage_boxes; (void)d_offsets; (void)max_comps; (void)cap; (void)st; (void)d_pages; (void)d_cnt;
    const int nspec = R.nspec;
    int* h_off = reinterpret_cast<int*>(reinterpret_cast<char*>(h_cnt) + sizeof(PageCounters) * n);
    RT_CUDA_OK(ctx, cudaEventSynchronize(ctx->ev_dp2));
    retto_b200_status ret = RETTO_B200_OK;
    for (int i = 0; i < n; ++i) {
        int s = h_cnt[i].status;
        // hole borders: find_contours would return extra (hole) contours; #holes = #components - euler
        h_page_status[i] = s;
        h_box_offsets[i] = h_off[i];
    }
    h_box_offsets[n] = h_off[n];
    ctx->dp_holes.assign(n, 0);
    ctx->dp_ncomp.assign(n, 0);
    ctx->dp_nruns.assign(n, 0);
    for (int i = 0; i < n; ++i) { ctx->dp_holes[i] = h_cnt[i].n_holes; ctx->dp_ncomp[i] = h_cnt[i].n_roots + h_cnt[i].n_holes; ctx->dp_nruns[i] = h_cnt[i].n_runs; }
    const int total = h_off[n];
    if (total > cap) {
        ctx->set_error("det_postprocess: " + std::to_string(total) + " boxes exceed max_boxes_total " + std::to_string(cap));
        ret = RETTO_B200_ERR_CAPACITY;
    }
    const int ncopy = std::min(total, cap);
    if (ncopy > 0 && h_boxes) {
        if (ncopy > nspec) {   // cannot happen (box_bound is an upper bound); kept as a safe path
            RT_CUDA_OK(ctx, cudaMemcpyAsync(h_stage_boxes, ctx->d_boxes_out.p, sizeof(retto_b200_box) * (size_t)ncopy, cudaMemcpyDeviceToHost, st));
            RT_CUDA_OK(ctx, cudaStreamSynchronize(st));
        }
        memcpy(h_boxes, h_stage_boxes, sizeof(retto_b200_box) * (size_t)ncopy);
    }
    return ret;
}

extern "C" retto_b200_status retto_b200_det_postprocess(retto_b200_ctx* ctx, const retto_b200_det_post_desc* h_descs, int32_t n,
                                                        int32_t* h_page_status, int32_t* h_box_offsets, retto_b200_box* h_boxes,
                                                        int32_t max_boxes_total) {
    RtDeviceGuard _dg(ctx);
    if (!ctx || n < 0 || (n > 0 && (!h_descs || !h_page_status)) || !h_box_offsets) return RETTO_B200_ERR_INVALID_ARG;
    RT_TRY(rt_det_post_begin(ctx, h_descs, n, max_boxes_total));
    RT_TRY(rt_det_post_mid(ctx));
    return rt_det_post_end(ctx, h_page_status, h_box_offsets, h_boxes);
}

extern "C" retto_b200_status retto_b200_det_post_fetch_bitmap(retto_b200_ctx* ctx, int32_t page, uint8_t* h_bitmap) {
    RtDeviceGuard _dg(ctx);
    if (!ctx || page < 0 || page >= (int)ctx->dp_pages.size() || !h_bitmap) return RETTO_B200_ERR_INVALID_ARG;
    const DetPostPage& pg = ctx->dp_pages[page];
    RT_CUDA_OK(ctx, cudaMemcpyAsync(h_bitmap, ctx->d_bitmap.as<unsigned char>() + pg.px_base, (size_t)pg.h * pg.w, cudaMemcpyDeviceToHost, ctx->stream));
    RT_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
    return RETTO_B200_OK;
}
extern "C" retto_b200_status retto_b200_det_post_fetch_labels(retto_b200_ctx* ctx, int32_t page, int32_t* h_labels) {
    RtDeviceGuard _dg(ctx);
    if (!ctx || page < 0 || page >= (int)ctx->dp_pages.size() || !h_labels) return RETTO_B200_ERR_INVALID_ARG;
    const DetPostPage& pg = ctx->dp_pages[page];
    if (!ctx->dp_labels_final[page] && ctx->dp_run_path) {
        // run-table path: the plane is materialised from the page's sorted run table
        int* plane = ctx->d_labels.as<int>() + pg.px_base;
        RT_CUDA_OK(ctx, cudaMemsetAsync(plane, 0xff, sizeof(int) * (size_t)pg.h * pg.w, ctx->stream));
        const int nr = page < (int)ctx->dp_nruns.size() ? ctx->dp_nruns[page] : 0;
        if (nr > 0) {
            const RunRec* sorted = ctx->d_runs.as<RunRec>() + (size_t)page * 3 * RUN_CAP + RUN_CAP;
            const int* lab = reinterpret_cast<const int*>(ctx->d_runs.as<RunRec>() + (size_t)page * 3 * RUN_CAP + 2 * RUN_CAP);
            RT_LAUNCH_BEGIN(ctx, "labels_from_runs_kernel");
            labels_from_runs_kernel<<<(nr + 7) / 8, 256, 0, ctx->stream>>>(sorted, lab, nr, pg.w, plane);
            RT_LAUNCH_CHECK(ctx);
        }
        ctx->dp_labels_final[page] = 1;
    }
    if (!ctx->dp_labels_final[page]) {
        RT_LAUNCH_BEGIN(ctx, "labels_finalize_kernel");
        labels_finalize_kernel<<<(pg.h * pg.w + 255) / 256, 256, 0, ctx->stream>>>(ctx->d_bitmap.as<unsigned char>(), ctx->d_labels.as<int>(), pg.px_base, pg.h * pg.w);
        RT_LAUNCH_CHECK(ctx);
        ctx->dp_labels_final[page] = 1;
    }
    RT_CUDA_OK(ctx, cudaMemcpyAsync(h_labels, ctx->d_labels.as<int>() + pg.px_base, sizeof(int) * (size_t)pg.h * pg.w, cudaMemcpyDeviceToHost, ctx->stream));
    RT_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
    return RETTO_B200_OK;
}

// PointBox::scale_and_clip (points.rs:179-194) for host-resident result boxes (session.rs:94-97): 8 numbers per box of
// f64 rounding, done in a tiny kernel so that no arithmetic of the path runs on the CPU.
__global__ void scale_clip_kernel(retto_b200_box* b, int n, double inv_w, double inv_h, double ori_w, double ori_h) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        b[i].xy[2 * k] = scale_clip_1(b[i].xy[2 * k], inv_w, ori_w);
        b[i].xy[2 * k + 1] = scale_clip_1(b[i].xy[2 * k + 1], inv_h, ori_h);
    }
}
extern "C" retto_b200_status retto_b200_scale_and_clip(retto_b200_ctx* ctx, retto_b200_box* h_boxes, int32_t n, double bitmap_w,
                                                       double bitmap_h, double ori_w, double ori_h) {
    RtDeviceGuard _dg(ctx);
    if (!ctx || n < 0 || (n > 0 && !h_boxes)) return RETTO_B200_ERR_INVALID_ARG;
    if (n == 0) return RETTO_B200_OK;
    RT_TRY(rt_upload(ctx, ctx->d_stage3, h_boxes, sizeof(retto_b200_box) * (size_t)n));
    RT_LAUNCH_BEGIN(ctx, "scale_clip_kernel");
    scale_clip_kernel<<<(n + 127) / 128, 128, 0, ctx->stream>>>(ctx->d_stage3.as<retto_b200_box>(), n, ori_w / bitmap_w, ori_h / bitmap_h, ori_w, ori_h);
    RT_LAUNCH_CHECK(ctx);
    RT_CUDA_OK(ctx, cudaMemcpyAsync(h_boxes, ctx->d_stage3.p, sizeof(retto_b200_box) * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
    RT_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
    return RETTO_B200_OK;
}

// parity tap: per-component trace of the last det_postprocess (enable with retto_b200_det_post_enable_trace)
extern "C" retto_b200_status retto_b200_det_post_enable_trace(retto_b200_ctx* ctx, int32_t on) {
    if (!ctx) return RETTO_B200_ERR_INVALID_ARG;
    ctx->dp_trace_enabled = on != 0;
    return RETTO_B200_OK;
}
extern "C" retto_b200_status retto_b200_det_post_fetch_trace(retto_b200_ctx* ctx, int32_t page, int32_t* n_components, int32_t* n_holes,
                                                             int32_t* h_key, int32_t* h_status, int32_t* h_rect1, float* h_sside1,
                                                             float* h_score, int32_t max_components) {
    RtDeviceGuard _dg(ctx);
    if (!ctx || page < 0 || page >= (int)ctx->dp_pages.size() || !n_components) return RETTO_B200_ERR_INVALID_ARG;
    const int n = ctx->dp_ncomp[page];
    *n_components = n;
    if (n_holes) *n_holes = ctx->dp_holes[page];
    if (!h_key) return RETTO_B200_OK;
    if (!ctx->dp_trace_valid && n > 0) { ctx->set_error("trace not enabled for the last det_postprocess"); return RETTO_B200_ERR_INVALID_ARG; }
    if (n > max_components) return RETTO_B200_ERR_CAPACITY;
    std::vector<TraceRec> t(n);
    if (n > 0) {
        RT_CUDA_OK(ctx, cudaMemcpyAsync(t.data(), ctx->d_trace.as<TraceRec>() + (size_t)page * ctx->cfg.max_components_per_page,
                                        sizeof(TraceRec) * n, cudaMemcpyDeviceToHost, ctx->stream));
        RT_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
    }
    for (int i = 0; i < n; ++i) {
        h_key[i] = t[i].key; h_status[i] = t[i].status; h_sside1[i] = t[i].sside1; h_score[i] = t[i].score;
        for (int k = 0; k < 8; ++k) h_rect1[8 * i + k] = t[i].rect1[k];
    }
    return RETTO_B200_OK;
}


// batched variant used by the session: per-box (inv_w, inv_h, ori_w, ori_h), one launch, one sync
__global__ void scale_clip_multi_kernel(retto_b200_box* b, const double4* __restrict__ prm, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double4 p = prm[i];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        b[i].xy[2 * k] = scale_clip_1(b[i].xy[2 * k], p.x, p.z);
        b[i].xy[2 * k + 1] = scale_clip_1(b[i].xy[2 * k + 1], p.y, p.w);
    }
}
// defer == true: the read-back lands in ctx->h_scale (pinned) and the caller copies it out after its next stream sync
retto_b200_status rt_scale_and_clip_multi(retto_b200_ctx* ctx, retto_b200_box* h_boxes, const double* h_params4, int n, bool defer) {
    if (n == 0) return RETTO_B200_OK;
    const size_t bb = (sizeof(retto_b200_box) * (size_t)n + 31) & ~size_t(31);
    std::vector<char> blob(bb + sizeof(double) * 4 * (size_t)n);
    memcpy(blob.data(), h_boxes, sizeof(retto_b200_box) * (size_t)n);
    memcpy(blob.data() + bb, h_params4, sizeof(double) * 4 * (size_t)n);
    RT_TRY(rt_upload(ctx, ctx->d_stage3, blob.data(), blob.size()));
    RT_LAUNCH_BEGIN(ctx, "scale_clip_multi_kernel");
    scale_clip_multi_kernel<<<(n + 127) / 128, 128, 0, ctx->stream>>>(ctx->d_stage3.as<retto_b200_box>(),
                                                                      reinterpret_cast<const double4*>(ctx->d_stage3.as<char>() + bb), n);
    RT_LAUNCH_CHECK(ctx);
    if (defer) {
        RT_CUDA_OK(ctx, ctx->h_scale.ensure(sizeof(retto_b200_box) * (size_t)n));
        RT_CUDA_OK(ctx, cudaMemcpyAsync(ctx->h_scale.p, ctx->d_stage3.p, sizeof(retto_b200_box) * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
        return RETTO_B200_OK;
    }
    RT_CUDA_OK(ctx, cudaMemcpyAsync(h_boxes, ctx->d_stage3.p, sizeof(retto_b200_box) * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
    RT_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
    return RETTO_B200_OK;
}
