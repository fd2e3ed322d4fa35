// common.cuh — context, buffers and launch helpers shared by the kernels of libretto_b200.so
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <cmath>
#include <string>
#include <vector>

#include "../../include/retto_b200.h"

#define RT_CUDA_OK(ctx, expr)                                                                    \
    do {                                                                                         \
        cudaError_t _e = (expr);                                                                 \
        if (_e != cudaSuccess) {                                                                 \
            (ctx)->set_error(std::string(#expr) + ": " + cudaGetErrorString(_e));                \
            return RETTO_B200_ERR_CUDA;                                                          \
        }                                                                                        \
    } while (0)

#define RT_TRY(expr)                                   \
    do {                                               \
        retto_b200_status _s = (expr);                 \
        if (_s != RETTO_B200_OK) return _s;            \
    } while (0)

// growable device buffer (never shrinks; growth synchronises the stream)
struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    DevBuf() = default;
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    ~DevBuf() { release(); }
    cudaError_t ensure(size_t bytes, cudaStream_t s) {
        if (bytes <= cap) return cudaSuccess;
        cudaError_t e = cudaStreamSynchronize(s);
        if (e != cudaSuccess) return e;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        size_t want = bytes + bytes / 4 + 256;
        e = cudaMalloc(&p, want);
        if (e != cudaSuccess) {
            e = cudaMalloc(&p, bytes);
            want = bytes;
        }
        if (e == cudaSuccess) cap = want;
        return e;
    }
    template <class T> T* as() const { return reinterpret_cast<T*>(p); }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};
struct HostBuf {  // pinned
    void* p = nullptr;
    size_t cap = 0;
    HostBuf() = default;
    HostBuf(const HostBuf&) = delete;
    HostBuf& operator=(const HostBuf&) = delete;
    ~HostBuf() { release(); }
    cudaError_t ensure(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
        size_t want = bytes + bytes / 4 + 256;
        cudaError_t e = cudaHostAlloc(&p, want, cudaHostAllocDefault);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    template <class T> T* as() const { return reinterpret_cast<T*>(p); }
    void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
};

// ---- device-side descriptors ------------------------------------------------------------------
struct DetPostPage {       // one page of a det_postprocess batch
    const float* prob;
    int h, w, ori_h, ori_w;
    int strips, rowblocks;   // tile grid: 128 px x 16 rows
    int tile_base;           // first tile index of this page in the flat tile list
    long long px_base;       // first pixel of this page in the flat bitmap / label arrays
    int comp_base;           // first slot of this page in the component tables
    int box_base;            // first slot in the candidate box table
};

struct PageCounters {      // per page, zeroed before each det_postprocess
    int n_roots;
    int euler;               // #components - #holes of the dilated bitmap
    int n_boxes;
    int status;
    int row_total;           // rows allocated in the row-extreme table
    int n_holes;             // hole borders (background components not connected to the frame)
    int n_runs;              // run-table CCL: horizontal foreground runs emitted by bitmap_runs2_kernel
    int fallback;            // run-table CCL: more runs than the on-chip table holds -> the batch is redone on the pixel path
    int nonfinite;           // the probability map holds NaN / +-Inf: box_score_fast takes the exact (whole-bbox, v * m) fold
    int pad;
};

struct CompRec {           // per connected component (dense id)
    int root;                // min linear index (page-local)
    int ymax, xmin, xmax;
    int row_off;             // offset into the row-extreme table (page-relative)
    int key;                 // raster index at which imageproc's scan discovers the border (INT_MAX: never)
    int ymin;                // first row of the contour's point set
    int pad;
};

struct CropDev {
    const uint8_t* page; int page_h, page_w;
    float box[8];
    float t[9]; int cls;     // inverse projection (crop -> page) and its class
    int w, h;                // final crop dims (after rotate270)
    int rot;
    int status;
    unsigned long long offset;
    int direct;              // the crop is the page rectangle at (tx, ty): translation by whole pixels, not rotated, every bicubic window inside the page
    int tx, ty;
    int pad;
};

struct LineDev {
    int crop, img_w, resized_w, kind;   // kind: BB_FF / BB_BB / BB_BF / BB_GEN (rec_batch.cu)
    unsigned long long dst_offset;
};

struct LogitsRow {         // one (line) entry of a CTC batch
    const float* logits; int t; int out_base;  // out_base: first (line, t) slot in idx/prob arrays
};

struct JpegInfo {          // host: the parsed markers of one baseline JPEG file (jpeg_decode.cu)
    int status;
    int X, Y, nc;
    int h[3], v[3], tq[3], td[3], ta[3];
    int max_h, max_v, mcux, mcuy, ri, n_seg;
    size_t ecs_off, ecs_len;   // entropy-coded data: offset in the file, bytes up to the end of the file
    uint16_t qt[4][64];        // natural order
    uint8_t qt_present[4];
    struct Huff { uint8_t present; uint8_t bits[17]; uint8_t vals[256]; } dc[4], ac[4];
};
retto_b200_status rt_jpeg_parse(const uint8_t* d, size_t n, JpegInfo* out);
struct retto_b200_ctx;
retto_b200_status rt_jpeg_entropy_enqueue(retto_b200_ctx* ctx, cudaStream_t st, const JpegInfo* infos, const uint8_t* const* d_bytes, int n);
retto_b200_status rt_jpeg_pixels_enqueue(retto_b200_ctx* owner, retto_b200_ctx* lane, int first, int n, uint8_t* const* d_out);

struct retto_b200_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    retto_b200_config cfg;
    std::string err;
    uint64_t launches = 0;

    void set_error(const std::string& e) { err = e; }

    // optional per-kernel CUDA-event timing (bench.py roofline): events on the launching stream
    struct TimedLaunch { int name_id; cudaEvent_t a, b; };
    bool timing_enabled = false;
    std::vector<std::string> timer_names;
    std::vector<TimedLaunch> timed;
    std::vector<cudaEvent_t> event_pool;
    int timer_pending = -1;
    cudaStream_t timer_stream = nullptr;   // set around a launch on an auxiliary stream
    std::vector<double> timer_total_ms;
    std::vector<uint64_t> timer_count;
    cudaEvent_t get_event() {
        if (!event_pool.empty()) { cudaEvent_t e = event_pool.back(); event_pool.pop_back(); return e; }
        cudaEvent_t e; cudaEventCreate(&e); return e;
    }
    void timer_begin(const char* name) {
        int id = -1;
        for (size_t i = 0; i < timer_names.size(); ++i) if (timer_names[i] == name) { id = (int)i; break; }
        if (id < 0) { id = (int)timer_names.size(); timer_names.push_back(name); timer_total_ms.push_back(0); timer_count.push_back(0); }
        TimedLaunch t{id, get_event(), get_event()};
        cudaEventRecord(t.a, timer_stream ? timer_stream : stream);
        timed.push_back(t);
        timer_pending = (int)timed.size() - 1;
    }
    void timer_end() {
        if (timer_pending < 0) return;
        cudaEventRecord(timed[timer_pending].b, timer_stream ? timer_stream : stream);
        timer_pending = -1;
    }
    void timer_collect() {  // caller synchronises the stream first
        for (auto& t : timed) {
            float ms = 0;
            if (cudaEventElapsedTime(&ms, t.a, t.b) == cudaSuccess) { timer_total_ms[t.name_id] += ms; timer_count[t.name_id]++; }
            event_pool.push_back(t.a); event_pool.push_back(t.b);
        }
        timed.clear();
    }

    // descriptor staging: device side, and pinned host slots for rt_upload (a cudaMemcpyAsync from pageable memory
    // larger than 64 KB waits for the stream to drain, which would serialise the host against the GPU)
    DevBuf d_stage, d_stage2, d_stage3, d_stage_cols, d_stage_tcols;
    struct StageSlot { void* p = nullptr; size_t cap = 0; cudaEvent_t ev = nullptr; bool busy = false; };
    std::vector<StageSlot> stage_slots;
    bool uploads_by_sm = false;   // descriptor uploads through stage_pull_kernel instead of the H2D copy engine (capi.cu)
    HostBuf h_scale;   // deferred scale_and_clip read-back (session.cu)

    // dictionary (rec_processor.rs:29-46)
    std::vector<std::string> dict;
    DevBuf d_dict_bytes, d_dict_offs;
    int dict_max_len = 0;

    // ctc scratch
    DevBuf d_ctc_rows, d_ctc_idx, d_ctc_prob, d_ctc_tok, d_ctc_cnt, d_ctc_score, d_ctc_text, d_ctc_tlen, d_ctc_flag;
    HostBuf h_ctc;
    struct CtcRun { int n_lines = 0, max_t = 0; size_t o_score = 0, o_offs = 0, o_text = 0, o_tok = 0; bool want_tokens = false; } ctc;   // begin -> end state

    // det post state (kept for the fetch_* taps and for crop jobs)
    std::vector<DetPostPage> dp_pages;
    std::vector<int> dp_ncomp;
    std::vector<int> dp_nruns;           // per page: runs in the run table of the last det_postprocess
    std::vector<char> dp_labels_final;   // per page: run-interior labels resolved (lazily, by fetch_labels)
    bool dp_trace_enabled = false, dp_trace_valid = false;
    DevBuf d_trace;
    std::vector<int> dp_holes;   // per page: #hole borders of the last det_postprocess (components - Euler number)
    DevBuf d_dp_pages, d_dp_counters, d_bitmap, d_labels, d_tileflags, d_roots, d_comps, d_cid_at, d_rowtab, d_cand, d_boxes_out, d_holes, d_hole_pages, d_key_at;
    DevBuf d_order;          // per page: permutation of the valid box candidates (page_sort_kernel -> pack_boxes_kernel)
    DevBuf d_biglist;        // det postprocess: [count, pad, (page, id) x SCORE_BIG_CAP] — boxes box_score_kernel scores first
    DevBuf d_runs;           // run-table CCL: per page RUN_CAP raw records + RUN_CAP sorted records (key, x1|flags) + RUN_CAP labels
    bool ccl_runs_attr_set = false;   // cudaFuncSetAttribute(ccl_runs_kernel, max dynamic shared memory) done on this context's device
    bool dp_run_path = false;   // the last det_postprocess used the run-table CCL (labels are materialised lazily from the runs)
    HostBuf h_dp;
    cudaEvent_t ev_dp = nullptr;   // behind the early counter read-back of det_postprocess
    cudaEvent_t ev_dp2 = nullptr;  // behind the box read-back of det_postprocess
    cudaStream_t aux_stream = nullptr;             // box_score_fast || unclip (db_post.cu): fork / join around the two kernels
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    struct DpRun { int n = 0, cap = 0, total_tiles = 0, total_tiles2 = 0, nspec = 0, w_or = 0; size_t hdr_bytes = 0; bool vec = true, a8 = true; } dp;   // begin -> mid -> end state

    // crops
    struct CropHost { int w, h, rot, status; unsigned long long offset; };   // host view of the current crop set
    std::vector<CropHost> crops;
    DevBuf d_crop_descs, d_crop_pix, d_crop_flip, d_crop_pages, d_crop_any;
    int crop_dev_cap = 0;                 // device-built descriptor table (crop.cu): capacities at enqueue time, 0 = not enqueued
    size_t crop_dev_cap_bytes = 0, crop_dev_desc_bytes = 0;
    bool crop_dev_check = false;          // rt_crop_finish compares the device's crop sizes with the host's
    int crops_seen_max = 0, crop_rows_seen_max = 0, crop_dev_row_cap = 0;
    bool crops_lazy = false;              // session path: direct crops (CropDev::direct) are not materialised — the batch build reads the page
    struct CropLaunch { const int* d_prefix = nullptr; int n = 0, rows = 0; const void* d_totals = nullptr; } crop_launch;   // for retto_b200_crop_fetch
    HostBuf h_crops;

    // batches
    DevBuf d_lines, d_lines_rec, d_batch_cls, d_batch_rec;
    std::vector<LineDev> bb_lines;            // build_batches host scratch (prepare -> launch, per kind: 0 cls, 1 rec)
    std::vector<char> bb_blob;
    size_t bb_chunks[2] = {0, 0}, bb_chunk_off[2] = {0, 0};
    DevBuf d_cls_idx, d_cls_out;
    HostBuf h_cls;

    // run_pages pipeline (session.cu): child contexts ("lanes") on the same device, and the tunables
    std::vector<retto_b200_ctx*> lanes;
    int pipe_lanes = 0, pipe_unit_pages = 0;     // 0 = default / environment
    retto_b200_stage_fn stage_cb = nullptr;      // per-stage result delivery (retto_b200_set_stage_callback)
    void* stage_user = nullptr;
    std::string dict_source;                     // raw dictionary text, replayed into the lanes
    uint64_t dict_version = 0;
    // sizes of the last run_pages call (bench.py algorithmic bytes): pages, lines, det px, crop px, cls floats, rec floats, rec rows
    uint64_t run_stats[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    // image decode (jpeg_decode.cu)
    DevBuf d_jpeg_blob, d_jpeg_desc, d_jpeg_seg, d_jpeg_coef, d_jpeg_planes, d_jpeg_clean, d_jpeg_out;
    struct JpegBatch {                           // the batch whose entropy phase ran last: what the pixel phase of its units needs
        int n = 0;
        size_t desc_bytes = 0, head_bytes = 0;
        std::vector<int> X, Y;
        std::vector<unsigned> n_blocks;
        std::vector<unsigned char> is420;   // 3 components, 2x2 luma : 1x1 chroma (the colour kernel has a specialisation for it)
        HostBuf h_desc, h_sub;
    } jpeg;
    DevBuf d_jpeg_sub;                           // K-J2s: slot tables, parse states, MCU-start bitmaps of the sub-sequence decode
    bool jpeg_sub_attr_set = false;
    cudaEvent_t ev_jpeg_zero = nullptr;          // behind the zero-fill of the coefficient planes (issued on `stream` while the copy stream uploads)
    bool jpeg_huff_attr_set = false;             // dynamic shared memory opt-in of jpeg_huff_kernel done on this context's device
    int* jpeg_status_dev = nullptr;              // per-file device status of the last decode (inside d_jpeg_seg)
    HostBuf h_jpeg_status;
    std::vector<JpegInfo> jpeg_infos;            // parsed headers of the encoded pages of the current run_pages call
    // session scratch
    DevBuf d_pages_raw, d_pages_rs, d_det_in, d_pages_up;
    cudaStream_t copy_stream = nullptr;          // H2D of later chunks overlaps the kernels of earlier ones (run_pages)
    std::vector<cudaEvent_t> copy_events;
    std::vector<retto_b200_page_result> r_pages;
    std::vector<retto_b200_box> r_boxes;
    std::vector<retto_b200_cls_result> r_cls;
    std::vector<uint32_t> r_text_offs;
    std::vector<char> r_text;
    std::vector<float> r_scores;
};

// Every extern "C" entry that takes a context makes the context's device current for the duration of the call and
// restores the caller's device on return: a host thread may own contexts on several GPUs (retto_b200_run_pages_multi,
// `--gpus N`), and torch or the host application may change the current device between calls.
struct RtDeviceGuard {
    int prev = -1;
    bool switched = false;
    explicit RtDeviceGuard(const retto_b200_ctx* c) {
        if (!c) return;
        if (cudaGetDevice(&prev) != cudaSuccess) { prev = -1; cudaGetLastError(); }
        if (prev != c->device) switched = cudaSetDevice(c->device) == cudaSuccess;
    }
    ~RtDeviceGuard() { if (switched && prev >= 0) cudaSetDevice(prev); }
    RtDeviceGuard(const RtDeviceGuard&) = delete;
    RtDeviceGuard& operator=(const RtDeviceGuard&) = delete;
};

#define RT_LAUNCH_BEGIN(ctx, name)                                                                \
    do {                                                                                          \
        if ((ctx)->timing_enabled) (ctx)->timer_begin(name);                                      \
    } while (0)

#define RT_LAUNCH_CHECK(ctx)                                                                      \
    do {                                                                                          \
        (ctx)->launches++;                                                                        \
        if ((ctx)->timing_enabled) (ctx)->timer_end();                                            \
        cudaError_t _e = cudaGetLastError();                                                      \
        if (_e != cudaSuccess) {                                                                  \
            (ctx)->set_error(std::string("kernel launch: ") + cudaGetErrorString(_e));            \
            return RETTO_B200_ERR_CUDA;                                                           \
        }                                                                                         \
    } while (0)

// copy a host descriptor array to the device, asynchronously on the stream: the bytes are snapshotted into a pinned
// staging slot (recycled once the event recorded behind its copy has completed), so callers may reuse `src` at once
retto_b200_status rt_upload(retto_b200_ctx* ctx, DevBuf& dst, const void* src, size_t bytes);
retto_b200_status rt_upload_to(retto_b200_ctx* ctx, void* d_dst, const void* src, size_t bytes);   // into an existing device range
// the same in two steps, for tables that are built in place in the pinned slot
retto_b200_status rt_stage_begin(retto_b200_ctx* ctx, size_t bytes, int* slot, void** p);
retto_b200_status rt_stage_commit(retto_b200_ctx* ctx, DevBuf& dst, int slot, size_t bytes);

// binary search: largest p with prefix[p] <= v   (prefix has n+1 entries, prefix[0] = 0)
__device__ __forceinline__ int rt_find_segment(const int* __restrict__ prefix, int n, int v) {
    int lo = 0, hi = n;
    while (hi - lo > 1) {
        int mid = (lo + hi) >> 1;
        if (prefix[mid] <= v) lo = mid; else hi = mid;
    }
    return lo;
}

