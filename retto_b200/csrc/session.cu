// session.cu — RettoSession::process_pipeline (session.rs:75-106) for a batch of pages, host side in C++
// above the stage entry points.  Order of operations per page is the reference's:
//   resize_both -> det.process (preprocess, worker.det, postprocess) -> get_crop_img per box ->
//   boxes.scale_and_clip to the original image -> cls.process (may flip crops) -> rec.process
// with every stage executed once for the whole batch (one launch per kernel, not per page).  Pages never
// touch host memory between stages; the forward passes are the caller's (worker.rs:69-73 seam).
//
// A batch is cut into units of pages; a unit is a resumable PageRun with three host steps (begin / mid / finish)
// separated by the only points where the host needs numbers from the device (component counts, boxes, strings).
// With retto_b200_set_pipeline(ctx, 2, unit) two units are kept in flight on two lanes (contexts with their own stream
// and buffers), software-pipelined from ONE host thread: while the host waits for / post-processes one unit, the
// kernels of the other are queued.  Measured on B200 this only pays for a continuous flow of units: a
// 256-page batch that must drain at the end of the call is faster as one unit on one stream (6.4 ms vs 7.1 ms for
// two units of 128), because the streams share the SMs evenly instead of letting the older unit run ahead — so the
// default is one lane.  Results are concatenated in page order; pages are independent, so the results do not depend
// on the cut.
#include "common.cuh"

retto_b200_status rt_scale_and_clip_multi(retto_b200_ctx* ctx, retto_b200_box* h_boxes, const double* h_params4, int n, bool defer);
retto_b200_status rt_cls_postprocess_ptrs(retto_b200_ctx* ctx, const std::vector<const float*>& logits, const int32_t* crop_index, int n,
                                          retto_b200_cls_result* h_results, bool defer);
retto_b200_status rt_cls_collect(retto_b200_ctx* ctx, int n, retto_b200_cls_result* h_results);
retto_b200_status rt_build_batches_prepare(retto_b200_ctx* ctx, int32_t kind, const retto_b200_line_job* h_lines, int32_t n_lines,
                                           uint64_t total_floats, float** d_base);
retto_b200_status rt_build_batches_launch(retto_b200_ctx* ctx, int32_t kind);
retto_b200_status rt_crop_launch_pages(retto_b200_ctx* ctx, const retto_b200_box* h_boxes, const int32_t* box_off, int n_pages,
                                       const uint8_t* const* page_ptr, const int* page_h, const int* page_w, retto_b200_crop_info* h_infos);
retto_b200_status rt_crop_finish(retto_b200_ctx* ctx, retto_b200_crop_info* h_infos, bool do_sync);
retto_b200_status rt_crop_enqueue_device(retto_b200_ctx* ctx, const retto_b200_box* d_boxes, const int* d_box_off, int n_pages,
                                         const uint8_t* const* page_ptr, const int* page_h, const int* page_w, int crops_hint, int max_boxes);
retto_b200_status rt_crop_adopt_device(retto_b200_ctx* ctx, const retto_b200_box* h_boxes, int n, retto_b200_crop_info* h_infos, bool* fits);
retto_b200_status rt_det_post_begin(retto_b200_ctx* ctx, const retto_b200_det_post_desc* h_descs, int32_t n, int32_t max_boxes_total);
retto_b200_status rt_det_post_mid(retto_b200_ctx* ctx);
retto_b200_status rt_det_post_end(retto_b200_ctx* ctx, int32_t* h_page_status, int32_t* h_box_offsets, retto_b200_box* h_boxes);
retto_b200_status rt_ctc_begin(retto_b200_ctx* ctx, const retto_b200_logits_desc* h_descs, int32_t n_descs, int32_t num_classes,
                               bool want_tokens, int32_t max_t_out);
retto_b200_status rt_ctc_end(retto_b200_ctx* ctx, uint32_t* h_text_offsets, char* h_text, size_t text_capacity, float* h_scores,
                             int32_t* h_tokens, int32_t* h_token_counts, int32_t max_t_out);

#include <atomic>
#include <chrono>
#include <memory>
#include <thread>
#define RUN_CHUNK_PAGES_MAX 64
#define PAGE_DEVICE_ENCODED 3   // internal: rgb = device copy of the file bytes, (h, w) from the parsed header
static int env_int(const char* name, int dflt, int lo, int hi) {
    const char* v = getenv(name);
    if (!v || !*v) return dflt;
    const int x = atoi(v);
    return x < lo ? lo : (x > hi ? hi : x);
}
namespace {
struct HostTrace {   // RETTO_B200_HOST_TRACE=1: wall-clock per stage of run_pages on stderr (host + waits)
    bool on;
    std::chrono::steady_clock::time_point t0, last;
    std::string line;
    HostTrace() : on(getenv("RETTO_B200_HOST_TRACE") != nullptr) { t0 = last = std::chrono::steady_clock::now(); }
    void mark(const char* name) {
        if (!on) return;
        auto now = std::chrono::steady_clock::now();
        char buf[96];
        snprintf(buf, sizeof(buf), " %s=%.3f", name, std::chrono::duration<double, std::milli>(now - last).count());
        line += buf;
        last = now;
    }
    ~HostTrace() {
        if (on) fprintf(stderr, "[run_pages ms]%s total=%.3f\n", line.c_str(), std::chrono::duration<double, std::milli>(last - t0).count());
    }
};
struct PageState {
    int ori_h, ori_w;     // decoded image
    int h, w;             // after resize_both
    const uint8_t* d_img; // after resize_both
    int det_h, det_w;
    float* d_det;
};
inline size_t align256(size_t v) { return (v + 255) & ~size_t(255); }
}  // namespace

namespace {
struct PageRun {
    retto_b200_ctx* ctx = nullptr;          // the lane this unit runs on
    const retto_b200_page* h_pages = nullptr;
    int n_pages = 0;
    retto_b200_forward_fn forward = nullptr;
    void* user = nullptr;
    retto_b200_stage_fn stage_cb = nullptr; // RettoWorkerStageResult delivery (session.rs:98,101,104), optional
    void* stage_user = nullptr;
    int first_page = 0;                     // index of this unit's first page in the caller's page array
    const JpegInfo* jinfo = nullptr;        // encoded pages: parsed headers, aligned with h_pages (the file bytes are already on the device)
    retto_b200_ctx* jowner = nullptr;       // the context whose entropy phase decoded the files of this call (the parent of a lane)
    int n_encoded = 0;
    std::vector<int> enc_page;              // page index of every decoded file of this unit
    // restart markers that do not match the DRI header are only seen by the device: per-file status, read after a stream sync
    void check_decode_status() {
        const int* hs = ctx->h_jpeg_status.as<int>();
        for (int k = 0; k < n_encoded; ++k)
            if (hs[k] != 0) {
                ctx->r_pages[enc_page[k]].status = hs[k];
                ret = (retto_b200_status)hs[k];
                ctx->set_error("run_pages: page " + std::to_string(first_page + enc_page[k]) + ": restart markers do not match the DRI header (damaged JPEG)");
            }
    }
    cudaEvent_t wait_for = nullptr;         // pages uploaded by the copy stream (chunked host batches)
    bool done = false;                      // finished early (no pages / no lines / error)
    retto_b200_status ret = RETTO_B200_OK;
    HostTrace tr;
    std::vector<PageState> ps;
    std::vector<retto_b200_tensor> det_in, det_out, tin, tout;
    std::vector<retto_b200_det_post_desc> dp_descs;
    std::vector<int32_t> page_status, box_off, cls_crop_idx;
    std::vector<retto_b200_crop_info> infos;
    std::vector<retto_b200_line_job> lines, rec_lines;
    std::vector<retto_b200_batch> batches, rec_batches;
    std::vector<int> batch_page, rec_batch_page;
    std::vector<uint32_t> toff;
    std::vector<float> sc;
    std::vector<char> text;
    int n_lines = 0, det_cap = 0;
    bool scaled = false;

    retto_b200_status fail(retto_b200_status s) { done = true; ret = s; return s; }
    // stage 0 = Det (boxes), 1 = Cls, 2 = Rec: the unit's results so far, in the reference's order and BEFORE the next stage's forward
    void emit_stage(int stage) {
        if (!stage_cb) return;
        retto_b200_stage_result r;
        memset(&r, 0, sizeof(r));
        r.stage = stage; r.first_page = first_page; r.n_pages = n_pages; r.pages = ctx->r_pages.data(); r.n_lines = n_lines;
        if (stage == 0) r.boxes = ctx->r_boxes.data();
        else if (stage == 1) r.cls = ctx->r_cls.data();
        else { r.text_offsets = ctx->r_text_offs.data(); r.text = ctx->r_text.data(); r.rec_scores = ctx->r_scores.data(); }
        stage_cb(stage_user, &r);
    }
    retto_b200_status begin();
    retto_b200_status mid();
    retto_b200_status finish();
    retto_b200_status plan_all(int kind, std::vector<retto_b200_line_job>& ln, std::vector<retto_b200_batch>& bt, std::vector<int>& bpage, uint64_t* total);
};

// ---- steps 1-4a: pages to the device, resize_both, det preprocess, worker.det, det postprocess up to the component table
retto_b200_status PageRun::begin() {
    cudaStream_t st = ctx->stream;
    const retto_b200_config& cfg = ctx->cfg;
    ctx->r_pages.assign(n_pages, retto_b200_page_result{0, 0, 0});
    ctx->r_boxes.clear(); ctx->r_cls.clear(); ctx->r_text_offs.assign(1, 0); ctx->r_text.assign(1, 0); ctx->r_scores.clear();
    memset(ctx->run_stats, 0, sizeof(ctx->run_stats));
    if (n_pages == 0) { done = true; return RETTO_B200_OK; }
    if (wait_for) RT_CUDA_OK(ctx, cudaStreamWaitEvent(st, wait_for, 0));

    // ---- 1. pages to the device, resize_both (image_helper.rs:106-148) ------------------------------------
    ps.resize(n_pages);
    size_t raw_bytes = 0, rs_bytes = 0, det_floats = 0;
    std::vector<std::vector<std::pair<int, int>>> rs_steps(n_pages);
    for (int i = 0; i < n_pages; ++i) {
        const retto_b200_page& p = h_pages[i];
        if (!p.rgb || p.h <= 0 || p.w <= 0) { ctx->set_error("run_pages: bad page " + std::to_string(i)); return fail(RETTO_B200_ERR_INVALID_ARG); }
        ps[i].ori_h = p.h; ps[i].ori_w = p.w;
        if (p.on_device == RETTO_B200_PAGE_HOST_RGB || p.on_device == PAGE_DEVICE_ENCODED) raw_bytes += align256((size_t)p.h * p.w * 3);
        int dims[4], ns = 0;
        RT_TRY(retto_b200_resize_both_plan(p.h, p.w, cfg.max_side_len, cfg.min_side_len, dims, &ns));
        int h = p.h, w = p.w;
        for (int s = 0; s < ns; ++s) {
            h = dims[2 * s]; w = dims[2 * s + 1];
            if (h <= 0 || w <= 0) { ctx->set_error("run_pages: page " + std::to_string(i) + " resizes to nothing"); return fail(RETTO_B200_ERR_INVALID_ARG); }
            rs_steps[i].push_back({h, w});
            rs_bytes += align256((size_t)h * w * 3);
        }
        ps[i].h = h; ps[i].w = w;
        RT_TRY(retto_b200_resize_either_plan(h, w, cfg.det_limit_type, cfg.det_limit_side_len, &ps[i].det_h, &ps[i].det_w));
        if (ps[i].det_h <= 0 || ps[i].det_w <= 0 || ps[i].det_h > cfg.max_det_side || ps[i].det_w > cfg.max_det_side) {
            ctx->set_error("run_pages: det tensor of page " + std::to_string(i) + " is " + std::to_string(ps[i].det_h) + "x" + std::to_string(ps[i].det_w) +
                           " (cap max_det_side=" + std::to_string(cfg.max_det_side) + ")");
            return fail(RETTO_B200_ERR_CAPACITY);
        }
        det_floats += (align256((size_t)3 * ps[i].det_h * ps[i].det_w * 4)) / 4;
    }
    RT_CUDA_OK(ctx, ctx->d_pages_raw.ensure(std::max<size_t>(raw_bytes, 256), st));
    RT_CUDA_OK(ctx, ctx->d_pages_rs.ensure(std::max<size_t>(rs_bytes, 256), st));
    RT_CUDA_OK(ctx, ctx->d_det_in.ensure(std::max<size_t>(det_floats * 4, 256), st));
    {
        size_t off = 0, roff = 0;
        std::vector<retto_b200_resize_desc> step1, step2;
        std::vector<uint8_t*> enc_dst;
        for (int i = 0; i < n_pages; ++i) {
            const retto_b200_page& p = h_pages[i];
            const uint8_t* cur = p.rgb;
            if (p.on_device == RETTO_B200_PAGE_HOST_RGB) {
                uint8_t* d = ctx->d_pages_raw.as<uint8_t>() + off;
                RT_CUDA_OK(ctx, cudaMemcpyAsync(d, p.rgb, (size_t)p.h * p.w * 3, cudaMemcpyHostToDevice, st));
                off += align256((size_t)p.h * p.w * 3);
                cur = d;
            } else if (p.on_device == PAGE_DEVICE_ENCODED) {   // image_helper.rs:34-44 on the device: the decoded page lands in the raw-page arena
                uint8_t* d = ctx->d_pages_raw.as<uint8_t>() + off;
                off += align256((size_t)p.h * p.w * 3);
                enc_dst.push_back(d); enc_page.push_back(i);
                cur = d;
            }
            int ch = p.h, cw = p.w;
            for (size_t s = 0; s < rs_steps[i].size(); ++s) {
                uint8_t* d = ctx->d_pages_rs.as<uint8_t>() + roff;
                roff += align256((size_t)rs_steps[i][s].first * rs_steps[i][s].second * 3);
                (s == 0 ? step1 : step2).push_back(retto_b200_resize_desc{cur, ch, cw, d, rs_steps[i][s].first, rs_steps[i][s].second});
                cur = d; ch = rs_steps[i][s].first; cw = rs_steps[i][s].second;
            }
            ps[i].d_img = cur;
        }
        n_encoded = (int)enc_dst.size();
        if (n_encoded) {   // pixel phase of this unit's files (the entropy phase of the whole call ran on the copy stream)
            if (n_encoded != n_pages || !jowner) { ctx->set_error("run_pages: internal: a unit mixes encoded and decoded pages"); return fail(RETTO_B200_ERR_INVALID_ARG); }
            retto_b200_status js = rt_jpeg_pixels_enqueue(jowner, ctx, first_page, n_encoded, enc_dst.data());
            if (js != RETTO_B200_OK) return fail(js);
            RT_CUDA_OK(ctx, ctx->h_jpeg_status.ensure(sizeof(int) * (size_t)n_encoded));
            RT_CUDA_OK(ctx, cudaMemcpyAsync(ctx->h_jpeg_status.p, jowner->jpeg_status_dev + first_page, sizeof(int) * (size_t)n_encoded, cudaMemcpyDeviceToHost, st));
        }
        if (!step1.empty()) RT_TRY(retto_b200_thumbnail(ctx, step1.data(), (int)step1.size()));
        if (!step2.empty()) RT_TRY(retto_b200_thumbnail(ctx, step2.data(), (int)step2.size()));
    }
    tr.mark("upload+resize");
    // ---- 2. det preprocess (det_processor.rs:256-274) --------------------------------------------------------
    det_in.resize(n_pages); det_out.resize(n_pages);
    {
        std::vector<retto_b200_det_pre_desc> descs(n_pages);
        size_t off = 0;
        for (int i = 0; i < n_pages; ++i) {
            ps[i].d_det = ctx->d_det_in.as<float>() + off;
            off += align256((size_t)3 * ps[i].det_h * ps[i].det_w * 4) / 4;
            descs[i] = retto_b200_det_pre_desc{ps[i].d_img, ps[i].h, ps[i].w, ps[i].d_det, ps[i].det_h, ps[i].det_w};
            det_in[i].d_data = ps[i].d_det;
            det_in[i].shape[0] = 1; det_in[i].shape[1] = 3; det_in[i].shape[2] = ps[i].det_h; det_in[i].shape[3] = ps[i].det_w;
            det_in[i].ndim = 4;
            memset(&det_out[i], 0, sizeof(retto_b200_tensor));
        }
        RT_TRY(retto_b200_det_preprocess(ctx, descs.data(), n_pages));
    }
    tr.mark("det_pre");
    // ---- 3. worker.det (session.rs:86) ------------------------------------------------------------------------
    if (forward(user, 0, n_pages, det_in.data(), det_out.data(), (void*)st) != 0) { ctx->set_error("run_pages: det forward failed"); return fail(RETTO_B200_ERR_WORKER); }
    tr.mark("det_fwd");
    // ---- 4. det postprocess (det_processor.rs:279-335), first part ------------------------------------------------
    page_status.assign(n_pages, 0); box_off.assign(n_pages + 1, 0);
    dp_descs.resize(n_pages);
    for (int i = 0; i < n_pages; ++i) {
        const retto_b200_tensor& t = det_out[i];
        if (!t.d_data || t.ndim != 4 || t.shape[0] != 1 || t.shape[1] != 1 || t.shape[2] <= 0 || t.shape[3] <= 0) {
            ctx->set_error("run_pages: det forward returned a bad tensor for page " + std::to_string(i));
            return fail(RETTO_B200_ERR_WORKER);
        }
        // DetProcessor::new(cfg, after_h, after_w): boxes are scaled to the resize_both-ed page (session.rs:85)
        dp_descs[i] = retto_b200_det_post_desc{t.d_data, (int32_t)t.shape[2], (int32_t)t.shape[3], ps[i].h, ps[i].w};
    }
    det_cap = std::max(4096, n_pages * 256);
    retto_b200_status s = rt_det_post_begin(ctx, dp_descs.data(), n_pages, det_cap);
    if (s != RETTO_B200_OK) return fail(s);
    tr.mark("det_post_begin");
    return RETTO_B200_OK;
}

retto_b200_status PageRun::plan_all(int kind, std::vector<retto_b200_line_job>& ln, std::vector<retto_b200_batch>& bt, std::vector<int>& bpage, uint64_t* total) {
    ln.assign(n_lines, retto_b200_line_job{});
    bt.clear(); bpage.clear();
    uint64_t off = 0;
    std::vector<retto_b200_batch> pb;
    for (int i = 0; i < n_pages; ++i) {
        const int nb = box_off[i + 1] - box_off[i];
        if (nb == 0) continue;
        pb.assign((size_t)nb, retto_b200_batch{});
        int32_t nbat = 0; uint64_t tot = 0;
        RT_TRY(retto_b200_plan_batches(&ctx->cfg, kind, infos.data() + box_off[i], nb, ln.data() + box_off[i], pb.data(), &nbat, &tot));
        for (int k = box_off[i]; k < box_off[i + 1]; ++k) { ln[k].crop += box_off[i]; ln[k].dst_offset += off; }
        for (int b = 0; b < nbat; ++b) { pb[b].first_line += box_off[i]; pb[b].offset += off; bt.push_back(pb[b]); bpage.push_back(i); }
        off += tot;
    }
    *total = off;
    return RETTO_B200_OK;
}

// ---- steps 4b-8: geometry, boxes to the host, crops, cls, rec enqueued up to the CTC read-back
retto_b200_status PageRun::mid() {
    if (done) return ret;
    cudaStream_t st = ctx->stream;
    const retto_b200_config& cfg = ctx->cfg;
    const bool dev_crops = getenv("RETTO_B200_HOST_CROP_TABLE") == nullptr;   // A/B + tests: the host-built descriptor table
    ctx->crops_lazy = getenv("RETTO_B200_CROP_EAGER") == nullptr;             // direct crops are read from the page by the batch build (A/B: materialise all)
    std::vector<const uint8_t*> pp(n_pages);
    std::vector<int> ph(n_pages), pw(n_pages);
    for (int i = 0; i < n_pages; ++i) { pp[i] = ps[i].d_img; ph[i] = ps[i].h; pw[i] = ps[i].w; }
    for (int attempt = 0; attempt < 2; ++attempt) {
        ctx->r_boxes.resize(det_cap);
        retto_b200_status s = rt_det_post_mid(ctx);
        if (s == RETTO_B200_OK && dev_crops) {
            // 5a. crops straight from the packed device boxes (session.rs:88-92), enqueued before the host waits for them
            const int* d_off = reinterpret_cast<const int*>(ctx->d_dp_counters.as<char>() + sizeof(PageCounters) * (size_t)n_pages);
            const int hint = ctx->crops_seen_max + ctx->crops_seen_max / 4 + 64;
            s = rt_crop_enqueue_device(ctx, ctx->d_boxes_out.as<retto_b200_box>(), d_off, n_pages, pp.data(), ph.data(), pw.data(), hint, det_cap);
        }
        if (s == RETTO_B200_OK) s = rt_det_post_end(ctx, page_status.data(), box_off.data(), ctx->r_boxes.data());
        if (s == RETTO_B200_ERR_CAPACITY && box_off[n_pages] > det_cap && attempt == 0) {   // more boxes than planned for: redo with room
            det_cap = box_off[n_pages];
            s = rt_det_post_begin(ctx, dp_descs.data(), n_pages, det_cap);
            if (s != RETTO_B200_OK) return fail(s);
            continue;
        }
        if (s != RETTO_B200_OK) return fail(s);
        break;
    }
    tr.mark("det_post");
    n_lines = box_off[n_pages];
    ctx->r_boxes.resize(n_lines);
    for (int i = 0; i < n_pages; ++i) {
        ctx->r_pages[i].status = page_status[i];
        ctx->r_pages[i].first_line = box_off[i];
        ctx->r_pages[i].n_lines = box_off[i + 1] - box_off[i];
        if (page_status[i] != RETTO_B200_OK) { ret = (retto_b200_status)page_status[i]; ctx->set_error("run_pages: det postprocess status on page " + std::to_string(i)); }
    }
    ctx->r_cls.assign(n_lines, retto_b200_cls_result{0, 0.0f});
    ctx->r_scores.assign(n_lines, 0.0f);
    ctx->r_text_offs.assign(n_lines + 1, 0);
    if (n_lines == 0) {   // no detections: cls / rec run zero batches and return empty results (App. A #20); the three stages still report
        done = true;
        check_decode_status();
        emit_stage(0); emit_stage(1); emit_stage(2);
        return ret;
    }

    // ---- 5. crops from the resize_both-ed page (session.rs:88-92) --------------------------------------------------
    // From here to the CTC read-back nothing waits for the GPU: the plans depend only on the crop dims (known on the
    // host), descriptor tables are built in pinned staging slots, and the crop statuses / cls results / rescaled boxes
    // are collected after the one final sync — the host runs ahead and the kernels queue back to back.
    infos.resize(n_lines);
    {
        bool fits = false;
        retto_b200_status s = RETTO_B200_OK;
        if (dev_crops) s = rt_crop_adopt_device(ctx, ctx->r_boxes.data(), n_lines, infos.data(), &fits);
        ctx->crops_seen_max = std::max(ctx->crops_seen_max, n_lines);
        if (s == RETTO_B200_OK && !fits)   // first batches of a context (tables still growing) or RETTO_B200_HOST_CROP_TABLE: host-built table
            s = rt_crop_launch_pages(ctx, ctx->r_boxes.data(), box_off.data(), n_pages, pp.data(), ph.data(), pw.data(), infos.data());
        if (s != RETTO_B200_OK) return fail(s);
    }
    tr.mark("crop_launch");
    // ---- 7. cls (cls_processor.rs:127-172) ------------------------------------------------------------------------------
    uint64_t total = 0, rec_total = 0;
    retto_b200_status s = plan_all(0, lines, batches, batch_page, &total);
    if (s != RETTO_B200_OK) return fail(s);
    float *d_base = nullptr, *d_base_rec = nullptr;
    if ((s = rt_build_batches_prepare(ctx, 0, lines.data(), n_lines, total, &d_base)) != RETTO_B200_OK) return fail(s);
    if ((s = rt_build_batches_launch(ctx, 0)) != RETTO_B200_OK) return fail(s);
    tr.mark("cls_build");
    // ---- 6. boxes back to original-image coordinates (session.rs:94-97) ----------------------------------------------
    {
        bool any = false;
        for (int i = 0; i < n_pages; ++i) if (ps[i].h != ps[i].ori_h || ps[i].w != ps[i].ori_w) any = true;  // identity otherwise: round(x * 1) clamped == x
        if (any) {
            std::vector<double> prm((size_t)n_lines * 4);
            for (int i = 0; i < n_pages; ++i)
                for (int k = box_off[i]; k < box_off[i + 1]; ++k) {
                    prm[4 * (size_t)k] = (double)ps[i].ori_w / (double)ps[i].w;
                    prm[4 * (size_t)k + 1] = (double)ps[i].ori_h / (double)ps[i].h;
                    prm[4 * (size_t)k + 2] = (double)ps[i].ori_w;
                    prm[4 * (size_t)k + 3] = (double)ps[i].ori_h;
                }
            if ((s = rt_scale_and_clip_multi(ctx, ctx->r_boxes.data(), prm.data(), n_lines, true)) != RETTO_B200_OK) return fail(s);
            scaled = true;
        }
    }
    // host work while the GPU crops and builds the cls batches: the rec plan and the rec descriptor tables
    if ((s = plan_all(1, rec_lines, rec_batches, rec_batch_page, &rec_total)) != RETTO_B200_OK) return fail(s);
    if ((s = rt_build_batches_prepare(ctx, 1, rec_lines.data(), n_lines, rec_total, &d_base_rec)) != RETTO_B200_OK) return fail(s);
    {
        uint64_t det_px = 0, crop_px = 0, rec_rows = 0;
        for (int i = 0; i < n_pages; ++i) det_px += (uint64_t)ps[i].det_h * ps[i].det_w;
        for (int k = 0; k < n_lines; ++k) crop_px += (uint64_t)infos[k].w * infos[k].h;
        for (const auto& b : rec_batches) rec_rows += (uint64_t)b.n * (b.img_w / 8);
        const uint64_t st8[8] = {(uint64_t)n_pages, (uint64_t)n_lines, det_px, crop_px, total, rec_total, rec_rows, 0};
        memcpy(ctx->run_stats, st8, sizeof(st8));
    }
    tr.mark("rec_prepare");
    auto fill_inputs = [&](const std::vector<retto_b200_batch>& bt, float* base, int img_h) {
        tin.assign(bt.size(), retto_b200_tensor{});
        tout.assign(bt.size(), retto_b200_tensor{});
        for (size_t b = 0; b < bt.size(); ++b) {
            tin[b].d_data = base + bt[b].offset;
            tin[b].shape[0] = bt[b].n; tin[b].shape[1] = 3; tin[b].shape[2] = img_h; tin[b].shape[3] = bt[b].img_w;
            tin[b].ndim = 4;
        }
    };
    if (stage_cb) {
        // streaming callers get the Det result before worker.cls runs (session.rs:98): that needs the rescaled boxes on the host
        // now instead of at the end of the unit — one extra stream sync per unit, paid only when a stage callback is set
        RT_CUDA_OK(ctx, cudaStreamSynchronize(st));
        if (scaled) { memcpy(ctx->r_boxes.data(), ctx->h_scale.p, sizeof(retto_b200_box) * (size_t)n_lines); scaled = false; }
        emit_stage(0);
    }
    fill_inputs(batches, d_base, cfg.cls_image_shape[1]);
    if (forward(user, 1, (int)batches.size(), tin.data(), tout.data(), (void*)st) != 0) { ctx->set_error("run_pages: cls forward failed"); return fail(RETTO_B200_ERR_WORKER); }
    {
        std::vector<const float*> ptrs(n_lines);
        cls_crop_idx.resize(n_lines);
        for (size_t b = 0; b < batches.size(); ++b) {
            const retto_b200_tensor& t = tout[b];
            if (!t.d_data || t.ndim != 2 || t.shape[0] != batches[b].n || t.shape[1] != 2) {
                ctx->set_error("run_pages: cls forward returned a bad tensor");
                return fail(RETTO_B200_ERR_WORKER);
            }
            for (int k = 0; k < batches[b].n; ++k) {
                ptrs[batches[b].first_line + k] = t.d_data + 2 * (size_t)k;
                cls_crop_idx[batches[b].first_line + k] = lines[batches[b].first_line + k].crop;
            }
        }
        if ((s = rt_cls_postprocess_ptrs(ctx, ptrs, cls_crop_idx.data(), n_lines, nullptr, true)) != RETTO_B200_OK) return fail(s);   // collected after the final sync
        if (stage_cb) {   // the Cls result before worker.rec runs (session.rs:101)
            RT_CUDA_OK(ctx, cudaStreamSynchronize(st));
            std::vector<retto_b200_cls_result> res(n_lines);
            if ((s = rt_cls_collect(ctx, n_lines, res.data())) != RETTO_B200_OK) return fail(s);
            for (int k = 0; k < n_lines; ++k) ctx->r_cls[cls_crop_idx[k]] = res[k];
            emit_stage(1);
        }
    }
    tr.mark("cls_fwd+post");
    // ---- 8. rec (rec_processor.rs:214-270) ---------------------------------------------------------------------------------
    if ((s = rt_build_batches_launch(ctx, 1)) != RETTO_B200_OK) return fail(s);
    fill_inputs(rec_batches, d_base_rec, cfg.rec_image_shape[1]);
    if (forward(user, 2, (int)rec_batches.size(), tin.data(), tout.data(), (void*)st) != 0) { ctx->set_error("run_pages: rec forward failed"); return fail(RETTO_B200_ERR_WORKER); }
    {
        std::vector<retto_b200_logits_desc> descs(rec_batches.size());
        int max_t = 1;
        for (size_t b = 0; b < rec_batches.size(); ++b) {
            const retto_b200_tensor& t = tout[b];
            if (!t.d_data || t.ndim != 3 || t.shape[0] != rec_batches[b].n || t.shape[1] <= 0 || t.shape[2] != (int64_t)ctx->dict.size()) {
                ctx->set_error("run_pages: rec forward returned a bad tensor (classes must equal the dictionary size " + std::to_string(ctx->dict.size()) + ")");
                return fail(ctx->dict.empty() ? RETTO_B200_ERR_NO_DICT : RETTO_B200_ERR_WORKER);
            }
            descs[b] = retto_b200_logits_desc{t.d_data, (int32_t)t.shape[0], (int32_t)t.shape[1]};
            max_t = std::max(max_t, (int)t.shape[1]);
        }
        if (ctx->dict.empty()) { ctx->set_error("ctc_decode: no dictionary loaded"); return fail(RETTO_B200_ERR_NO_DICT); }
        toff.assign(n_lines + 1, 0);
        sc.assign(n_lines, 0.0f);
        text.resize((size_t)n_lines * max_t * std::max(ctx->dict_max_len, 1) + 16);
        if ((s = rt_ctc_begin(ctx, descs.data(), (int)descs.size(), (int)ctx->dict.size(), false, 0)) != RETTO_B200_OK) return fail(s);
    }
    tr.mark("rec_fwd+ctc_begin");
    return RETTO_B200_OK;
}

// ---- the one final sync of the unit + results in detection order
retto_b200_status PageRun::finish() {
    if (done) return ret;
    done = true;
    retto_b200_status s = rt_ctc_end(ctx, toff.data(), text.data(), text.size(), sc.data(), nullptr, nullptr, 0);
    if (s != RETTO_B200_OK) return (ret = s);
    tr.mark("ctc_wait");
    // the CTC call synchronised the stream: the deferred cls results, crop statuses and rescaled boxes are on the host now
    std::vector<retto_b200_cls_result> res(n_lines);
    if ((s = rt_cls_collect(ctx, n_lines, res.data())) != RETTO_B200_OK) return (ret = s);
    for (int k = 0; k < n_lines; ++k) ctx->r_cls[cls_crop_idx[k]] = res[k];  // final_res[idx].label = label (cls_processor.rs:167)
    if (scaled) memcpy(ctx->r_boxes.data(), ctx->h_scale.p, sizeof(retto_b200_box) * (size_t)n_lines);
    if ((s = rt_crop_finish(ctx, infos.data(), false)) != RETTO_B200_OK) return (ret = s);
    // scatter from plan order back to detection order (rec_processor.rs:259-264)
    std::vector<uint32_t> len(n_lines, 0);
    for (int k = 0; k < n_lines; ++k) len[rec_lines[k].crop] = toff[k + 1] - toff[k];
    for (int k = 0; k < n_lines; ++k) ctx->r_text_offs[k + 1] = ctx->r_text_offs[k] + len[k];
    ctx->r_text.assign(ctx->r_text_offs[n_lines] + 1, 0);
    for (int k = 0; k < n_lines; ++k) {
        const int dst = rec_lines[k].crop;
        memcpy(ctx->r_text.data() + ctx->r_text_offs[dst], text.data() + toff[k], toff[k + 1] - toff[k]);
        ctx->r_scores[dst] = sc[k];
    }
    tr.mark("collect");
    check_decode_status();
    emit_stage(2);   // session.rs:104
    return ret;
}

void fill_results(retto_b200_ctx* ctx, retto_b200_results* out) {
    out->n_pages = (int)ctx->r_pages.size();
    out->pages = ctx->r_pages.data();
    out->n_lines = (int)ctx->r_boxes.size();
    out->boxes = ctx->r_boxes.data();
    out->cls = ctx->r_cls.data();
    out->text_offsets = ctx->r_text_offs.data();
    out->text = ctx->r_text.data();
    out->rec_scores = ctx->r_scores.data();
}
}  // namespace

// Page upload by the SMs: pinned host memory is device-accessible under UVA, so a copy kernel on the copy stream can
// pull the pages over PCIe without occupying the DMA engine — the many small descriptor uploads of the compute stream
// would otherwise queue behind megabytes of page copies in the same H2D engine FIFO and serialise the pipeline.
struct PullArgs { const uint4* src[RUN_CHUNK_PAGES_MAX]; uint4* dst[RUN_CHUNK_PAGES_MAX]; unsigned n16[RUN_CHUNK_PAGES_MAX]; int n; };
// A SMALL grid on purpose (PULL_BLOCKS blocks of 512 threads, 8 x 16 B in flight per thread = 64 KB per block): PCIe needs
// only ~0.2 MB in flight to saturate, and a grid that filled the SMs with threads waiting on PCIe would starve the
// compute stream's kernels of block slots — the overlap this pipeline exists for.
__global__ void __launch_bounds__(512) pull_pages_kernel(PullArgs a) {
    const unsigned stride = gridDim.x * blockDim.x;
    for (int p = 0; p < a.n; ++p) {
        const uint4* __restrict__ s = a.src[p];
        uint4* __restrict__ d = a.dst[p];
        const unsigned n = a.n16[p];
        unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
        for (; i + 7 * stride < n; i += 8 * stride) {
            uint4 v[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) v[k] = __ldcs(s + i + k * stride);
#pragma unroll
            for (int k = 0; k < 8; ++k) d[i + k * stride] = v[k];
        }
        for (; i < n; i += stride) d[i] = __ldcs(s + i);
    }
}

// lane 1: a child context on the same device (own stream, buffers, dictionary copy), created on first use
static retto_b200_status lane_ctx(retto_b200_ctx* ctx, int lane, retto_b200_ctx** out) {
    if (lane == 0) { *out = ctx; return RETTO_B200_OK; }
    while ((int)ctx->lanes.size() < lane) {
        retto_b200_ctx* c = nullptr;
        retto_b200_status s = retto_b200_create(ctx->device, &ctx->cfg, &c);
        if (s != RETTO_B200_OK) { ctx->set_error("run_pages: cannot create a pipeline lane"); return s; }
        ctx->lanes.push_back(c);
    }
    retto_b200_ctx* c = ctx->lanes[lane - 1];
    if (c->dict_version != ctx->dict_version) {
        retto_b200_status s = retto_b200_dict_load(c, ctx->dict_source.data(), ctx->dict_source.size());
        if (s != RETTO_B200_OK) return s;
        c->dict_version = ctx->dict_version;
    }
    c->cfg = ctx->cfg;
    c->uploads_by_sm = ctx->uploads_by_sm;
    c->timing_enabled = false;
    *out = c;
    return RETTO_B200_OK;
}

struct Unit { const retto_b200_page* pages; int n; cudaEvent_t wait_for; int first_page; const JpegInfo* jinfo; };

// Software pipeline over the units, two in flight: finish(u[k-2]) -> begin(u[k]) -> mid(u[k-1]).  Unit k runs on lane
// k % 2, which unit k-2 has just left.  With one lane the units run back to back.
static retto_b200_status run_units(retto_b200_ctx* ctx, const std::vector<Unit>& units, int n_lanes, retto_b200_forward_fn forward, void* user,
                                   retto_b200_results* out) {
    const int nu = (int)units.size();
    n_lanes = std::max(1, std::min(n_lanes, std::min(nu, 2)));
    std::vector<retto_b200_ctx*> lane(n_lanes);
    for (int l = 0; l < n_lanes; ++l) RT_TRY(lane_ctx(ctx, l, &lane[l]));
    std::vector<retto_b200_page_result> a_pages;
    std::vector<retto_b200_box> a_boxes;
    std::vector<retto_b200_cls_result> a_cls;
    std::vector<uint32_t> a_toffs(1, 0);
    std::vector<char> a_text;
    std::vector<float> a_scores;
    uint64_t stats[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    retto_b200_status ret = RETTO_B200_OK;
    std::vector<std::unique_ptr<PageRun>> runs(nu);
    auto hard = [](retto_b200_status s) { return s != RETTO_B200_OK && s != RETTO_B200_ERR_DEGENERATE_QUAD && s != RETTO_B200_ERR_CAPACITY; };
    auto drain = [&](retto_b200_status s, retto_b200_ctx* from) {
        for (auto* l : lane) cudaStreamSynchronize(l->stream);
        if (from != ctx) ctx->set_error(from->err);
        return s;
    };
    auto collect = [&](PageRun& r) {   // append the unit's results (units complete in page order)
        retto_b200_results u;
        fill_results(r.ctx, &u);
        const int line0 = (int)a_boxes.size();
        for (int i = 0; i < u.n_pages; ++i) { retto_b200_page_result pr = u.pages[i]; pr.first_line += line0; a_pages.push_back(pr); }
        a_boxes.insert(a_boxes.end(), u.boxes, u.boxes + u.n_lines);
        a_cls.insert(a_cls.end(), u.cls, u.cls + u.n_lines);
        a_scores.insert(a_scores.end(), u.rec_scores, u.rec_scores + u.n_lines);
        const uint32_t t0 = a_toffs.back();
        for (int k = 0; k < u.n_lines; ++k) a_toffs.push_back(t0 + u.text_offsets[k + 1]);
        if (u.n_lines) a_text.insert(a_text.end(), u.text, u.text + u.text_offsets[u.n_lines]);
        for (int k = 0; k < 8; ++k) stats[k] += r.ctx->run_stats[k];
    };
    auto start = [&](int k) {
        runs[k].reset(new PageRun());
        PageRun& r = *runs[k];
        r.ctx = lane[k % n_lanes]; r.h_pages = units[k].pages; r.n_pages = units[k].n; r.forward = forward; r.user = user; r.wait_for = units[k].wait_for;
        r.stage_cb = ctx->stage_cb; r.stage_user = ctx->stage_user; r.first_page = units[k].first_page; r.jinfo = units[k].jinfo; r.jowner = ctx;
        return r.begin();
    };
    if (n_lanes == 1) {
        for (int k = 0; k < nu; ++k) {
            retto_b200_status s = start(k);
            if (!hard(s)) s = runs[k]->mid();
            if (!hard(s)) s = runs[k]->finish();
            if (hard(s)) return drain(s, lane[0]);
            if (s != RETTO_B200_OK) ret = s;
            if (nu == 1) { fill_results(ctx, out); return ret; }   // single unit: the context's own vectors are the result
            collect(*runs[k]);
            runs[k].reset();
        }
    } else {
        for (int k = 0; k < nu + 2; ++k) {
            if (k >= 2) {
                retto_b200_status s = runs[k - 2]->finish();
                if (hard(s)) return drain(s, runs[k - 2]->ctx);
                if (s != RETTO_B200_OK) { ret = s; if (runs[k - 2]->ctx != ctx) ctx->set_error(runs[k - 2]->ctx->err); }
                collect(*runs[k - 2]);
                runs[k - 2].reset();
            }
            if (k < nu) {
                retto_b200_status s = start(k);
                if (hard(s)) return drain(s, runs[k]->ctx);
            }
            if (k >= 1 && k <= nu) {
                retto_b200_status s = runs[k - 1]->mid();
                if (hard(s)) return drain(s, runs[k - 1]->ctx);
            }
        }
    }
    memcpy(ctx->run_stats, stats, sizeof(stats));
    a_text.push_back(0);
    ctx->r_pages.swap(a_pages); ctx->r_boxes.swap(a_boxes); ctx->r_cls.swap(a_cls);
    ctx->r_text_offs.swap(a_toffs); ctx->r_text.swap(a_text); ctx->r_scores.swap(a_scores);
    fill_results(ctx, out);
    return ret;
}

// Host-resident batches larger than one unit are uploaded by the pull kernel on a separate copy stream up front (one event
// per unit); a unit starts as soon as its pages have landed, so the PCIe transfer of the later units overlaps with the
// kernels of the earlier ones.  Device-resident batches are cut into units only to keep two lanes busy.
extern "C" retto_b200_status retto_b200_run_pages(retto_b200_ctx* ctx, const retto_b200_page* h_pages, int32_t n_pages,
                                                  retto_b200_forward_fn forward, void* user, retto_b200_results* out) {
    RtDeviceGuard _dg(ctx);
    if (!ctx || n_pages < 0 || (n_pages > 0 && !h_pages) || !forward || !out) return RETTO_B200_ERR_INVALID_ARG;
    memset(out, 0, sizeof(*out));
    // tunables (retto_b200_set_pipeline or the environment); defaults measured on B200
    const int n_lanes = ctx->pipe_lanes > 0 ? ctx->pipe_lanes : env_int("RETTO_B200_LANES", 1, 1, 2);
    const int unit_dev = ctx->pipe_unit_pages > 0 ? ctx->pipe_unit_pages : env_int("RETTO_B200_UNIT_PAGES", 64, 1, 1 << 20);
    const int RUN_CHUNK_PAGES = std::min(RUN_CHUNK_PAGES_MAX, ctx->pipe_unit_pages > 0 ? ctx->pipe_unit_pages : env_int("RETTO_B200_CHUNK_PAGES", 32, 1, RUN_CHUNK_PAGES_MAX));
    static const int PULL_BLOCKS = env_int("RETTO_B200_PULL_BLOCKS", 16, 1, 1024);
    // pages by the copy engine (55 GB/s measured) with the descriptor tables of the compute stream pulled by the SMs, or
    // pages pulled by pull_pages_kernel (48 GB/s) with ordinary descriptor copies
    static const int USE_DMA = env_int("RETTO_B200_PULL_DMA", 1, 0, 1);
    // ---- encoded pages (file bytes): parse the headers here, upload the files on the copy stream, decode per unit on the device
    {
        int n_enc = 0;
        for (int i = 0; i < n_pages; ++i) n_enc += h_pages[i].on_device == RETTO_B200_PAGE_HOST_ENCODED;
        if (n_enc > 0) {
            if (n_enc != n_pages) { ctx->set_error("run_pages: encoded and decoded pages cannot be mixed in one call"); return RETTO_B200_ERR_INVALID_ARG; }
            std::vector<JpegInfo>& infos = ctx->jpeg_infos;
            infos.resize(n_pages);
            size_t blob = 0;
            for (int i = 0; i < n_pages; ++i) blob += ((size_t)h_pages[i].n_bytes + 31) & ~size_t(15);
            const int unit = ctx->pipe_unit_pages > 0 ? ctx->pipe_unit_pages : env_int("RETTO_B200_ENC_UNIT_PAGES", 1 << 20, 1, 1 << 20);   // default: one unit
            if (!ctx->copy_stream) {
                int lo = 0, hi = 0;
                cudaDeviceGetStreamPriorityRange(&lo, &hi);
                RT_CUDA_OK(ctx, cudaStreamCreateWithPriority(&ctx->copy_stream, cudaStreamNonBlocking, hi));
            }
            const int n_units = (n_pages + unit - 1) / unit;
            if (ctx->copy_events.empty()) {
                cudaEvent_t e;
                RT_CUDA_OK(ctx, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
                ctx->copy_events.push_back(e);
            }
            RT_CUDA_OK(ctx, ctx->d_jpeg_blob.ensure(blob + 16, ctx->stream));
            std::vector<retto_b200_page> dev_pages(n_pages);
            std::vector<const uint8_t*> d_files(n_pages);
            size_t off = 0;
            uint64_t enc_bytes = 0;
            // header parse (host, ~1 us per file) and upload of file i in one loop: the copy engine works while the next headers are read
            for (int i = 0; i < n_pages; ++i) {
                const retto_b200_status s = h_pages[i].rgb ? rt_jpeg_parse(h_pages[i].rgb, (size_t)h_pages[i].n_bytes, &infos[i]) : RETTO_B200_ERR_INVALID_ARG;
                if (s != RETTO_B200_OK) {
                    cudaStreamSynchronize(ctx->copy_stream);
                    ctx->set_error("run_pages: page " + std::to_string(i) + (s == RETTO_B200_ERR_UNSUPPORTED ? ": not a baseline JPEG this decoder covers (decode it on the host and pass RGB)" : ": damaged or unknown image file"));
                    return s;
                }
                uint8_t* d = ctx->d_jpeg_blob.as<uint8_t>() + off;
                off += ((size_t)h_pages[i].n_bytes + 31) & ~size_t(15);
                enc_bytes += h_pages[i].n_bytes;
                RT_CUDA_OK(ctx, cudaMemcpyAsync(d, h_pages[i].rgb, (size_t)h_pages[i].n_bytes, cudaMemcpyHostToDevice, ctx->copy_stream));
                dev_pages[i] = retto_b200_page{d, infos[i].Y, infos[i].X, PAGE_DEVICE_ENCODED, h_pages[i].n_bytes};
                d_files[i] = d;
            }
            // entropy phase of ALL files behind the uploads, on the copy stream: its duration is the longest restart interval's
            // serial chain whatever the number of files, so it is paid once per call; the units' pixel phases wait for its event
            RT_TRY(rt_jpeg_entropy_enqueue(ctx, ctx->copy_stream, infos.data(), d_files.data(), n_pages));
            RT_CUDA_OK(ctx, cudaEventRecord(ctx->copy_events[0], ctx->copy_stream));
            std::vector<Unit> units;
            for (int u = 0; u < n_units; ++u) units.push_back(Unit{dev_pages.data() + u * unit, std::min(unit, n_pages - u * unit), ctx->copy_events[0], u * unit, infos.data() + u * unit});
            const retto_b200_status ret = run_units(ctx, units, n_lanes, forward, user, out);
            ctx->run_stats[7] = enc_bytes;
            cudaStreamSynchronize(ctx->copy_stream);
            return ret;
        }
    }
    bool all_host = n_pages > RUN_CHUNK_PAGES, all_dev = true;
    for (int i = 0; i < n_pages; ++i) {
        if (h_pages[i].on_device || !h_pages[i].rgb || h_pages[i].h <= 0 || h_pages[i].w <= 0) all_host = false;
        if (!h_pages[i].on_device) all_dev = false;
    }
    std::vector<Unit> units;
    if (!all_host) {
        // one unit, or (device-resident pages, two lanes) units of unit_dev pages
        const int up = (all_dev && n_lanes > 1 && n_pages > unit_dev) ? unit_dev : std::max(n_pages, 1);
        for (int p0 = 0; p0 < std::max(n_pages, 1); p0 += up) units.push_back(Unit{h_pages + p0, std::min(up, n_pages - p0), nullptr, p0, nullptr});
        return run_units(ctx, units, n_lanes, forward, user, out);
    }
    if (!ctx->copy_stream) {
        int lo = 0, hi = 0;
        cudaDeviceGetStreamPriorityRange(&lo, &hi);
        RT_CUDA_OK(ctx, cudaStreamCreateWithPriority(&ctx->copy_stream, cudaStreamNonBlocking, hi));   // page pulls first: they are the critical path
    }
    const int n_chunks = (n_pages + RUN_CHUNK_PAGES - 1) / RUN_CHUNK_PAGES;
    while ((int)ctx->copy_events.size() < n_chunks) {
        cudaEvent_t e;
        RT_CUDA_OK(ctx, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        ctx->copy_events.push_back(e);
    }
    size_t raw_bytes = 0;
    for (int i = 0; i < n_pages; ++i) raw_bytes += align256((size_t)h_pages[i].h * h_pages[i].w * 3);
    RT_CUDA_OK(ctx, ctx->d_pages_up.ensure(std::max<size_t>(raw_bytes, 256), ctx->stream));
    std::vector<retto_b200_page> dev_pages(n_pages);
    {
        size_t off = 0;
        PullArgs pa;
        pa.n = 0;
        bool pull_ok = true;
        for (int i = 0; i < n_pages; ++i) {
            uint8_t* d = ctx->d_pages_up.as<uint8_t>() + off;
            const size_t bytes = (size_t)h_pages[i].h * h_pages[i].w * 3;
            off += align256(bytes);
            dev_pages[i] = retto_b200_page{d, h_pages[i].h, h_pages[i].w, 1};
            cudaPointerAttributes at;
            const bool pinned = cudaPointerGetAttributes(&at, h_pages[i].rgb) == cudaSuccess && at.type == cudaMemoryTypeHost && at.devicePointer != nullptr &&
                                ((uintptr_t)at.devicePointer % 16 == 0);
            cudaGetLastError();
            if (pinned && pull_ok && !USE_DMA) {
                pa.src[pa.n] = reinterpret_cast<const uint4*>(at.devicePointer);
                pa.dst[pa.n] = reinterpret_cast<uint4*>(d);
                pa.n16[pa.n] = (unsigned)(bytes / 16);
                ++pa.n;
                if (bytes % 16) RT_CUDA_OK(ctx, cudaMemcpyAsync(d + (bytes & ~size_t(15)), h_pages[i].rgb + (bytes & ~size_t(15)), bytes % 16, cudaMemcpyHostToDevice, ctx->copy_stream));
            } else {
                pull_ok = false;   // pageable memory: plain DMA copy (no overlap guarantee)
                RT_CUDA_OK(ctx, cudaMemcpyAsync(d, h_pages[i].rgb, bytes, cudaMemcpyHostToDevice, ctx->copy_stream));
            }
            if ((i + 1) % RUN_CHUNK_PAGES == 0 || i + 1 == n_pages) {
                if (pa.n > 0) {
                    pull_pages_kernel<<<PULL_BLOCKS, 512, 0, ctx->copy_stream>>>(pa);
                    ctx->launches++;
                    pa.n = 0;
                }
                RT_CUDA_OK(ctx, cudaEventRecord(ctx->copy_events[i / RUN_CHUNK_PAGES], ctx->copy_stream));
            }
        }
    }
    for (int c = 0; c < n_chunks; ++c) {
        const int p0 = c * RUN_CHUNK_PAGES;
        units.push_back(Unit{dev_pages.data() + p0, std::min(RUN_CHUNK_PAGES, n_pages - p0), ctx->copy_events[c], p0, nullptr});
    }
    const bool by_sm = USE_DMA != 0;
    ctx->uploads_by_sm = by_sm;
    for (retto_b200_ctx* l : ctx->lanes) l->uploads_by_sm = by_sm;
    retto_b200_status ret = run_units(ctx, units, n_lanes, forward, user, out);
    ctx->uploads_by_sm = false;
    for (retto_b200_ctx* l : ctx->lanes) l->uploads_by_sm = false;
    cudaStreamSynchronize(ctx->copy_stream);
    return ret;
}

extern "C" retto_b200_status retto_b200_set_stage_callback(retto_b200_ctx* ctx, retto_b200_stage_fn fn, void* user) {
    if (!ctx) return RETTO_B200_ERR_INVALID_ARG;
    ctx->stage_cb = fn;
    ctx->stage_user = user;
    return RETTO_B200_OK;
}

extern "C" retto_b200_status retto_b200_set_pipeline(retto_b200_ctx* ctx, int32_t lanes, int32_t unit_pages) {
    if (!ctx || lanes < 0 || lanes > 2 || unit_pages < 0) return RETTO_B200_ERR_INVALID_ARG;
    ctx->pipe_lanes = lanes;
    ctx->pipe_unit_pages = unit_pages;
    return RETTO_B200_OK;
}

extern "C" retto_b200_status retto_b200_last_run_stats(const retto_b200_ctx* ctx, uint64_t* out8) {
    if (!ctx || !out8) return RETTO_B200_ERR_INVALID_ARG;
    memcpy(out8, ctx->run_stats, sizeof(ctx->run_stats));
    return RETTO_B200_OK;
}

// ---- several contexts, one batch: RettoSession::run over pages sharded per image ---------------------------------------
// Pages are independent (session.rs:75-106 reads no state of another page), so a batch is spread over N contexts — normally one per
// GPU of the box — with no collective: one host thread per context pulls chunks of pages from a shared atomic cursor over the page
// list sorted by descending H*W (longest-processing-time first, so the stragglers are the small pages), runs retto_b200_run_pages on
// its own context / device / stream, and files the results under the pages' original indices.  The assembled result (page order ==
// the sequential CLI's order) is owned by ctxs[0].
namespace {
struct PageOut {
    int32_t status = 0;
    std::vector<retto_b200_box> boxes;
    std::vector<retto_b200_cls_result> cls;
    std::vector<float> scores;
    std::vector<uint32_t> text_len;
    std::string text;
};
}  // namespace

extern "C" retto_b200_status retto_b200_run_pages_multi(retto_b200_ctx* const* ctxs, int32_t n_ctx, const retto_b200_page* h_pages, int32_t n_pages,
                                                        int32_t chunk_pages, retto_b200_forward_fn forward, void* const* users,
                                                        retto_b200_results* out) {
    if (!ctxs || n_ctx <= 0 || !ctxs[0] || n_pages < 0 || (n_pages > 0 && !h_pages) || !forward || !out || chunk_pages < 0) return RETTO_B200_ERR_INVALID_ARG;
    retto_b200_ctx* c0 = ctxs[0];
    memset(out, 0, sizeof(*out));
    for (int i = 0; i < n_ctx; ++i) {
        if (!ctxs[i]) { c0->set_error("run_pages_multi: null context"); return RETTO_B200_ERR_INVALID_ARG; }
        for (int j = 0; j < i; ++j) if (ctxs[j] == ctxs[i]) { c0->set_error("run_pages_multi: a context is listed twice"); return RETTO_B200_ERR_INVALID_ARG; }
    }
    for (int i = 0; i < n_pages; ++i)
        if (h_pages[i].on_device == RETTO_B200_PAGE_DEVICE_RGB) { c0->set_error("run_pages_multi: pages must be host-resident (a device pointer belongs to one GPU)"); return RETTO_B200_ERR_INVALID_ARG; }
    std::vector<int> order(n_pages);
    for (int i = 0; i < n_pages; ++i) order[i] = i;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) {
        return (long long)h_pages[a].h * h_pages[a].w > (long long)h_pages[b].h * h_pages[b].w;
    });
    const int chunk = chunk_pages > 0 ? chunk_pages : std::max(1, std::min(32, n_pages / (n_ctx * 3)));
    std::vector<PageOut> outs(n_pages);
    std::atomic<int> cursor{0};
    std::atomic<int> first_error{RETTO_B200_OK};
    std::vector<retto_b200_status> soft(n_ctx, RETTO_B200_OK);
    std::vector<uint64_t> stats((size_t)n_ctx * 8, 0);
    auto work = [&](int ci) {
        retto_b200_ctx* ctx = ctxs[ci];
        std::vector<retto_b200_page> mine;
        for (;;) {
            if (first_error.load() != RETTO_B200_OK) return;
            const int start = cursor.fetch_add(chunk);
            if (start >= n_pages) return;
            const int n = std::min(chunk, n_pages - start);
            mine.resize(n);
            for (int k = 0; k < n; ++k) mine[k] = h_pages[order[start + k]];
            retto_b200_results r;
            const retto_b200_status s = retto_b200_run_pages(ctx, mine.data(), n, forward, users ? users[ci] : nullptr, &r);
            const bool hard = s != RETTO_B200_OK && s != RETTO_B200_ERR_DEGENERATE_QUAD && s != RETTO_B200_ERR_CAPACITY;
            if (hard || r.n_pages != n) {
                int expect = RETTO_B200_OK;
                if (first_error.compare_exchange_strong(expect, hard ? (int)s : (int)RETTO_B200_ERR_CUDA) && ctx != c0) c0->set_error(ctx->err);
                return;
            }
            if (s != RETTO_B200_OK) soft[ci] = s;
            for (int k = 0; k < n; ++k) {
                PageOut& o = outs[order[start + k]];
                const retto_b200_page_result& pr = r.pages[k];
                o.status = pr.status;
                o.boxes.assign(r.boxes + pr.first_line, r.boxes + pr.first_line + pr.n_lines);
                o.cls.assign(r.cls + pr.first_line, r.cls + pr.first_line + pr.n_lines);
                o.scores.assign(r.rec_scores + pr.first_line, r.rec_scores + pr.first_line + pr.n_lines);
                o.text_len.resize(pr.n_lines);
                for (int l = 0; l < pr.n_lines; ++l) o.text_len[l] = r.text_offsets[pr.first_line + l + 1] - r.text_offsets[pr.first_line + l];
                if (pr.n_lines) o.text.assign(r.text + r.text_offsets[pr.first_line], r.text + r.text_offsets[pr.first_line + pr.n_lines]);
            }
            for (int k = 0; k < 8; ++k) stats[(size_t)ci * 8 + k] += ctx->run_stats[k];
        }
    };
    if (n_ctx == 1) work(0);
    else {
        std::vector<std::thread> th;
        for (int ci = 0; ci < n_ctx; ++ci) th.emplace_back(work, ci);
        for (auto& t : th) t.join();
    }
    if (first_error.load() != RETTO_B200_OK) return (retto_b200_status)first_error.load();
    retto_b200_status ret = RETTO_B200_OK;
    for (int ci = 0; ci < n_ctx; ++ci) if (soft[ci] != RETTO_B200_OK) { ret = soft[ci]; if (ctxs[ci] != c0) c0->set_error(ctxs[ci]->err); }
    // assemble in page order
    c0->r_pages.assign(n_pages, retto_b200_page_result{0, 0, 0});
    c0->r_boxes.clear(); c0->r_cls.clear(); c0->r_scores.clear(); c0->r_text_offs.assign(1, 0); c0->r_text.clear();
    for (int p = 0; p < n_pages; ++p) {
        const PageOut& o = outs[p];
        c0->r_pages[p] = retto_b200_page_result{o.status, (int32_t)c0->r_boxes.size(), (int32_t)o.boxes.size()};
        c0->r_boxes.insert(c0->r_boxes.end(), o.boxes.begin(), o.boxes.end());
        c0->r_cls.insert(c0->r_cls.end(), o.cls.begin(), o.cls.end());
        c0->r_scores.insert(c0->r_scores.end(), o.scores.begin(), o.scores.end());
        for (uint32_t l : o.text_len) c0->r_text_offs.push_back(c0->r_text_offs.back() + l);
        c0->r_text.insert(c0->r_text.end(), o.text.begin(), o.text.end());
    }
    c0->r_text.push_back(0);
    memset(c0->run_stats, 0, sizeof(c0->run_stats));
    for (int ci = 0; ci < n_ctx; ++ci) for (int k = 0; k < 8; ++k) c0->run_stats[k] += stats[(size_t)ci * 8 + k];
    fill_results(c0, out);
    return ret;
}
