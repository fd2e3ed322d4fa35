// session.cu — RettoSession::process_pipeline (session.rs:75-106) for a batch of pages, host side in C++
// above the stage entry points.  Order of operations per page is the reference's:
//   resize_both -> det.process (preprocess, worker.det, postprocess) -> get_crop_img per box ->
//   boxes.scale_and_clip to the original image -> cls.process (may flip crops) -> rec.process
// with every stage executed once for the whole batch (one launch per kernel, not per page).  Pages never
// touch host memory between stages; the forward passes are the caller's (worker.rs:69-73 seam).
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

#include <chrono>
#define RUN_CHUNK_PAGES_MAX 64
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

static retto_b200_status run_pages_once(retto_b200_ctx* ctx, const retto_b200_page* h_pages, int32_t n_pages,
                                        retto_b200_forward_fn forward, void* user, retto_b200_results* out) {
    if (!ctx || n_pages < 0 || (n_pages > 0 && !h_pages) || !forward || !out) return RETTO_B200_ERR_INVALID_ARG;
    cudaStream_t st = ctx->stream;
    HostTrace tr;
    const retto_b200_config& cfg = ctx->cfg;
    ctx->r_pages.assign(n_pages, retto_b200_page_result{0, 0, 0});
    ctx->r_boxes.clear(); ctx->r_cls.clear(); ctx->r_text_offs.assign(1, 0); ctx->r_text.clear(); ctx->r_scores.clear();
    memset(out, 0, sizeof(*out));
    out->n_pages = n_pages;
    out->pages = ctx->r_pages.data();
    out->text_offsets = ctx->r_text_offs.data();
    if (n_pages == 0) return RETTO_B200_OK;

    // ---- 1. pages to the device, resize_both (image_helper.rs:106-148) ------------------------------------
    std::vector<PageState> ps(n_pages);
    size_t raw_bytes = 0, rs_bytes = 0, det_floats = 0;
    std::vector<std::vector<std::pair<int, int>>> rs_steps(n_pages);
    for (int i = 0; i < n_pages; ++i) {
        const retto_b200_page& p = h_pages[i];
        if (!p.rgb || p.h <= 0 || p.w <= 0) { ctx->set_error("run_pages: bad page " + std::to_string(i)); return RETTO_B200_ERR_INVALID_ARG; }
        ps[i].ori_h = p.h; ps[i].ori_w = p.w;
        if (!p.on_device) raw_bytes += align256((size_t)p.h * p.w * 3);
        int dims[4], ns = 0;
        RT_TRY(retto_b200_resize_both_plan(p.h, p.w, cfg.max_side_len, cfg.min_side_len, dims, &ns));
        int h = p.h, w = p.w;
        for (int s = 0; s < ns; ++s) {
            h = dims[2 * s]; w = dims[2 * s + 1];
            if (h <= 0 || w <= 0) { ctx->set_error("run_pages: page " + std::to_string(i) + " resizes to nothing"); return RETTO_B200_ERR_INVALID_ARG; }
            rs_steps[i].push_back({h, w});
            rs_bytes += align256((size_t)h * w * 3);
        }
        ps[i].h = h; ps[i].w = w;
        RT_TRY(retto_b200_resize_either_plan(h, w, cfg.det_limit_type, cfg.det_limit_side_len, &ps[i].det_h, &ps[i].det_w));
        if (ps[i].det_h <= 0 || ps[i].det_w <= 0 || ps[i].det_h > cfg.max_det_side || ps[i].det_w > cfg.max_det_side) {
            ctx->set_error("run_pages: det tensor of page " + std::to_string(i) + " is " + std::to_string(ps[i].det_h) + "x" + std::to_string(ps[i].det_w) +
                           " (cap max_det_side=" + std::to_string(cfg.max_det_side) + ")");
            return RETTO_B200_ERR_CAPACITY;
        }
        det_floats += (align256((size_t)3 * ps[i].det_h * ps[i].det_w * 4)) / 4;
    }
    RT_CUDA_OK(ctx, ctx->d_pages_raw.ensure(std::max<size_t>(raw_bytes, 256), st));
    RT_CUDA_OK(ctx, ctx->d_pages_rs.ensure(std::max<size_t>(rs_bytes, 256), st));
    RT_CUDA_OK(ctx, ctx->d_det_in.ensure(std::max<size_t>(det_floats * 4, 256), st));
    {
        size_t off = 0, roff = 0;
        std::vector<retto_b200_resize_desc> step1, step2;
        for (int i = 0; i < n_pages; ++i) {
            const retto_b200_page& p = h_pages[i];
            const uint8_t* cur = p.rgb;
            if (!p.on_device) {
                uint8_t* d = ctx->d_pages_raw.as<uint8_t>() + off;
                RT_CUDA_OK(ctx, cudaMemcpyAsync(d, p.rgb, (size_t)p.h * p.w * 3, cudaMemcpyHostToDevice, st));
                off += align256((size_t)p.h * p.w * 3);
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
        if (!step1.empty()) RT_TRY(retto_b200_thumbnail(ctx, step1.data(), (int)step1.size()));
        if (!step2.empty()) RT_TRY(retto_b200_thumbnail(ctx, step2.data(), (int)step2.size()));
    }

    tr.mark("upload+resize");
    // ---- 2. det preprocess (det_processor.rs:256-274) --------------------------------------------------------
    std::vector<retto_b200_tensor> det_in(n_pages), det_out(n_pages);
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
    if (forward(user, 0, n_pages, det_in.data(), det_out.data(), (void*)st) != 0) { ctx->set_error("run_pages: det forward failed"); return RETTO_B200_ERR_WORKER; }
    tr.mark("det_fwd");
    // ---- 4. det postprocess (det_processor.rs:279-335) ----------------------------------------------------------
    std::vector<int32_t> page_status(n_pages, 0), box_off(n_pages + 1, 0);
    {
        std::vector<retto_b200_det_post_desc> descs(n_pages);
        for (int i = 0; i < n_pages; ++i) {
            const retto_b200_tensor& t = det_out[i];
            if (!t.d_data || t.ndim != 4 || t.shape[0] != 1 || t.shape[1] != 1 || t.shape[2] <= 0 || t.shape[3] <= 0) {
                ctx->set_error("run_pages: det forward returned a bad tensor for page " + std::to_string(i));
                return RETTO_B200_ERR_WORKER;
            }
            // DetProcessor::new(cfg, after_h, after_w): boxes are scaled to the resize_both-ed page (session.rs:85)
            descs[i] = retto_b200_det_post_desc{t.d_data, (int32_t)t.shape[2], (int32_t)t.shape[3], ps[i].h, ps[i].w};
        }
        int cap = std::max(4096, n_pages * 256);
        for (int attempt = 0; attempt < 2; ++attempt) {
            ctx->r_boxes.resize(cap);
            retto_b200_status s = retto_b200_det_postprocess(ctx, descs.data(), n_pages, page_status.data(), box_off.data(), ctx->r_boxes.data(), cap);
            if (s == RETTO_B200_ERR_CAPACITY && box_off[n_pages] > cap && attempt == 0) { cap = box_off[n_pages]; continue; }
            if (s != RETTO_B200_OK) return s;
            break;
        }
    }
    tr.mark("det_post");
    const int n_lines = box_off[n_pages];
    ctx->r_boxes.resize(n_lines);
    retto_b200_status ret = RETTO_B200_OK;
    for (int i = 0; i < n_pages; ++i) {
        ctx->r_pages[i].status = page_status[i];
        ctx->r_pages[i].first_line = box_off[i];
        ctx->r_pages[i].n_lines = box_off[i + 1] - box_off[i];
        if (page_status[i] != RETTO_B200_OK) { ret = (retto_b200_status)page_status[i]; ctx->set_error("run_pages: det postprocess status on page " + std::to_string(i)); }
    }
    ctx->r_cls.assign(n_lines, retto_b200_cls_result{0, 0.0f});
    ctx->r_scores.assign(n_lines, 0.0f);
    ctx->r_text_offs.assign(n_lines + 1, 0);
    out->n_lines = n_lines;
    out->boxes = ctx->r_boxes.data();
    out->cls = ctx->r_cls.data();
    out->text_offsets = ctx->r_text_offs.data();
    out->rec_scores = ctx->r_scores.data();
    out->text = ctx->r_text.data();
    if (n_lines == 0) return ret;

    // ---- 5. crops from the resize_both-ed page (session.rs:88-92) --------------------------------------------------
    std::vector<retto_b200_crop_info> infos(n_lines);
    {
        std::vector<const uint8_t*> pp(n_pages);
        std::vector<int> ph(n_pages), pw(n_pages);
        for (int i = 0; i < n_pages; ++i) { pp[i] = ps[i].d_img; ph[i] = ps[i].h; pw[i] = ps[i].w; }
        // async: the host plans the batches meanwhile
        RT_TRY(rt_crop_launch_pages(ctx, ctx->r_boxes.data(), box_off.data(), n_pages, pp.data(), ph.data(), pw.data(), infos.data()));
    }
    tr.mark("crop_launch");
    // ---- 7. cls (cls_processor.rs:127-172) ------------------------------------------------------------------------------
    auto plan_all = [&](int kind, std::vector<retto_b200_line_job>& lines, std::vector<retto_b200_batch>& batches, std::vector<int>& batch_page,
                        uint64_t* total) -> retto_b200_status {
        lines.assign(n_lines, retto_b200_line_job{});
        batches.clear(); batch_page.clear();
        uint64_t off = 0;
        std::vector<retto_b200_batch> pb;
        for (int i = 0; i < n_pages; ++i) {
            const int nb = box_off[i + 1] - box_off[i];
            if (nb == 0) continue;
            pb.assign((size_t)nb, retto_b200_batch{});
            int32_t nbat = 0; uint64_t tot = 0;
            RT_TRY(retto_b200_plan_batches(&cfg, kind, infos.data() + box_off[i], nb, lines.data() + box_off[i], pb.data(), &nbat, &tot));
            for (int k = box_off[i]; k < box_off[i + 1]; ++k) { lines[k].crop += box_off[i]; lines[k].dst_offset += off; }
            for (int b = 0; b < nbat; ++b) { pb[b].first_line += box_off[i]; pb[b].offset += off; batches.push_back(pb[b]); batch_page.push_back(i); }
            off += tot;
        }
        *total = off;
        return RETTO_B200_OK;
    };
    std::vector<retto_b200_line_job> lines, rec_lines;
    std::vector<retto_b200_batch> batches, rec_batches;
    std::vector<int> batch_page, rec_batch_page;
    uint64_t total = 0, rec_total = 0;
    // From here to the CTC read-back nothing waits for the GPU: the plans depend only on the crop dims (known on the
    // host), descriptor uploads go through pinned staging slots, and the crop statuses / cls results / rescaled boxes
    // are collected after the one final sync — the host runs ahead and the kernels queue back to back.
    RT_TRY(plan_all(0, lines, batches, batch_page, &total));
    tr.mark("crops+plans");
    // ---- 6. boxes back to original-image coordinates (session.rs:94-97) ----------------------------------------------
    bool scaled = false;
    {
        bool any = false;
        std::vector<double> prm((size_t)n_lines * 4);
        for (int i = 0; i < n_pages; ++i) {
            if (ps[i].h != ps[i].ori_h || ps[i].w != ps[i].ori_w) any = true;  // identity otherwise: round(x * 1) clamped == x
            for (int k = box_off[i]; k < box_off[i + 1]; ++k) {
                prm[4 * (size_t)k] = (double)ps[i].ori_w / (double)ps[i].w;
                prm[4 * (size_t)k + 1] = (double)ps[i].ori_h / (double)ps[i].h;
                prm[4 * (size_t)k + 2] = (double)ps[i].ori_w;
                prm[4 * (size_t)k + 3] = (double)ps[i].ori_h;
            }
        }
        if (any) { RT_TRY(rt_scale_and_clip_multi(ctx, ctx->r_boxes.data(), prm.data(), n_lines, true)); scaled = true; }
    }

    tr.mark("scale");
    float *d_base = nullptr, *d_base_rec = nullptr;
    RT_TRY(rt_build_batches_prepare(ctx, 0, lines.data(), n_lines, total, &d_base));
    RT_TRY(rt_build_batches_launch(ctx, 0));
    tr.mark("cls_build");
    // host work while the GPU crops and builds the cls batches: the rec plan and the rec descriptor tables
    RT_TRY(plan_all(1, rec_lines, rec_batches, rec_batch_page, &rec_total));
    RT_TRY(rt_build_batches_prepare(ctx, 1, rec_lines.data(), n_lines, rec_total, &d_base_rec));
    {
        uint64_t det_px = 0, crop_px = 0, rec_rows = 0;
        for (int i = 0; i < n_pages; ++i) det_px += (uint64_t)ps[i].det_h * ps[i].det_w;
        for (int k = 0; k < n_lines; ++k) crop_px += (uint64_t)infos[k].w * infos[k].h;
        for (const auto& b : rec_batches) rec_rows += (uint64_t)b.n * (b.img_w / 8);
        const uint64_t st8[8] = {(uint64_t)n_pages, (uint64_t)n_lines, det_px, crop_px, total, rec_total, rec_rows, 0};
        memcpy(ctx->run_stats, st8, sizeof(st8));
    }
    tr.mark("rec_prepare");
    std::vector<int32_t> cls_crop_idx;
    std::vector<retto_b200_tensor> tin(batches.size()), tout(batches.size());
    auto fill_inputs = [&](int img_h) {
        for (size_t b = 0; b < batches.size(); ++b) {
            tin[b].d_data = d_base + batches[b].offset;
            tin[b].shape[0] = batches[b].n; tin[b].shape[1] = 3; tin[b].shape[2] = img_h; tin[b].shape[3] = batches[b].img_w;
            tin[b].ndim = 4;
            memset(&tout[b], 0, sizeof(retto_b200_tensor));
        }
    };
    fill_inputs(cfg.cls_image_shape[1]);
    if (forward(user, 1, (int)batches.size(), tin.data(), tout.data(), (void*)st) != 0) { ctx->set_error("run_pages: cls forward failed"); return RETTO_B200_ERR_WORKER; }
    {
        std::vector<const float*> ptrs(n_lines);
        std::vector<int32_t> crop_idx(n_lines);
        std::vector<retto_b200_cls_result> res(n_lines);
        for (size_t b = 0; b < batches.size(); ++b) {
            const retto_b200_tensor& t = tout[b];
            if (!t.d_data || t.ndim != 2 || t.shape[0] != batches[b].n || t.shape[1] != 2) {
                ctx->set_error("run_pages: cls forward returned a bad tensor");
                return RETTO_B200_ERR_WORKER;
            }
            for (int k = 0; k < batches[b].n; ++k) {
                ptrs[batches[b].first_line + k] = t.d_data + 2 * (size_t)k;
                crop_idx[batches[b].first_line + k] = lines[batches[b].first_line + k].crop;
            }
        }
        RT_TRY(rt_cls_postprocess_ptrs(ctx, ptrs, crop_idx.data(), n_lines, nullptr, true));   // results collected after the final sync
        cls_crop_idx = crop_idx;
    }

    tr.mark("cls_fwd+post");
    // ---- 8. rec (rec_processor.rs:214-270) ---------------------------------------------------------------------------------
    lines.swap(rec_lines); batches.swap(rec_batches); total = rec_total; d_base = d_base_rec;
    RT_TRY(rt_build_batches_launch(ctx, 1));
    tin.assign(batches.size(), retto_b200_tensor{});
    tout.assign(batches.size(), retto_b200_tensor{});
    tr.mark("rec_build");
    fill_inputs(cfg.rec_image_shape[1]);
    if (forward(user, 2, (int)batches.size(), tin.data(), tout.data(), (void*)st) != 0) { ctx->set_error("run_pages: rec forward failed"); return RETTO_B200_ERR_WORKER; }
    {
        std::vector<retto_b200_logits_desc> descs(batches.size());
        int max_t = 1;
        for (size_t b = 0; b < batches.size(); ++b) {
            const retto_b200_tensor& t = tout[b];
            if (!t.d_data || t.ndim != 3 || t.shape[0] != batches[b].n || t.shape[1] <= 0 || t.shape[2] != (int64_t)ctx->dict.size()) {
                ctx->set_error("run_pages: rec forward returned a bad tensor (classes must equal the dictionary size " + std::to_string(ctx->dict.size()) + ")");
                return ctx->dict.empty() ? RETTO_B200_ERR_NO_DICT : RETTO_B200_ERR_WORKER;
            }
            descs[b] = retto_b200_logits_desc{t.d_data, (int32_t)t.shape[0], (int32_t)t.shape[1]};
            max_t = std::max(max_t, (int)t.shape[1]);
        }
        std::vector<uint32_t> toff(n_lines + 1, 0);
        std::vector<float> sc(n_lines, 0.0f);
        std::vector<char> text((size_t)n_lines * max_t * std::max(ctx->dict_max_len, 1) + 16);
        retto_b200_status s = retto_b200_ctc_decode(ctx, descs.data(), (int)descs.size(), (int)ctx->dict.size(), toff.data(), text.data(), text.size(),
                                                    sc.data(), nullptr, nullptr, 0);
        if (s != RETTO_B200_OK) return s;
        {   // the CTC call synchronised the stream: the deferred cls results, crop statuses and rescaled boxes are on the host now
            std::vector<retto_b200_cls_result> res(n_lines);
            RT_TRY(rt_cls_collect(ctx, n_lines, res.data()));
            for (int k = 0; k < n_lines; ++k) ctx->r_cls[cls_crop_idx[k]] = res[k];  // final_res[idx].label = label (cls_processor.rs:167)
            if (scaled) memcpy(ctx->r_boxes.data(), ctx->h_scale.p, sizeof(retto_b200_box) * (size_t)n_lines);
            RT_TRY(rt_crop_finish(ctx, infos.data(), false));
        }
        // scatter from plan order back to detection order (rec_processor.rs:259-264)
        std::vector<uint32_t> len(n_lines, 0);
        for (int k = 0; k < n_lines; ++k) len[lines[k].crop] = toff[k + 1] - toff[k];
        for (int k = 0; k < n_lines; ++k) ctx->r_text_offs[k + 1] = ctx->r_text_offs[k] + len[k];
        ctx->r_text.assign(ctx->r_text_offs[n_lines] + 1, 0);
        for (int k = 0; k < n_lines; ++k) {
            const int dst = lines[k].crop;
            memcpy(ctx->r_text.data() + ctx->r_text_offs[dst], text.data() + toff[k], toff[k + 1] - toff[k]);
            ctx->r_scores[dst] = sc[k];
        }
    }
    tr.mark("rec_fwd+ctc");
    out->text = ctx->r_text.data();
    out->text_offsets = ctx->r_text_offs.data();
    return ret;
}

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

// Host-resident batches larger than one chunk are pipelined: every page is uploaded on a separate copy stream up front
// (one event per chunk), and chunk k is processed on the compute stream as soon as its pages have landed, so the PCIe
// transfer of the later chunks overlaps with the kernels of the earlier ones.  Results are concatenated in page order.
extern "C" retto_b200_status retto_b200_run_pages(retto_b200_ctx* ctx, const retto_b200_page* h_pages, int32_t n_pages,
                                                  retto_b200_forward_fn forward, void* user, retto_b200_results* out) {
    if (!ctx || n_pages < 0 || (n_pages > 0 && !h_pages) || !forward || !out) return RETTO_B200_ERR_INVALID_ARG;
    // pages per chunk / blocks of the pull kernel: tunables for experiments, defaults measured on B200
    static const int RUN_CHUNK_PAGES = env_int("RETTO_B200_CHUNK_PAGES", 32, 1, RUN_CHUNK_PAGES_MAX);
    static const int PULL_BLOCKS = env_int("RETTO_B200_PULL_BLOCKS", 16, 1, 1024);
    static const int USE_DMA = env_int("RETTO_B200_PULL_DMA", 0, 0, 1);
    bool all_host = n_pages > RUN_CHUNK_PAGES;
    for (int i = 0; i < n_pages && all_host; ++i) if (h_pages[i].on_device || !h_pages[i].rgb || h_pages[i].h <= 0 || h_pages[i].w <= 0) all_host = false;
    if (!all_host) return run_pages_once(ctx, h_pages, n_pages, forward, user, out);

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
    std::vector<retto_b200_page_result> a_pages;
    std::vector<retto_b200_box> a_boxes;
    std::vector<retto_b200_cls_result> a_cls;
    std::vector<uint32_t> a_toffs(1, 0);
    std::vector<char> a_text;
    std::vector<float> a_scores;
    uint64_t stats[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    retto_b200_status ret = RETTO_B200_OK;
    for (int c = 0; c < n_chunks; ++c) {
        const int p0 = c * RUN_CHUNK_PAGES, np = std::min(RUN_CHUNK_PAGES, n_pages - p0);
        RT_CUDA_OK(ctx, cudaStreamWaitEvent(ctx->stream, ctx->copy_events[c], 0));
        retto_b200_results r;
        retto_b200_status s = run_pages_once(ctx, dev_pages.data() + p0, np, forward, user, &r);
        if (s != RETTO_B200_OK && s != RETTO_B200_ERR_DEGENERATE_QUAD && s != RETTO_B200_ERR_CAPACITY) { cudaStreamSynchronize(ctx->copy_stream); return s; }
        if (s != RETTO_B200_OK) ret = s;
        const int line0 = (int)a_boxes.size();
        for (int i = 0; i < r.n_pages; ++i) {
            retto_b200_page_result pr = r.pages[i];
            pr.first_line += line0;
            a_pages.push_back(pr);
        }
        a_boxes.insert(a_boxes.end(), r.boxes, r.boxes + r.n_lines);
        a_cls.insert(a_cls.end(), r.cls, r.cls + r.n_lines);
        a_scores.insert(a_scores.end(), r.rec_scores, r.rec_scores + r.n_lines);
        const uint32_t t0 = a_toffs.back();
        for (int k = 0; k < r.n_lines; ++k) a_toffs.push_back(t0 + r.text_offsets[k + 1]);
        if (r.n_lines) a_text.insert(a_text.end(), r.text, r.text + r.text_offsets[r.n_lines]);
        for (int k = 0; k < 8; ++k) stats[k] += ctx->run_stats[k];
    }
    RT_CUDA_OK(ctx, cudaStreamSynchronize(ctx->copy_stream));
    memcpy(ctx->run_stats, stats, sizeof(stats));
    a_text.push_back(0);
    ctx->r_pages.swap(a_pages); ctx->r_boxes.swap(a_boxes); ctx->r_cls.swap(a_cls);
    ctx->r_text_offs.swap(a_toffs); ctx->r_text.swap(a_text); ctx->r_scores.swap(a_scores);
    out->n_pages = n_pages;
    out->pages = ctx->r_pages.data();
    out->n_lines = (int)ctx->r_boxes.size();
    out->boxes = ctx->r_boxes.data();
    out->cls = ctx->r_cls.data();
    out->text_offsets = ctx->r_text_offs.data();
    out->text = ctx->r_text.data();
    out->rec_scores = ctx->r_scores.data();
    return ret;
}

extern "C" retto_b200_status retto_b200_last_run_stats(const retto_b200_ctx* ctx, uint64_t* out8) {
    if (!ctx || !out8) return RETTO_B200_ERR_INVALID_ARG;
    memcpy(out8, ctx->run_stats, sizeof(ctx->run_stats));
    return RETTO_B200_OK;
}
