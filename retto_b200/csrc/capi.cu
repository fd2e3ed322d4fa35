// capi.cu — context lifetime, memory helpers and configuration defaults of the C ABI
// (include/retto_b200.h).  Stage entry points live next to their kernels.
#include "common.cuh"

extern "C" int32_t retto_b200_abi_version(void) { return RETTO_B200_ABI_VERSION; }

extern "C" void retto_b200_config_default(retto_b200_config* c) {
    if (!c) return;
    memset(c, 0, sizeof(*c));
    c->max_side_len = 2000;                 // session.rs:33
    c->min_side_len = 30;                   // session.rs:34
    c->det_limit_side_len = 736;            // det_processor.rs:78
    c->det_limit_type = 0;                  // LimitType::Min
    for (int i = 0; i < 3; ++i) { c->det_mean[i] = 0.5f; c->det_std[i] = 0.5f; }
    c->det_scale = 1.0f / 255.0f;           // 1f32 / 255.0
    c->det_thresh = 0.3f;
    c->det_box_thresh = 0.5f;
    c->det_max_candidates = 1000;
    c->det_unclip_ratio = 1.6f;
    c->det_use_dilation = 1;
    c->det_score_mode = 1;                  // ScoreMode::Fast
    c->det_min_mini_box_size = 3;
    c->det_dilation_2x2 = 1;
    c->cls_image_shape[0] = 3; c->cls_image_shape[1] = 48; c->cls_image_shape[2] = 192;  // cls_processor.rs:30
    c->cls_batch_num = 6;
    c->cls_thresh = 0.9f;
    c->cls_label[0] = 0; c->cls_label[1] = 180;
    c->rec_image_shape[0] = 3; c->rec_image_shape[1] = 48; c->rec_image_shape[2] = 320;  // rec_processor.rs:132
    c->rec_batch_num = 6;
    c->max_components_per_page = 16384;
    c->max_det_side = 4096;
}

extern "C" retto_b200_status retto_b200_create(int32_t device_id, const retto_b200_config* cfg, retto_b200_ctx** out) {
    if (!out) return RETTO_B200_ERR_INVALID_ARG;
    *out = nullptr;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0 || device_id < 0 || device_id >= n) return RETTO_B200_ERR_CUDA;
    retto_b200_ctx* c = new retto_b200_ctx();
    c->device = device_id;
    RtDeviceGuard _dg(c);   // the caller's current device is restored on return
    if (!_dg.switched) {
        int cur = -1;
        if (cudaGetDevice(&cur) != cudaSuccess || cur != device_id) { delete c; return RETTO_B200_ERR_CUDA; }
    }
    if (cfg) c->cfg = *cfg; else retto_b200_config_default(&c->cfg);
    if (c->cfg.max_components_per_page <= 0) c->cfg.max_components_per_page = 16384;
    if (c->cfg.max_det_side <= 0) c->cfg.max_det_side = 4096;
    if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) { delete c; return RETTO_B200_ERR_CUDA; }
    *out = c;
    return RETTO_B200_OK;
}

extern "C" void retto_b200_destroy(retto_b200_ctx* c) {
    if (!c) return;
    for (retto_b200_ctx* l : c->lanes) retto_b200_destroy(l);
    c->lanes.clear();
    RtDeviceGuard _dg(c);
    cudaStreamSynchronize(c->stream);
    cudaStream_t s = c->stream;
    if (c->copy_stream) { cudaStreamSynchronize(c->copy_stream); cudaStreamDestroy(c->copy_stream); }
    for (cudaEvent_t e : c->copy_events) cudaEventDestroy(e);
    for (auto& t : c->timed) { cudaEventDestroy(t.a); cudaEventDestroy(t.b); }
    for (cudaEvent_t e : c->event_pool) cudaEventDestroy(e);
    if (c->ev_dp) cudaEventDestroy(c->ev_dp);
    if (c->ev_dp2) cudaEventDestroy(c->ev_dp2);
    if (c->aux_stream) { cudaStreamSynchronize(c->aux_stream); cudaStreamDestroy(c->aux_stream); }
    if (c->ev_fork) cudaEventDestroy(c->ev_fork);
    if (c->ev_join) cudaEventDestroy(c->ev_join);
    if (c->ev_jpeg_zero) cudaEventDestroy(c->ev_jpeg_zero);
    for (auto& sl : c->stage_slots) { if (sl.p) cudaFreeHost(sl.p); if (sl.ev) cudaEventDestroy(sl.ev); }
    delete c;  // DevBuf / HostBuf members free their memory
    cudaStreamDestroy(s);
}

extern "C" const char* retto_b200_last_error(const retto_b200_ctx* c) { return c ? c->err.c_str() : "null context"; }
extern "C" void* retto_b200_stream(retto_b200_ctx* c) { return c ? (void*)c->stream : nullptr; }
extern "C" uint64_t retto_b200_launch_count(const retto_b200_ctx* c) {
    if (!c) return 0;
    uint64_t n = c->launches;
    for (const retto_b200_ctx* l : c->lanes) n += l->launches;   // pipeline lanes of run_pages
    return n;
}

extern "C" retto_b200_status retto_b200_sync(retto_b200_ctx* c) {
    RtDeviceGuard _dg(c);
    if (!c) return RETTO_B200_ERR_INVALID_ARG;
    RT_CUDA_OK(c, cudaStreamSynchronize(c->stream));
    return RETTO_B200_OK;
}

extern "C" retto_b200_status retto_b200_dev_alloc(retto_b200_ctx* c, size_t bytes, void** d_out) {
    RtDeviceGuard _dg(c);
    if (!c || !d_out) return RETTO_B200_ERR_INVALID_ARG;
    cudaError_t e = cudaMalloc(d_out, bytes ? bytes : 1);
    if (e != cudaSuccess) { c->set_error(std::string("cudaMalloc: ") + cudaGetErrorString(e)); cudaGetLastError(); return RETTO_B200_ERR_OOM; }
    return RETTO_B200_OK;
}
extern "C" retto_b200_status retto_b200_dev_free(retto_b200_ctx* c, void* p) {
    RtDeviceGuard _dg(c);
    if (!c) return RETTO_B200_ERR_INVALID_ARG;
    RT_CUDA_OK(c, cudaStreamSynchronize(c->stream));
    RT_CUDA_OK(c, cudaFree(p));
    return RETTO_B200_OK;
}
extern "C" retto_b200_status retto_b200_host_alloc(retto_b200_ctx* c, size_t bytes, void** h_out) {
    RtDeviceGuard _dg(c);
    if (!c || !h_out) return RETTO_B200_ERR_INVALID_ARG;
    cudaError_t e = cudaHostAlloc(h_out, bytes ? bytes : 1, cudaHostAllocDefault);
    if (e != cudaSuccess) { c->set_error(std::string("cudaHostAlloc: ") + cudaGetErrorString(e)); cudaGetLastError(); return RETTO_B200_ERR_OOM; }
    return RETTO_B200_OK;
}
extern "C" retto_b200_status retto_b200_host_free(retto_b200_ctx* c, void* p) {
    RtDeviceGuard _dg(c);
    if (!c) return RETTO_B200_ERR_INVALID_ARG;
    RT_CUDA_OK(c, cudaFreeHost(p));
    return RETTO_B200_OK;
}
extern "C" retto_b200_status retto_b200_h2d(retto_b200_ctx* c, void* d_dst, const void* h_src, size_t bytes) {
    RtDeviceGuard _dg(c);
    if (!c) return RETTO_B200_ERR_INVALID_ARG;
    RT_CUDA_OK(c, cudaMemcpyAsync(d_dst, h_src, bytes, cudaMemcpyHostToDevice, c->stream));
    return RETTO_B200_OK;
}
extern "C" retto_b200_status retto_b200_d2h(retto_b200_ctx* c, void* h_dst, const void* d_src, size_t bytes) {
    RtDeviceGuard _dg(c);
    if (!c) return RETTO_B200_ERR_INVALID_ARG;
    RT_CUDA_OK(c, cudaMemcpyAsync(h_dst, d_src, bytes, cudaMemcpyDeviceToHost, c->stream));
    return RETTO_B200_OK;
}

// ---- per-kernel timing taps (bench.py roofline) ---------------------------------------------------------
extern "C" retto_b200_status retto_b200_enable_kernel_timing(retto_b200_ctx* c, int32_t on) {
    RtDeviceGuard _dg(c);
    if (!c) return RETTO_B200_ERR_INVALID_ARG;
    RT_CUDA_OK(c, cudaStreamSynchronize(c->stream));
    c->timer_collect();
    c->timing_enabled = on != 0;
    return RETTO_B200_OK;
}
extern "C" retto_b200_status retto_b200_reset_kernel_times(retto_b200_ctx* c) {
    RtDeviceGuard _dg(c);
    if (!c) return RETTO_B200_ERR_INVALID_ARG;
    RT_CUDA_OK(c, cudaStreamSynchronize(c->stream));
    c->timer_collect();
    for (auto& v : c->timer_total_ms) v = 0;
    for (auto& v : c->timer_count) v = 0;
    return RETTO_B200_OK;
}
// writes "name\tcount\ttotal_ms\n" lines into buf (NUL terminated); returns ERR_CAPACITY if it does not fit
extern "C" retto_b200_status retto_b200_kernel_times(retto_b200_ctx* c, char* buf, size_t cap) {
    RtDeviceGuard _dg(c);
    if (!c || !buf || cap == 0) return RETTO_B200_ERR_INVALID_ARG;
    RT_CUDA_OK(c, cudaStreamSynchronize(c->stream));
    c->timer_collect();
    std::string s;
    for (size_t i = 0; i < c->timer_names.size(); ++i) {
        char line[256];
        snprintf(line, sizeof(line), "%s\t%llu\t%.6f\n", c->timer_names[i].c_str(), (unsigned long long)c->timer_count[i], c->timer_total_ms[i]);
        s += line;
    }
    if (s.size() + 1 > cap) return RETTO_B200_ERR_CAPACITY;
    memcpy(buf, s.c_str(), s.size() + 1);
    return RETTO_B200_OK;
}

// a free pinned staging slot of at least `bytes`: never used, or the copy that read it has completed
retto_b200_status rt_stage_begin(retto_b200_ctx* ctx, size_t bytes, int* slot, void** p) {
    int pick = -1;
    for (size_t i = 0; i < ctx->stage_slots.size(); ++i) {
        auto& sl = ctx->stage_slots[i];
        if (sl.busy && cudaEventQuery(sl.ev) == cudaSuccess) sl.busy = false;
        if (!sl.busy && (pick < 0 || (ctx->stage_slots[pick].cap < bytes && sl.cap > ctx->stage_slots[pick].cap))) pick = (int)i;
    }
    cudaGetLastError();   // cudaErrorNotReady from the queries is not an error
    if (pick < 0) {
        retto_b200_ctx::StageSlot sl;
        RT_CUDA_OK(ctx, cudaEventCreateWithFlags(&sl.ev, cudaEventDisableTiming));
        ctx->stage_slots.push_back(sl);
        pick = (int)ctx->stage_slots.size() - 1;
    }
    auto& sl = ctx->stage_slots[pick];
    if (sl.cap < bytes) {
        if (sl.p) cudaFreeHost(sl.p);
        sl.p = nullptr; sl.cap = 0;
        const size_t want = std::max<size_t>(bytes + bytes / 2, 64 << 10);
        RT_CUDA_OK(ctx, cudaHostAlloc(&sl.p, want, cudaHostAllocDefault));
        sl.cap = want;
    }
    sl.busy = true;   // reserved until the commit's event completes
    *slot = pick;
    *p = sl.p;
    return RETTO_B200_OK;
}
// Descriptor upload by the SMs (pinned memory is device-accessible): used while run_pages streams host pages through the
// H2D copy engine, where a small cudaMemcpyAsync on the compute stream would queue behind megabytes of page copies in
// the same engine FIFO and stall the kernels waiting for their tables.
__global__ void __launch_bounds__(256) stage_pull_kernel(unsigned char* __restrict__ dst, const unsigned char* __restrict__ src, size_t bytes) {
    const size_t n16 = bytes / 16;
    const uint4* __restrict__ s4 = reinterpret_cast<const uint4*>(src);
    uint4* __restrict__ d4 = reinterpret_cast<uint4*>(dst);
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) d4[i] = s4[i];
    for (size_t i = n16 * 16 + blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < bytes; i += (size_t)gridDim.x * blockDim.x) dst[i] = src[i];
}
static retto_b200_status stage_copy(retto_b200_ctx* ctx, void* d_dst, const void* h_pinned, size_t bytes) {
    if (ctx->uploads_by_sm && ((uintptr_t)d_dst % 16 == 0)) {
        void* dp = nullptr;
        RT_CUDA_OK(ctx, cudaHostGetDevicePointer(&dp, const_cast<void*>(h_pinned), 0));
        const unsigned blocks = (unsigned)std::min<size_t>(8, (bytes + 16383) / 16384);
        stage_pull_kernel<<<std::max(1u, blocks), 256, 0, ctx->stream>>>(reinterpret_cast<unsigned char*>(d_dst), reinterpret_cast<const unsigned char*>(dp), bytes);
        ctx->launches++;
        RT_CUDA_OK(ctx, cudaGetLastError());
        return RETTO_B200_OK;
    }
    RT_CUDA_OK(ctx, cudaMemcpyAsync(d_dst, h_pinned, bytes, cudaMemcpyHostToDevice, ctx->stream));
    return RETTO_B200_OK;
}
// enqueue the H2D copy of a slot filled in place (no intermediate host copy)
retto_b200_status rt_stage_commit(retto_b200_ctx* ctx, DevBuf& dst, int slot, size_t bytes) {
    auto& sl = ctx->stage_slots[slot];
    RT_CUDA_OK(ctx, dst.ensure(bytes ? bytes : 16, ctx->stream));
    if (bytes) RT_TRY(stage_copy(ctx, dst.p, sl.p, bytes));
    RT_CUDA_OK(ctx, cudaEventRecord(sl.ev, ctx->stream));
    return RETTO_B200_OK;
}
retto_b200_status rt_upload_to(retto_b200_ctx* ctx, void* d_dst, const void* src, size_t bytes) {
    if (!bytes) return RETTO_B200_OK;
    int slot = -1;
    void* p = nullptr;
    RT_TRY(rt_stage_begin(ctx, bytes, &slot, &p));
    memcpy(p, src, bytes);
    RT_TRY(stage_copy(ctx, d_dst, p, bytes));
    RT_CUDA_OK(ctx, cudaEventRecord(ctx->stage_slots[slot].ev, ctx->stream));
    return RETTO_B200_OK;
}
retto_b200_status rt_upload(retto_b200_ctx* ctx, DevBuf& dst, const void* src, size_t bytes) {
    if (!bytes) { RT_CUDA_OK(ctx, dst.ensure(16, ctx->stream)); return RETTO_B200_OK; }
    int slot = -1;
    void* p = nullptr;
    RT_TRY(rt_stage_begin(ctx, bytes, &slot, &p));
    memcpy(p, src, bytes);
    return rt_stage_commit(ctx, dst, slot, bytes);
}
