// ctc_decode.cu — K10: rec postprocess (argmax + max over the class axis) and CTC greedy decode.
// Replaces RecProcessor::postprocess (rec_processor.rs:190-208: two full scalar passes, argmax and
// max) and RecCharacter::decode (rec_processor.rs:48-97).
//
// Roofline: HBM.  Algorithmic bytes = 4 * rows * C (each logit read once).  One warp per (line, t)
// row: 128-bit streaming loads (row starts are only 4-byte aligned: 26500 B rows for C = 6625, so
// up to 3 head elements are peeled), per-lane running (value, index) with the reference's
// first-maximum rule, warp-shuffle reduction "greater value wins, lower index on ties".
#include <math_constants.h>

#include "common.cuh"

struct CtcTensor {
    const float* logits;  // [n, t, C]
    int n, t;
    int line_base;        // first line index of this tensor
    int pad;
};

__device__ __forceinline__ void amax_upd(float v, int i, float& bv, int& bi, int& nan) {
    nan |= (v != v);
    if (v > bv || (v == bv && i < bi)) { bv = v; bi = i; }
}

// rows_prefix[k] = number of rows before tensor k.  idx/prob are laid out [line][max_t].
template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32) ctc_argmax_kernel(const CtcTensor* __restrict__ tensors, const int* __restrict__ rows_prefix,
                                                                 int n_tensors, int total_rows, int C, int max_t,
                                                                 int* __restrict__ idx_out, float* __restrict__ prob_out,
                                                                 int* __restrict__ line_nan) {
    const int lane = threadIdx.x & 31;
    const int row_g = blockIdx.x * WARPS + (threadIdx.x >> 5);
    if (row_g >= total_rows) return;
    const int k = rt_find_segment(rows_prefix, n_tensors, row_g);
    const CtcTensor tn = tensors[k];
    const int r = row_g - rows_prefix[k];
    const float* __restrict__ row = tn.logits + (size_t)r * (size_t)C;

    float bv = -CUDART_INF_F;
    int bi = 0x7fffffff;
    int nan = 0;

    int head = (int)(((16u - (unsigned)((uintptr_t)row & 15u)) & 15u) >> 2);
    if (head > C) head = C;
    if (lane < head) amax_upd(__ldcs(row + lane), lane, bv, bi, nan);
    const float4* __restrict__ body = reinterpret_cast<const float4*>(row + head);
    const int nvec = (C - head) >> 2;
    int i = lane;
    // 4 independent 128-bit loads in flight per lane
    for (; i + 96 < nvec; i += 128) {
        const float4 a = __ldcs(body + i);
        const float4 b = __ldcs(body + i + 32);
        const float4 c = __ldcs(body + i + 64);
        const float4 d = __ldcs(body + i + 96);
        int base = head + 4 * i;
        amax_upd(a.x, base, bv, bi, nan); amax_upd(a.y, base + 1, bv, bi, nan); amax_upd(a.z, base + 2, bv, bi, nan); amax_upd(a.w, base + 3, bv, bi, nan);
        base += 128;
        amax_upd(b.x, base, bv, bi, nan); amax_upd(b.y, base + 1, bv, bi, nan); amax_upd(b.z, base + 2, bv, bi, nan); amax_upd(b.w, base + 3, bv, bi, nan);
        base += 128;
        amax_upd(c.x, base, bv, bi, nan); amax_upd(c.y, base + 1, bv, bi, nan); amax_upd(c.z, base + 2, bv, bi, nan); amax_upd(c.w, base + 3, bv, bi, nan);
        base += 128;
        amax_upd(d.x, base, bv, bi, nan); amax_upd(d.y, base + 1, bv, bi, nan); amax_upd(d.z, base + 2, bv, bi, nan); amax_upd(d.w, base + 3, bv, bi, nan);
    }
    for (; i < nvec; i += 32) {
        const float4 a = __ldcs(body + i);
        const int base = head + 4 * i;
        amax_upd(a.x, base, bv, bi, nan); amax_upd(a.y, base + 1, bv, bi, nan); amax_upd(a.z, base + 2, bv, bi, nan); amax_upd(a.w, base + 3, bv, bi, nan);
    }
    for (int j = head + 4 * nvec + lane; j < C; j += 32) amax_upd(__ldcs(row + j), j, bv, bi, nan);

#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, bv, off);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, off);
        nan |= __shfl_xor_sync(0xffffffffu, nan, off);
        if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
    }
    if (lane == 0) {
        const int line = tn.line_base + r / tn.t;
        const int t = r - (r / tn.t) * tn.t;
        idx_out[(size_t)line * max_t + t] = bi;
        prob_out[(size_t)line * max_t + t] = bv;
        if (nan) line_nan[line] = 1;
    }
}

// one warp per line: CTC collapse (rec_processor.rs:57-95): keep t iff idx != 0 && (t == 0 || idx[t]
// != idx[t-1]) && idx not in ignored_tokens ([0]); score = sum(p) / count, sequential f32 (NaN when
// count == 0).  Text = concatenation of dict[idx] written at a fixed per-line stride.
// 32 time steps per trip: the keep mask is a ballot, token / text positions are prefix counts over it, and the score
// is still the reference's sequential fold — the kept probabilities are added one by one in time order (a shuffle
// broadcast per kept step, every lane carrying the same accumulator).
__global__ void __launch_bounds__(256) ctc_collapse_kernel(const int* __restrict__ idx, const float* __restrict__ prob, const int* __restrict__ line_t,
                                    int n_lines, int max_t, const unsigned* __restrict__ dict_offs,
                                    const unsigned char* __restrict__ dict_bytes, int n_dict, int text_stride, int* __restrict__ tokens,
                                    int* __restrict__ counts, float* __restrict__ scores, unsigned char* __restrict__ text,
                                    int* __restrict__ text_len) {
    const int line = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (line >= n_lines) return;
    const int T = line_t[line];
    const int* li = idx + (size_t)line * max_t;
    const float* lp = prob + (size_t)line * max_t;
    int* tk = tokens + (size_t)line * max_t;
    unsigned char* tx = text + (size_t)line * text_stride;
    float acc = 0.0f;
    int cnt = 0, tl = 0, carry = -1;
    for (int base = 0; base < T; base += 32) {
        const int t = base + lane;
        const int c = t < T ? li[t] : 0;
        const float p = t < T ? lp[t] : 0.0f;
        int prev = __shfl_up_sync(0xffffffffu, c, 1);
        if (lane == 0) prev = carry;
        carry = __shfl_sync(0xffffffffu, c, 31);
        const bool sel = (t < T) && (c != 0) && (t == 0 || c != prev);
        const unsigned m = __ballot_sync(0xffffffffu, sel);
        const unsigned below = m & ((1u << lane) - 1u);
        unsigned b = 0, e = 0;
        if (sel) {
            tk[cnt + __popc(below)] = c;
            if (c < n_dict) { b = dict_offs[c]; e = dict_offs[c + 1]; }
        }
        // byte offsets of the kept entries: inclusive scan of their lengths
        const int len = (int)(e - b);
        int incl = len;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
        unsigned char* dst = tx + tl + incl - len;
        for (int q = 0; q < len; ++q) dst[q] = dict_bytes[b + q];
        tl += __shfl_sync(0xffffffffu, incl, 31);
        cnt += __popc(m);
        for (unsigned r = m; r; r &= r - 1) acc = __fadd_rn(acc, __shfl_sync(0xffffffffu, p, __ffs(r) - 1));   // time order
    }
    for (int t = cnt + lane; t < max_t; t += 32) tk[t] = -1;
    if (lane == 0) {
        counts[line] = cnt;
        scores[line] = __fdiv_rn(acc, (float)cnt);  // 0/0 -> NaN like the reference (rec_processor.rs:94)
        text_len[line] = tl;
    }
}

// Packing of the decoded strings: the fixed-stride text buffer is mostly padding (stride = max_t * longest entry), so
// instead of copying it back whole, one block scans the line lengths into byte offsets and a warp per line writes
// its bytes at that offset — both straight into pinned, device-mapped host memory, so only the real text crosses PCIe
// and the host needs no further copy after its stream sync.
__global__ void __launch_bounds__(1024) ctc_text_scan_kernel(const int* __restrict__ text_len, int n_lines, unsigned* __restrict__ d_offs,
                                                             unsigned* __restrict__ h_offs) {
    __shared__ unsigned s_warp[32];
    __shared__ unsigned s_run;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (threadIdx.x == 0) { s_run = 0; d_offs[0] = 0; h_offs[0] = 0; }
    __syncthreads();
    for (int base = 0; base < n_lines; base += 1024) {
        const int i = base + threadIdx.x;
        const unsigned v = i < n_lines ? (unsigned)text_len[i] : 0u;
        unsigned x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const unsigned y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
        if (lane == 31) s_warp[w] = x;
        __syncthreads();
        if (w == 0) {
            unsigned t = s_warp[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const unsigned y = __shfl_up_sync(0xffffffffu, t, o); if (lane >= o) t += y; }
            s_warp[lane] = t;   // inclusive over warps
        }
        __syncthreads();
        const unsigned incl = s_run + (w ? s_warp[w - 1] : 0u) + x;
        if (i < n_lines) { d_offs[i + 1] = incl; h_offs[i + 1] = incl; }
        __syncthreads();
        if (threadIdx.x == 1023) s_run = incl;
        __syncthreads();
    }
}
__global__ void ctc_text_pack_kernel(const unsigned char* __restrict__ text, const unsigned* __restrict__ d_offs, int n_lines, int text_stride,
                                     unsigned char* __restrict__ h_text, unsigned capacity) {
    const int line = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (line >= n_lines) return;
    const unsigned b = d_offs[line], e = d_offs[line + 1];
    if (e > capacity) return;
    const unsigned char* src = text + (size_t)line * text_stride;
    for (unsigned k = lane; k < e - b; k += 32) h_text[b + k] = src[k];
}

// ---- host side ---------------------------------------------------------------------------------
static retto_b200_status ctc_prepare(retto_b200_ctx* ctx, const retto_b200_logits_desc* h_descs, int n_descs, int C,
                                     std::vector<CtcTensor>& tensors, std::vector<int>& prefix, std::vector<int>& line_t, int* max_t) {
    if (!h_descs || n_descs < 0 || C <= 0) { ctx->set_error("ctc: bad arguments"); return RETTO_B200_ERR_INVALID_ARG; }
    tensors.resize(n_descs);
    prefix.assign(n_descs + 1, 0);
    line_t.clear();
    int mt = 1;
    long long rows = 0;
    for (int k = 0; k < n_descs; ++k) {
        if (h_descs[k].n < 0 || h_descs[k].t <= 0) { ctx->set_error("ctc: bad tensor dims"); return RETTO_B200_ERR_INVALID_ARG; }
        tensors[k] = CtcTensor{h_descs[k].d_logits, h_descs[k].n, h_descs[k].t, (int)line_t.size(), 0};
        prefix[k] = (int)rows;
        rows += (long long)h_descs[k].n * h_descs[k].t;
        if (rows > 0x7fffffffLL) { ctx->set_error("ctc: too many rows"); return RETTO_B200_ERR_CAPACITY; }
        for (int i = 0; i < h_descs[k].n; ++i) line_t.push_back(h_descs[k].t);
        if (h_descs[k].n > 0) mt = std::max(mt, h_descs[k].t);
    }
    prefix[n_descs] = (int)rows;
    *max_t = mt;
    return RETTO_B200_OK;
}

static retto_b200_status ctc_run_argmax(retto_b200_ctx* ctx, const std::vector<CtcTensor>& tensors, const std::vector<int>& prefix, int C,
                                        int max_t, int n_lines, int* d_idx, float* d_prob, int* d_line_nan) {
    const int total_rows = prefix.back();
    if (total_rows == 0) return RETTO_B200_OK;
    std::vector<char> blob(tensors.size() * sizeof(CtcTensor) + prefix.size() * sizeof(int));
    memcpy(blob.data(), tensors.data(), tensors.size() * sizeof(CtcTensor));
    memcpy(blob.data() + tensors.size() * sizeof(CtcTensor), prefix.data(), prefix.size() * sizeof(int));
    RT_TRY(rt_upload(ctx, ctx->d_ctc_rows, blob.data(), blob.size()));
    const CtcTensor* d_t = ctx->d_ctc_rows.as<CtcTensor>();
    const int* d_pre = reinterpret_cast<const int*>(ctx->d_ctc_rows.as<char>() + tensors.size() * sizeof(CtcTensor));
    RT_CUDA_OK(ctx, cudaMemsetAsync(d_line_nan, 0, sizeof(int) * (size_t)n_lines, ctx->stream));
    constexpr int WARPS = 8;
    const int grid = (total_rows + WARPS - 1) / WARPS;
    RT_LAUNCH_BEGIN(ctx, "ctc_argmax_kernel<WARPS>");
    ctc_argmax_kernel<WARPS><<<grid, WARPS * 32, 0, ctx->stream>>>(d_t, d_pre, (int)tensors.size(), total_rows, C, max_t, d_idx, d_prob, d_line_nan);
    RT_LAUNCH_CHECK(ctx);
    return RETTO_B200_OK;
}

extern "C" retto_b200_status retto_b200_ctc_argmax(retto_b200_ctx* ctx, const retto_b200_logits_desc* h_descs, int32_t n_descs,
                                                   int32_t num_classes, int32_t* d_idx, float* d_prob) {
    RtDeviceGuard _dg(ctx);
    if (!ctx) return RETTO_B200_ERR_INVALID_ARG;
    std::vector<CtcTensor> tensors;
    std::vector<int> prefix, line_t;
    int max_t = 1;
    RT_TRY(ctc_prepare(ctx, h_descs, n_descs, num_classes, tensors, prefix, line_t, &max_t));
    // this tap writes [line][t] densely only when every tensor has the same t; otherwise stride = max_t
    const int n_lines = (int)line_t.size();
    RT_CUDA_OK(ctx, ctx->d_ctc_flag.ensure(sizeof(int) * (size_t)std::max(n_lines, 1), ctx->stream));
    return ctc_run_argmax(ctx, tensors, prefix, num_classes, max_t, n_lines, d_idx, d_prob, ctx->d_ctc_flag.as<int>());
}

// ctc_decode in two host steps (session.cu keeps several page batches in flight): begin enqueues argmax, collapse,
// text packing and the read-backs; end synchronises the stream and hands the results to the caller's arrays.
// want_tokens: also read the token ids back (max_t_out columns per line in rt_ctc_end).
retto_b200_status rt_ctc_begin(retto_b200_ctx* ctx, const retto_b200_logits_desc* h_descs, int32_t n_descs, int32_t num_classes,
                               bool want_tokens, int32_t max_t_out) {
    retto_b200_ctx::CtcRun& R = ctx->ctc;
    R = retto_b200_ctx::CtcRun{};
    int32_t* const h_tokens = want_tokens ? reinterpret_cast<int32_t*>(1) : nullptr;   // only tested for null below
    if (ctx->dict.empty()) { ctx->set_error("ctc_decode: no dictionary loaded"); return RETTO_B200_ERR_NO_DICT; }
    if ((int)ctx->dict.size() != num_classes) {
        ctx->set_error("ctc_decode: num_classes " + std::to_string(num_classes) + " != dictionary size " + std::to_string(ctx->dict.size()));
        return RETTO_B200_ERR_INVALID_ARG;
    }
    std::vector<CtcTensor> tensors;
    std::vector<int> prefix, line_t;
    int max_t = 1;
    RT_TRY(ctc_prepare(ctx, h_descs, n_descs, num_classes, tensors, prefix, line_t, &max_t));
    const int n_lines = (int)line_t.size();
    if (n_lines == 0) return RETTO_B200_OK;
    if (h_tokens && max_t_out < max_t) { ctx->set_error("ctc_decode: max_t too small for token output"); return RETTO_B200_ERR_INVALID_ARG; }
    const int text_stride = max_t * std::max(ctx->dict_max_len, 1);
    const size_t nl = (size_t)n_lines;
    RT_CUDA_OK(ctx, ctx->d_ctc_idx.ensure(sizeof(int) * nl * max_t, ctx->stream));
    RT_CUDA_OK(ctx, ctx->d_ctc_prob.ensure(sizeof(float) * nl * max_t, ctx->stream));
    RT_CUDA_OK(ctx, ctx->d_ctc_tok.ensure(sizeof(int) * nl * max_t, ctx->stream));
    RT_CUDA_OK(ctx, ctx->d_ctc_cnt.ensure(sizeof(int) * nl * 4, ctx->stream));  // counts | text_len | line_t | nan flags
    RT_CUDA_OK(ctx, ctx->d_ctc_score.ensure(sizeof(float) * nl, ctx->stream));
    RT_CUDA_OK(ctx, ctx->d_ctc_text.ensure(nl * text_stride, ctx->stream));
    int* d_cnt = ctx->d_ctc_cnt.as<int>();
    int* d_tlen = d_cnt + nl;
    int* d_linet = d_cnt + 2 * nl;
    int* d_nan = d_cnt + 3 * nl;
    RT_TRY(rt_upload_to(ctx, d_linet, line_t.data(), sizeof(int) * nl));
    RT_TRY(ctc_run_argmax(ctx, tensors, prefix, num_classes, max_t, n_lines, ctx->d_ctc_idx.as<int>(), ctx->d_ctc_prob.as<float>(), d_nan));
    RT_LAUNCH_BEGIN(ctx, "ctc_collapse_kernel");
    ctc_collapse_kernel<<<(n_lines + 7) / 8, 256, 0, ctx->stream>>>(
        ctx->d_ctc_idx.as<int>(), ctx->d_ctc_prob.as<float>(), d_linet, n_lines, max_t, ctx->d_dict_offs.as<unsigned>(),
        ctx->d_dict_bytes.as<unsigned char>(), (int)ctx->dict.size(), text_stride, ctx->d_ctc_tok.as<int>(), d_cnt,
        ctx->d_ctc_score.as<float>(), ctx->d_ctc_text.as<unsigned char>(), d_tlen);
    RT_LAUNCH_CHECK(ctx);
    // results -> host: counters/scores by copy, offsets + packed text written by the pack kernels into mapped pinned memory
    const size_t text_cap = nl * text_stride;
    const size_t o_score = sizeof(int) * nl * 4, o_offs = o_score + sizeof(float) * nl, o_text = (o_offs + sizeof(unsigned) * (nl + 1) + 15) & ~size_t(15);
    const size_t o_tok = (o_text + text_cap + 15) & ~size_t(15);
    const size_t hbytes = o_tok + (h_tokens ? sizeof(int) * nl * max_t : 0);
    RT_CUDA_OK(ctx, ctx->h_ctc.ensure(hbytes));
    RT_CUDA_OK(ctx, ctx->d_ctc_tlen.ensure(sizeof(unsigned) * (nl + 1), ctx->stream));
    char* hb = ctx->h_ctc.as<char>();
    char* hb_dev = nullptr;
    RT_CUDA_OK(ctx, cudaHostGetDevicePointer((void**)&hb_dev, hb, 0));
    int* hc = reinterpret_cast<int*>(hb);
    float* hs = reinterpret_cast<float*>(hb + o_score);
    const unsigned* hoff = reinterpret_cast<const unsigned*>(hb + o_offs);
    const unsigned char* ht = reinterpret_cast<const unsigned char*>(hb + o_text);
    int* htok = reinterpret_cast<int*>(hb + o_tok);
    RT_LAUNCH_BEGIN(ctx, "ctc_text_scan_kernel");
    ctc_text_scan_kernel<<<1, 1024, 0, ctx->stream>>>(d_tlen, n_lines, ctx->d_ctc_tlen.as<unsigned>(), reinterpret_cast<unsigned*>(hb_dev + o_offs));
    RT_LAUNCH_CHECK(ctx);
    RT_LAUNCH_BEGIN(ctx, "ctc_text_pack_kernel");
    ctc_text_pack_kernel<<<(n_lines + 7) / 8, 256, 0, ctx->stream>>>(ctx->d_ctc_text.as<unsigned char>(), ctx->d_ctc_tlen.as<unsigned>(), n_lines, text_stride,
                                                                     reinterpret_cast<unsigned char*>(hb_dev + o_text), (unsigned)std::min<size_t>(text_cap, 0xffffffffu));
    RT_LAUNCH_CHECK(ctx);
    RT_CUDA_OK(ctx, cudaMemcpyAsync(hc, d_cnt, sizeof(int) * nl * 4, cudaMemcpyDeviceToHost, ctx->stream));
    RT_CUDA_OK(ctx, cudaMemcpyAsync(hs, ctx->d_ctc_score.p, sizeof(float) * nl, cudaMemcpyDeviceToHost, ctx->stream));
    if (h_tokens) RT_CUDA_OK(ctx, cudaMemcpyAsync(htok, ctx->d_ctc_tok.p, sizeof(int) * nl * max_t, cudaMemcpyDeviceToHost, ctx->stream));
    R.n_lines = n_lines; R.max_t = max_t; R.o_score = o_score; R.o_offs = o_offs; R.o_text = o_text; R.o_tok = o_tok; R.want_tokens = want_tokens;
    return RETTO_B200_OK;
}

retto_b200_status rt_ctc_end(retto_b200_ctx* ctx, uint32_t* h_text_offsets, char* h_text, size_t text_capacity, float* h_scores,
                             int32_t* h_tokens, int32_t* h_token_counts, int32_t max_t_out) {
    const retto_b200_ctx::CtcRun& R = ctx->ctc;
    h_text_offsets[0] = 0;
    const int n_lines = R.n_lines, max_t = R.max_t;
    if (n_lines == 0) return RETTO_B200_OK;
    const size_t nl = (size_t)n_lines;
    char* hb = ctx->h_ctc.as<char>();
    int* hc = reinterpret_cast<int*>(hb);
    float* hs = reinterpret_cast<float*>(hb + R.o_score);
    const unsigned* hoff = reinterpret_cast<const unsigned*>(hb + R.o_offs);
    const unsigned char* ht = reinterpret_cast<const unsigned char*>(hb + R.o_text);
    int* htok = reinterpret_cast<int*>(hb + R.o_tok);
    if (!R.want_tokens) h_tokens = nullptr;
    RT_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
    const int* h_cnt = hc;
    const int* h_nan = hc + 3 * nl;
    retto_b200_status st = RETTO_B200_OK;
    const size_t total_text = hoff[n_lines];
    if (total_text > text_capacity) { ctx->set_error("ctc_decode: text buffer too small"); return RETTO_B200_ERR_CAPACITY; }
    memcpy(h_text, ht, total_text);
    for (int i = 0; i < n_lines; ++i) {
        if (h_nan[i]) { ctx->set_error("ctc_decode: NaN logits in line " + std::to_string(i)); st = RETTO_B200_ERR_NAN_LOGITS; }
        h_text_offsets[i + 1] = hoff[i + 1];
        h_scores[i] = hs[i];
        if (h_token_counts) h_token_counts[i] = h_cnt[i];
        if (h_tokens) {
            for (int t = 0; t < max_t_out; ++t) h_tokens[(size_t)i * max_t_out + t] = t < max_t ? htok[(size_t)i * max_t + t] : -1;
        }
    }
    return st;
}

extern "C" retto_b200_status retto_b200_ctc_decode(retto_b200_ctx* ctx, const retto_b200_logits_desc* h_descs, int32_t n_descs,
                                                   int32_t num_classes, uint32_t* h_text_offsets, char* h_text, size_t text_capacity,
                                                   float* h_scores, int32_t* h_tokens, int32_t* h_token_counts, int32_t max_t_out) {
    RtDeviceGuard _dg(ctx);
    if (!ctx || !h_text_offsets) return RETTO_B200_ERR_INVALID_ARG;
    RT_TRY(rt_ctc_begin(ctx, h_descs, n_descs, num_classes, h_tokens != nullptr, max_t_out));
    return rt_ctc_end(ctx, h_text_offsets, h_text, text_capacity, h_scores, h_tokens, h_token_counts, max_t_out);
}

extern "C" retto_b200_status retto_b200_dict_load(retto_b200_ctx* ctx, const char* utf8, size_t len) {
    RtDeviceGuard _dg(ctx);
    if (!ctx || (!utf8 && len)) return RETTO_B200_ERR_INVALID_ARG;
    // RecCharacter::new (rec_processor.rs:29-46): content.lines().map(str::trim) ; insert(0,"blank") ; push(" ")
    std::vector<std::string> d;
    d.push_back("blank");
    size_t i = 0;
    auto is_ws = [](const std::string& s, size_t pos, size_t* adv) -> bool {
        // Rust char::is_whitespace (White_Space property) for the UTF-8 sequence at pos
        const unsigned char c = (unsigned char)s[pos];
        if (c == ' ' || (c >= 0x09 && c <= 0x0d)) { *adv = 1; return true; }
        if (c == 0xC2 && pos + 1 < s.size()) {
            const unsigned char c1 = (unsigned char)s[pos + 1];
            if (c1 == 0x85 || c1 == 0xA0) { *adv = 2; return true; }
        }
        if (c == 0xE1 && pos + 2 < s.size() && (unsigned char)s[pos + 1] == 0x9A && (unsigned char)s[pos + 2] == 0x80) { *adv = 3; return true; }
        if (c == 0xE2 && pos + 2 < s.size()) {
            const unsigned char c1 = (unsigned char)s[pos + 1], c2 = (unsigned char)s[pos + 2];
            if (c1 == 0x80 && ((c2 >= 0x80 && c2 <= 0x8A) || c2 == 0xA8 || c2 == 0xA9 || c2 == 0xAF)) { *adv = 3; return true; }
            if (c1 == 0x81 && c2 == 0x9F) { *adv = 3; return true; }
        }
        if (c == 0xE3 && pos + 2 < s.size() && (unsigned char)s[pos + 1] == 0x80 && (unsigned char)s[pos + 2] == 0x80) { *adv = 3; return true; }
        return false;
    };
    while (i < len) {
        size_t j = i;
        while (j < len && utf8[j] != '\n') ++j;
        size_t e = j;
        if (e > i && utf8[e - 1] == '\r') --e;  // str::lines strips "\r\n"
        std::string line(utf8 + i, e - i);
        // trim both ends
        size_t b = 0, adv = 0;
        while (b < line.size() && is_ws(line, b, &adv)) b += adv;
        size_t en = line.size();
        for (;;) {
            bool cut = false;
            for (size_t back = 1; back <= 3 && back <= en - b; ++back) {
                size_t a2 = 0;
                if (en - back >= b && is_ws(line, en - back, &a2) && a2 == back) { en -= back; cut = true; break; }
            }
            if (!cut) break;
        }
        d.push_back(line.substr(b, en - b));
        i = j + 1;
    }
    d.push_back(" ");
    std::vector<unsigned> offs(d.size() + 1, 0);
    std::string bytes;
    int mx = 1;
    for (size_t k = 0; k < d.size(); ++k) {
        offs[k] = (unsigned)bytes.size();
        bytes += d[k];
        if (k >= 1) mx = std::max(mx, (int)d[k].size());
    }
    offs[d.size()] = (unsigned)bytes.size();
    RT_TRY(rt_upload(ctx, ctx->d_dict_offs, offs.data(), offs.size() * sizeof(unsigned)));
    if (bytes.empty()) bytes.push_back('\0');
    RT_TRY(rt_upload(ctx, ctx->d_dict_bytes, bytes.data(), bytes.size()));
    ctx->dict = std::move(d);
    ctx->dict_max_len = mx;
    if (utf8 != ctx->dict_source.data()) ctx->dict_source.assign(utf8 ? utf8 : "", len);
    ctx->dict_version++;
    return RETTO_B200_OK;
}

extern "C" int32_t retto_b200_dict_size(const retto_b200_ctx* ctx) { return ctx ? (int32_t)ctx->dict.size() : 0; }
