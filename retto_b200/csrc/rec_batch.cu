// rec_batch.cu — K8 / K9: cls + rec batch building and the cls postprocess / 180-degree flip.
//   plan_batches  (host) : ordering + batching rules of ClsProcessor::process / RecProcessor::process
//                          (cls_processor.rs:135-140, rec_processor.rs:223-238) — sizes only, no pixels
//   build_batches (K8)   : ImageHelper::resize_norm_image + concatenate (image_helper.rs:176-209):
//                          thumbnail to 48 x resized_w, px/255, (v-.5)/.5, RGB planes, zero pad to img_w,
//                          written straight into the [n,3,48,img_w] batch tensor; rec reads the crop through
//                          its flip flag (rotate_180_in_place as an index transform: zero bytes moved)
//   cls_postprocess (K9) : first-max argmax over [n,2], label/score, flip flag when label==180 && score>=thresh
// Roofline: HBM; algorithmic bytes per line = 3*w*h (crop read) + 12*48*img_w (tensor write).
#include "common.cuh"
#include "thumbnail.cuh"

struct FlipReader {   // crops are stored RGBX, one aligned word per pixel
    const uchar4* __restrict__ src;
    unsigned w, h;
    int flip;
    __device__ __forceinline__ uchar3 operator()(unsigned x, unsigned y) const {
        if (flip) { x = w - 1 - x; y = h - 1 - y; }
        const uchar4 p = __ldg(src + (size_t)y * w + x);
        return make_uchar3(p.x, p.y, p.z);
    }
};

// One block = one (line, 128-column chunk); one thread = one output column over all img_h rows.
// Everything that depends only on x is computed once per thread, everything that depends only on y once per block
// (shared memory); the normalisation `(px as f32 / 255 - .5) / .5` (image_helper.rs:200-203) comes from a 256-entry
// table built with the same three correctly-rounded operations.  Text crops are 20-60 px tall, so both thumbnail
// windows are at most 2 px wide almost always: that case runs a branch-light path on the 2x2 source pixels
// (i0|i1) x (j0|j1) that evaluates exactly the expression of the matching imageops::thumbnail branch; larger
// windows fall back to the generic thumbnail_pixel.
#define BB_COLS 128
#define BB_MAX_H 64
struct ChunkDev { int line, x0; };
struct AxisS {
    int i0, i1;      // the two source indices the window touches (i1 == i0 for a 1-px block)
    int n;           // block length (0 => fractional case between i0 and i1)
    float fract;
};
__device__ __forceinline__ AxisS axis_small(const ThumbAxis a, unsigned size) {
    AxisS r;
    if (a.lo != a.hi) { r.n = (int)(a.hi - a.lo); r.i0 = (int)a.lo; r.i1 = (int)a.hi - 1; r.fract = 0.0f; }
    else { r.n = 0; r.i0 = (int)a.hi - 1; r.i1 = (int)((a.hi > size - 1) ? size - 1 : a.hi); r.fract = a.fract; }
    return r;
}
struct RowS { int o0, o1; int n; float fract, omf, ft1, fb1, ft2, fb2; };   // per output row (pixel offsets of rows j0/j1)

__global__ void __launch_bounds__(BB_COLS, 6) build_batches_kernel(const LineDev* __restrict__ lines, const ChunkDev* __restrict__ chunks,
                                                                 const CropDev* __restrict__ crops, const unsigned char* __restrict__ crop_pix,
                                                                 const int* __restrict__ flip_flags, int use_flip, int img_h,
                                                                 float* __restrict__ out) {
    __shared__ float s_lut[256];
    __shared__ ThumbAxis s_ay[BB_MAX_H];
    __shared__ RowS s_row[BB_MAX_H];
    const ChunkDev ck = chunks[blockIdx.x];
    const LineDev ln = lines[ck.line];
    const CropDev& c = crops[ln.crop];
    const unsigned cw = (unsigned)c.w, chh = (unsigned)c.h;
    const int flip = use_flip ? flip_flags[ln.crop] : 0;
    for (int v = threadIdx.x; v < 256; v += BB_COLS) s_lut[v] = __fdiv_rn(__fsub_rn(__fdiv_rn((float)v, 255.0f), 0.5f), 0.5f);
    const float yr = __fdiv_rn((float)chh, (float)img_h);
    for (int y = threadIdx.x; y < img_h; y += BB_COLS) {
        const ThumbAxis ay = thumb_axis(y, yr, chh);
        s_ay[y] = ay;
        const AxisS a = axis_small(ay, chh);
        RowS r;
        const int j0 = flip ? (int)chh - 1 - a.i0 : a.i0, j1 = flip ? (int)chh - 1 - a.i1 : a.i1;
        r.o0 = j0 * (int)cw; r.o1 = j1 * (int)cw; r.n = a.n; r.fract = a.fract; r.omf = __fsub_rn(1.0f, a.fract);
        r.ft1 = a.fract; r.fb1 = r.omf;                                   // fract / 1, (1 - fract) / 1
        r.ft2 = __fdiv_rn(a.fract, 2.0f); r.fb2 = __fdiv_rn(r.omf, 2.0f);
        s_row[y] = r;
    }
    __syncthreads();
    const int x = ck.x0 + threadIdx.x;
    if (x >= ln.img_w) return;
    const size_t plane = (size_t)img_h * ln.img_w;
    float* dst = out + ln.dst_offset + x;
    if (x >= ln.resized_w) {
        for (int y = 0; y < img_h; ++y) {
            float* d = dst + (size_t)y * ln.img_w;
            d[0] = 0.0f; d[plane] = 0.0f; d[2 * plane] = 0.0f;
        }
        return;
    }
    const uchar4* __restrict__ src = reinterpret_cast<const uchar4*>(crop_pix + c.offset);
    const FlipReader rd{src, cw, chh, flip};
    const float xr = __fdiv_rn((float)cw, (float)ln.resized_w);
    const ThumbAxis ax = thumb_axis(x, xr, cw);
    const AxisS xs = axis_small(ax, cw);
    const int c0 = flip ? (int)cw - 1 - xs.i0 : xs.i0, c1 = flip ? (int)cw - 1 - xs.i1 : xs.i1;
    const float fh = xs.fract, omfh = __fsub_rn(1.0f, fh);
    const float fr1 = fh, fl1 = omfh, fr2 = __fdiv_rn(fh, 2.0f), fl2 = __fdiv_rn(omfh, 2.0f);
    const float fhu = xs.n ? 0.0f : fh, omfhu = xs.n ? 1.0f : omfh;   // (1, 0) weights for a 1-px block column
    // software pipeline: the four source pixels of row y+1 are in flight while row y is mixed and stored
    float* d = dst;
    const int stride = ln.img_w;
    RowS r = s_row[0];
    uchar4 p00 = __ldg(src + r.o0 + c0), p10 = __ldg(src + r.o0 + c1), p01 = __ldg(src + r.o1 + c0), p11 = __ldg(src + r.o1 + c1);
#pragma unroll 2
    for (int y = 0; y < img_h; ++y, d += stride) {
        const RowS rn = s_row[(y + 1 < img_h) ? y + 1 : y];
        const uchar4 n00 = __ldg(src + rn.o0 + c0), n10 = __ldg(src + rn.o0 + c1), n01 = __ldg(src + rn.o1 + c0), n11 = __ldg(src + rn.o1 + c1);
        unsigned char px[3];
        if (xs.n <= 1 && r.n <= 1) {
            // Up-scaling regime (the common one: crops are shorter than 48 px): every window is a single pixel or a
            // fractional pair, and all four imageops::thumbnail branches collapse into the bilinear expression when a
            // 1-px block axis is given the weights (1, 0):  x*1 = x, x*0 = 0, 0 + a = a are exact, so
            //   block/block -> p00;  h-fraction -> fl*p00 + fr*p10;  v-fraction -> fb*p00 + ft*p01   bit for bit.
            // One divergence-free path for the whole warp.
            const float fv = r.n ? 0.0f : r.fract, omv = r.n ? 1.0f : r.omf;
            const float f_tr = __fmul_rn(fv, fhu), f_tl = __fmul_rn(fv, omfhu), f_br = __fmul_rn(omv, fhu), f_bl = __fmul_rn(omv, omfhu);
#define RT_BL(ch) (unsigned char)__float2uint_rz(__fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(f_br, (float)p10.ch), __fmul_rn(f_tr, (float)p11.ch)), __fmul_rn(f_bl, (float)p00.ch)), __fmul_rn(f_tl, (float)p01.ch)))
            px[0] = RT_BL(x); px[1] = RT_BL(y); px[2] = RT_BL(z);
#undef RT_BL
        } else if (xs.n > 2 || r.n > 2) {
            thumbnail_pixel(rd, cw, chh, ax, s_ay[y], px);
        } else if (xs.n > 0 && r.n > 0) {          // block mean over (1|2) x (1|2) pixels: (sum + n/2) / n
            const unsigned n = (unsigned)(xs.n * r.n), h2 = n >> 1;
            const unsigned w10 = xs.n > 1, w01 = r.n > 1, w11 = w10 & w01;
            const unsigned sh = (n == 4) ? 2 : (n == 2) ? 1 : 0;   // n is 1, 2 or 4 here: the division is a shift
            px[0] = (unsigned char)((p00.x + w10 * p10.x + w01 * p01.x + w11 * p11.x + h2) >> sh);
            px[1] = (unsigned char)((p00.y + w10 * p10.y + w01 * p01.y + w11 * p11.y + h2) >> sh);
            px[2] = (unsigned char)((p00.z + w10 * p10.z + w01 * p01.z + w11 * p11.z + h2) >> sh);
        } else if (xs.n == 0 && r.n > 0) {  // horizontal fraction between columns i0,i1 summed over r.n rows
            const float fl = r.n > 1 ? fl2 : fl1, fr = r.n > 1 ? fr2 : fr1;
            const unsigned w01 = r.n > 1;
#define RT_HF(ch) f32_to_u8_numcast(__fadd_rn(__fmul_rn(fl, (float)(p00.ch + w01 * p01.ch)), __fmul_rn(fr, (float)(p10.ch + w01 * p11.ch))))
            px[0] = RT_HF(x); px[1] = RT_HF(y); px[2] = RT_HF(z);
#undef RT_HF
        } else if (xs.n > 0 && r.n == 0) {  // vertical fraction between rows j0,j1 summed over xs.n columns
            const float fb = xs.n > 1 ? r.fb2 : r.fb1, ft = xs.n > 1 ? r.ft2 : r.ft1;
            const unsigned w10 = xs.n > 1;
#define RT_VF(ch) f32_to_u8_numcast(__fadd_rn(__fmul_rn(fb, (float)(p00.ch + w10 * p10.ch)), __fmul_rn(ft, (float)(p01.ch + w10 * p11.ch))))
            px[0] = RT_VF(x); px[1] = RT_VF(y); px[2] = RT_VF(z);
#undef RT_VF
        } else {                             // both fractional: bilinear on the 2x2
            const float fv = r.fract;
            const float f_tr = __fmul_rn(fv, fh), f_tl = __fmul_rn(fv, omfh), f_br = __fmul_rn(r.omf, fh), f_bl = __fmul_rn(r.omf, omfh);
#define RT_BL(ch) f32_to_u8_numcast(__fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(f_br, (float)p10.ch), __fmul_rn(f_tr, (float)p11.ch)), __fmul_rn(f_bl, (float)p00.ch)), __fmul_rn(f_tl, (float)p01.ch)))
            px[0] = RT_BL(x); px[1] = RT_BL(y); px[2] = RT_BL(z);
#undef RT_BL
        }
        d[0] = s_lut[px[0]];
        d[plane] = s_lut[px[1]];
        d[2 * plane] = s_lut[px[2]];
        r = rn; p00 = n00; p10 = n10; p01 = n01; p11 = n11;
    }
}

// ---- K9 ------------------------------------------------------------------------------------------------
struct ClsLine { const float* logits; int crop; int pad; };
__global__ void cls_post_kernel(const ClsLine* __restrict__ lines, int n, int ncls, int label0, int label1, float thresh,
                                int* __restrict__ flip_flags, int* __restrict__ out_label, float* __restrict__ out_score,
                                int* __restrict__ nan_flag) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float* r = lines[i].logits;
    int best = 0;
    float bv = r[0];
    bool nan = bv != bv;
    for (int k = 1; k < ncls; ++k) {
        const float v = r[k];
        nan |= (v != v);
        if (v > bv) { bv = v; best = k; }
    }
    if (nan) { *nan_flag = 1; }
    const int label = best == 0 ? label0 : label1;
    out_label[i] = label;
    out_score[i] = bv;
    // cls_processor.rs:164-166
    if (label == 180 && bv >= thresh) flip_flags[lines[i].crop] ^= 1;
}

retto_b200_status rt_cls_collect(retto_b200_ctx* ctx, int n, retto_b200_cls_result* h_results);

// ---- host ------------------------------------------------------------------------------------------------
extern "C" retto_b200_status retto_b200_plan_batches(const retto_b200_config* cfg, int32_t kind, const retto_b200_crop_info* crops,
                                                     int32_t n, retto_b200_line_job* h_lines, retto_b200_batch* h_batches,
                                                     int32_t* n_batches, uint64_t* total_floats) {
    if (!cfg || n < 0 || (n > 0 && (!crops || !h_lines || !h_batches)) || !n_batches || !total_floats) return RETTO_B200_ERR_INVALID_ARG;
    const int* shape = kind == 0 ? cfg->cls_image_shape : cfg->rec_image_shape;
    const int img_h = shape[1], img_w_cfg = shape[2];
    const int batch_num = std::max(1, kind == 0 ? cfg->cls_batch_num : cfg->rec_batch_num);
    // stable sort by Reverse(OrderedFloat(ori_ratio)), ori_ratio = h as f64 / w as f64 (image_helper.rs:79-82)
    std::vector<int> idx(n);
    for (int i = 0; i < n; ++i) idx[i] = i;
    std::stable_sort(idx.begin(), idx.end(), [&](int a, int b) {
        const double ra = (double)crops[a].h / (double)crops[a].w, rb = (double)crops[b].h / (double)crops[b].w;
        return ra > rb;  // descending; OrderedFloat total order (NaN greatest) is unreachable for w,h > 0
    });
    float max_wh_ratio = (float)img_w_cfg / (float)img_h;  // rec_processor.rs:226-227, carried across batches
    uint64_t off = 0;
    int nb = 0;
    for (int b0 = 0; b0 < n; b0 += batch_num) {
        const int nl = std::min(batch_num, n - b0);
        int img_w = img_w_cfg;
        if (kind == 1) {
            for (int k = 0; k < nl; ++k) {
                const retto_b200_crop_info& c = crops[idx[b0 + k]];
                const float wh = (float)c.w / (float)c.h;  // rec_processor.rs:234-236 (current dims == crop dims)
                if (wh > max_wh_ratio) max_wh_ratio = wh;
            }
            img_w = (int)(size_t)((float)img_h * max_wh_ratio);  // image_helper.rs:179
        }
        retto_b200_batch& bt = h_batches[nb++];
        bt.first_line = b0; bt.n = nl; bt.img_w = img_w; bt.max_wh_ratio = max_wh_ratio; bt.offset = off;
        for (int k = 0; k < nl; ++k) {
            const retto_b200_crop_info& c = crops[idx[b0 + k]];
            const double rw = std::ceil((double)img_h * (double)c.w / (double)c.h);  // image_helper.rs:183
            retto_b200_line_job& l = h_lines[b0 + k];
            l.crop = idx[b0 + k];
            l.img_w = img_w;
            l.resized_w = (int)std::min<size_t>((size_t)img_w, (size_t)rw);
            l.dst_offset = off + (uint64_t)k * 3 * img_h * img_w;
        }
        off += (uint64_t)nl * 3 * img_h * img_w;
    }
    *n_batches = nb;
    *total_floats = off;
    return RETTO_B200_OK;
}

extern "C" retto_b200_status retto_b200_build_batches(retto_b200_ctx* ctx, int32_t kind, const retto_b200_line_job* h_lines, int32_t n_lines,
                                                      uint64_t total_floats, float** d_base) {
    if (!ctx || n_lines < 0 || (n_lines > 0 && !h_lines) || !d_base) return RETTO_B200_ERR_INVALID_ARG;
    DevBuf& buf = kind == 0 ? ctx->d_batch_cls : ctx->d_batch_rec;
    RT_CUDA_OK(ctx, buf.ensure(std::max<size_t>((size_t)total_floats * 4, 16), ctx->stream));
    *d_base = buf.as<float>();
    if (n_lines == 0) return RETTO_B200_OK;
    const int img_h = kind == 0 ? ctx->cfg.cls_image_shape[1] : ctx->cfg.rec_image_shape[1];
    if (img_h > BB_MAX_H) { ctx->set_error("build_batches: image_shape height > 64 is not supported"); return RETTO_B200_ERR_UNSUPPORTED; }
    std::vector<LineDev> lines(n_lines);
    std::vector<ChunkDev> chunks;
    chunks.reserve((size_t)n_lines * 6);
    for (int i = 0; i < n_lines; ++i) {
        const retto_b200_line_job& l = h_lines[i];
        if (l.crop < 0 || l.crop >= (int)ctx->crops.size() || l.img_w <= 0 || l.resized_w < 0 || l.resized_w > l.img_w ||
            l.dst_offset + (uint64_t)3 * img_h * l.img_w > total_floats || ctx->crops[l.crop].status != RETTO_B200_OK) {
            ctx->set_error("build_batches: bad line " + std::to_string(i));
            return RETTO_B200_ERR_INVALID_ARG;
        }
        lines[i] = LineDev{l.crop, l.img_w, l.resized_w, 0, l.dst_offset};
        for (int x0 = 0; x0 < l.img_w; x0 += BB_COLS) chunks.push_back(ChunkDev{i, x0});
    }
    const size_t lb = (sizeof(LineDev) * n_lines + 15) & ~size_t(15);
    std::vector<char> blob(lb + sizeof(ChunkDev) * chunks.size());
    memcpy(blob.data(), lines.data(), sizeof(LineDev) * n_lines);
    memcpy(blob.data() + lb, chunks.data(), sizeof(ChunkDev) * chunks.size());
    RT_TRY(rt_upload(ctx, ctx->d_lines, blob.data(), blob.size()));
    const LineDev* d_lines = ctx->d_lines.as<LineDev>();
    const ChunkDev* d_chunks = reinterpret_cast<const ChunkDev*>(ctx->d_lines.as<char>() + lb);
    RT_LAUNCH_BEGIN(ctx, "build_batches_kernel");
    build_batches_kernel<<<(unsigned)chunks.size(), BB_COLS, 0, ctx->stream>>>(d_lines, d_chunks, ctx->d_crop_descs.as<CropDev>(),
                                                                             ctx->d_crop_pix.as<unsigned char>(), ctx->d_crop_flip.as<int>(),
                                                                             kind == 1 ? 1 : 0, img_h, buf.as<float>());
    RT_LAUNCH_CHECK(ctx);
    return RETTO_B200_OK;
}

// enqueue K9 + the result read-back; with defer == true the caller collects after its next stream sync
retto_b200_status rt_cls_postprocess_ptrs(retto_b200_ctx* ctx, const std::vector<const float*>& logits, const int32_t* crop_index, int n,
                                          retto_b200_cls_result* h_results, bool defer) {
    if (n == 0) return RETTO_B200_OK;
    std::vector<ClsLine> lines(n);
    for (int i = 0; i < n; ++i) {
        if (crop_index[i] < 0 || crop_index[i] >= (int)ctx->crops.size()) { ctx->set_error("cls_postprocess: bad crop index"); return RETTO_B200_ERR_INVALID_ARG; }
        lines[i] = ClsLine{logits[i], crop_index[i], 0};
    }
    RT_TRY(rt_upload(ctx, ctx->d_cls_idx, lines.data(), sizeof(ClsLine) * n));
    RT_CUDA_OK(ctx, ctx->d_cls_out.ensure(sizeof(int) * (2 * (size_t)n + 1), ctx->stream));
    int* d_label = ctx->d_cls_out.as<int>();
    float* d_score = reinterpret_cast<float*>(d_label + n);
    int* d_nan = d_label + 2 * n;
    RT_CUDA_OK(ctx, cudaMemsetAsync(d_nan, 0, sizeof(int), ctx->stream));
    RT_LAUNCH_BEGIN(ctx, "cls_post_kernel");
    cls_post_kernel<<<(n + 127) / 128, 128, 0, ctx->stream>>>(ctx->d_cls_idx.as<ClsLine>(), n, 2, ctx->cfg.cls_label[0], ctx->cfg.cls_label[1],
                                                              ctx->cfg.cls_thresh, ctx->d_crop_flip.as<int>(), d_label, d_score, d_nan);
    RT_LAUNCH_CHECK(ctx);
    RT_CUDA_OK(ctx, ctx->h_cls.ensure(sizeof(int) * (2 * (size_t)n + 1)));
    RT_CUDA_OK(ctx, cudaMemcpyAsync(ctx->h_cls.p, d_label, sizeof(int) * (2 * (size_t)n + 1), cudaMemcpyDeviceToHost, ctx->stream));
    if (defer) return RETTO_B200_OK;
    RT_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
    return rt_cls_collect(ctx, n, h_results);
}
// after a stream sync: results of the last rt_cls_postprocess_ptrs
retto_b200_status rt_cls_collect(retto_b200_ctx* ctx, int n, retto_b200_cls_result* h_results) {
    if (n == 0) return RETTO_B200_OK;
    const int* hl = ctx->h_cls.as<int>();
    const float* hs = reinterpret_cast<const float*>(hl + n);
    for (int i = 0; i < n; ++i) { h_results[i].label = hl[i]; h_results[i].score = hs[i]; }
    if (hl[2 * n]) { ctx->set_error("cls_postprocess: NaN logits (reference: argmax().unwrap() panics, cls_processor.rs:113)"); return RETTO_B200_ERR_NAN_LOGITS; }
    return RETTO_B200_OK;
}

extern "C" retto_b200_status retto_b200_cls_postprocess(retto_b200_ctx* ctx, const float* d_logits, int32_t n, const int32_t* h_crop_index,
                                                        retto_b200_cls_result* h_results) {
    if (!ctx || n < 0 || (n > 0 && (!d_logits || !h_crop_index || !h_results))) return RETTO_B200_ERR_INVALID_ARG;
    std::vector<const float*> ptrs(n);
    for (int i = 0; i < n; ++i) ptrs[i] = d_logits + 2 * (size_t)i;
    return rt_cls_postprocess_ptrs(ctx, ptrs, h_crop_index, n, h_results, false);
}
