/* retto_b200.h — C ABI of libretto_b200.so: the B200-native (sm_100a) replacement for the CPU image
 * code of retto-core's det / cls / rec processors.
 *
 * The reference (NekoImageLand/retto) has no FFI for this path: its seams are Rust traits
 * (retto-core/src/processor.rs:20-43 `Processor::{preprocess, postprocess, process}`,
 * retto-core/src/worker.rs:69-73 `RettoInnerWorker::{det, cls, rec}`) and `ImageHelper`
 * (retto-core/src/image_helper.rs).  Each entry point below names the reference function(s) it
 * replaces; INTEGRATION.md shows the `retto-b200-sys` extern block and the `backend-b200` call
 * sites a maintainer would add.
 *
 * Conventions
 *  - every function returns a retto_b200_status (0 = OK); retto_b200_last_error(ctx) has the text.
 *  - `d_` pointers are DEVICE pointers on the context's GPU; `h_` pointers are host memory.
 *  - all device work is enqueued on the context's stream; calls that fill host outputs synchronise
 *    that stream before returning, all others are asynchronous.
 *  - a context is NOT thread-safe (mirrors `RettoSession::run(&mut self)`, session.rs:108);
 *    multi-GPU = one context per GPU, one host thread / process each, no collective.
 *  - no CPU fallback exists: every pixel/logit operation below runs in a CUDA kernel.
 */
#ifndef RETTO_B200_H
#define RETTO_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RETTO_B200_ABI_VERSION 2

typedef enum retto_b200_status {
    RETTO_B200_OK = 0,
    RETTO_B200_ERR_INVALID_ARG = 1,
    RETTO_B200_ERR_CUDA = 2,
    RETTO_B200_ERR_OOM = 3,
    RETTO_B200_ERR_CAPACITY = 4,        /* a per-page cap (components, boxes, hull points) was exceeded   */
    RETTO_B200_ERR_NAN_LOGITS = 5,      /* reference: argmax().unwrap() panics (cls_processor.rs:113, rec_processor.rs:198) */
    RETTO_B200_ERR_DEGENERATE_QUAD = 6, /* reference: draw_polygon_mut / from_control_points().unwrap() panic (det_processor.rs:209, image_helper.rs:237) */
    RETTO_B200_ERR_NO_DICT = 7,
    RETTO_B200_ERR_WORKER = 8,          /* forward callback failed */
    RETTO_B200_ERR_UNSUPPORTED = 9,
    RETTO_B200_ERR_DECODE = 10          /* not an image / truncated or corrupt file (reference: ImageError from image::load_from_memory, error.rs:2-19) */
} retto_b200_status;

typedef struct retto_b200_ctx retto_b200_ctx;

/* Mirrors RettoSessionConfig (session.rs:17-40) + DetProcessorConfig (det_processor.rs:44-93) +
 * ClsProcessorConfig (cls_processor.rs:12-36) + RecProcessorConfig (rec_processor.rs:100-136).
 * Fields the reference never reads (max_candidates, score_mode, use_dilation) are kept for
 * struct-for-struct mapping and ignored here too. */
typedef struct retto_b200_config {
    int32_t max_side_len;          /* 2000 */
    int32_t min_side_len;          /* 30   */
    /* det */
    int32_t det_limit_side_len;    /* 736 */
    int32_t det_limit_type;        /* 0 = Min (default), 1 = Max */
    float   det_mean[3];           /* .5 .5 .5 */
    float   det_std[3];            /* .5 .5 .5 */
    float   det_scale;             /* 1f32/255f32 */
    float   det_thresh;            /* 0.3  ("threch") */
    float   det_box_thresh;        /* 0.5 */
    int32_t det_max_candidates;    /* 1000, unused by the reference */
    float   det_unclip_ratio;      /* 1.6 */
    int32_t det_use_dilation;      /* unused by the reference */
    int32_t det_score_mode;        /* unused by the reference */
    int32_t det_min_mini_box_size; /* 3 */
    int32_t det_dilation_2x2;      /* 1 = dilation_kernel Some(2x2 ones), 0 = None */
    /* cls */
    int32_t cls_image_shape[3];    /* 3,48,192 */
    int32_t cls_batch_num;         /* 6 */
    float   cls_thresh;            /* 0.9 */
    int32_t cls_label[2];          /* 0,180 */
    /* rec */
    int32_t rec_image_shape[3];    /* 3,48,320 */
    int32_t rec_batch_num;         /* 6 */
    /* capacity knobs of this implementation (0 = default) */
    int32_t max_components_per_page;  /* default 16384 */
    int32_t max_det_side;             /* default 4096: det tensor side cap */
} retto_b200_config;

void retto_b200_config_default(retto_b200_config* cfg);
int32_t retto_b200_abi_version(void);

/* ---- context ------------------------------------------------------------------------------- */
/* replaces: RettoSession::new's worker/device set-up (session.rs:62-73) for the image path */
retto_b200_status retto_b200_create(int32_t device_id, const retto_b200_config* cfg, retto_b200_ctx** out);
void retto_b200_destroy(retto_b200_ctx* ctx);
const char* retto_b200_last_error(const retto_b200_ctx* ctx);
/* the CUDA stream (cudaStream_t) all work of this context is enqueued on */
void* retto_b200_stream(retto_b200_ctx* ctx);
retto_b200_status retto_b200_sync(retto_b200_ctx* ctx);
/* number of kernels this context has launched so far (bench.py `gpu_launches`) */
uint64_t retto_b200_launch_count(const retto_b200_ctx* ctx);

/* measurement taps: when enabled every kernel launch is bracketed by CUDA events on the context's stream;
 * retto_b200_kernel_times returns "name\tcount\ttotal_ms\n" lines accumulated since the last reset */
retto_b200_status retto_b200_enable_kernel_timing(retto_b200_ctx* ctx, int32_t on);
retto_b200_status retto_b200_reset_kernel_times(retto_b200_ctx* ctx);
retto_b200_status retto_b200_kernel_times(retto_b200_ctx* ctx, char* buf, size_t cap);

/* memory helpers so a host-language binding needs no CUDA runtime of its own */
retto_b200_status retto_b200_dev_alloc(retto_b200_ctx* ctx, size_t bytes, void** d_out);
retto_b200_status retto_b200_dev_free(retto_b200_ctx* ctx, void* d_ptr);
retto_b200_status retto_b200_host_alloc(retto_b200_ctx* ctx, size_t bytes, void** h_out); /* pinned */
retto_b200_status retto_b200_host_free(retto_b200_ctx* ctx, void* h_ptr);
retto_b200_status retto_b200_h2d(retto_b200_ctx* ctx, void* d_dst, const void* h_src, size_t bytes); /* async */
retto_b200_status retto_b200_d2h(retto_b200_ctx* ctx, void* h_dst, const void* d_src, size_t bytes); /* async */

/* ---- image decode ------------------------------------------------------------------------------ */
/* ImageHelper::new_from_raw_img_flow (image_helper.rs:34-44: image::load_from_memory(bytes).to_rgb8()) on the device, for the file
 * kind that matters at scale: baseline JPEG (Huffman, 8 bit, YCbCr 4:4:4 / 4:2:2 / 4:2:0 / 4:4:0 or grayscale, interleaved scan).
 * Pixels are bit-exact with libjpeg-turbo's default decode (integer "islow" IDCT, fancy up-sampling — what Pillow / OpenCV give).
 * Restart intervals (DRI) are decoded in parallel, one GPU thread each; files with long intervals — above all files without DRI, which
 * are ONE interval — are cut into 1-KB sub-sequences that are parsed self-synchronisingly, so they decode in
 * parallel too.
 * Anything else (progressive / arithmetic / 12-bit JPEG, CMYK, PNG, ...) reports RETTO_B200_ERR_UNSUPPORTED — the caller decodes
 * those on the host and passes RGB; a damaged file reports RETTO_B200_ERR_DECODE.  There is no silent CPU fallback. */
typedef struct retto_b200_encoded {
    const uint8_t* bytes;      /* host memory: the whole file */
    uint64_t n_bytes;
} retto_b200_encoded;
typedef struct retto_b200_image_info_t {
    int32_t h, w;
    int32_t format;            /* 1 = baseline JPEG */
    int32_t components;        /* 1 or 3 */
    int32_t subsampling;       /* 0 = 4:4:4, 1 = 4:2:2, 2 = 4:2:0, 3 = grayscale, 4 = 4:4:0 */
    int32_t restart_interval;  /* MCUs per restart interval, 0 = none */
    int32_t status;            /* OK / ERR_UNSUPPORTED / ERR_DECODE */
} retto_b200_image_info_t;
/* header parse only (host, no context needed): the (h, w) a caller needs to size its buffers */
retto_b200_status retto_b200_image_info(const uint8_t* bytes, uint64_t n_bytes, retto_b200_image_info_t* out);
/* decode n files into caller-provided device buffers (d_out[i]: h*w*3 bytes HWC u8 RGB, sized from retto_b200_image_info).
 * h_status[i] is per file; the call returns the last non-OK status (the other files are still decoded).  Synchronises. */
retto_b200_status retto_b200_decode_images(retto_b200_ctx* ctx, const retto_b200_encoded* h_imgs, int32_t n, uint8_t* const* d_out,
                                           int32_t* h_status);

/* ---- resizes -------------------------------------------------------------------------------- */
/* ImageHelper::resize_both size math (image_helper.rs:106-148). dims = up to 2 (h,w) pairs applied
 * in order; returns the number of resizes (0, 1 or 2) in *n_steps. */
retto_b200_status retto_b200_resize_both_plan(int32_t h, int32_t w, int32_t max_side_len, int32_t min_side_len,
                                              int32_t dims[4], int32_t* n_steps);
/* ImageHelper::resize_either size math (image_helper.rs:150-174) */
retto_b200_status retto_b200_resize_either_plan(int32_t h, int32_t w, int32_t limit_type, int32_t limit_len,
                                                int32_t* out_h, int32_t* out_w);
/* image::imageops::thumbnail (image_helper.rs:124,139,168,184) on HWC u8 RGB, batched */
typedef struct retto_b200_resize_desc {
    const uint8_t* d_src; int32_t h, w;
    uint8_t* d_dst;       int32_t out_h, out_w;
} retto_b200_resize_desc;
retto_b200_status retto_b200_thumbnail(retto_b200_ctx* ctx, const retto_b200_resize_desc* h_descs, int32_t n);

/* ---- det preprocess --------------------------------------------------------------------------- */
/* DetProcessor::preprocess (det_processor.rs:256-274): resize_either -> rgb2bgr -> normalize ->
 * HWC->CHW.  One descriptor per page; d_out is [1,3,out_h,out_w] f32 (out dims from
 * retto_b200_resize_either_plan).  One launch covers all n pages. */
typedef struct retto_b200_det_pre_desc {
    const uint8_t* d_rgb; int32_t h, w;     /* page after resize_both, HWC u8 RGB */
    float* d_out;         int32_t out_h, out_w;
} retto_b200_det_pre_desc;
retto_b200_status retto_b200_det_preprocess(retto_b200_ctx* ctx, const retto_b200_det_pre_desc* h_descs, int32_t n);

/* ---- det postprocess -------------------------------------------------------------------------- */
/* DetProcessor::postprocess (det_processor.rs:279-335): threshold -> 2x2 dilate -> contours
 * (connected components + hole borders) -> min_area_rect -> box_score_fast -> unclip ->
 * min_area_rect -> scale_and_clip -> filters -> sorted_boxes.
 * One descriptor per page: d_prob is the [1,1,h,w] f32 probability map; (ori_h, ori_w) are the page
 * dims the boxes are scaled to (DetProcessor::new's ori_h/ori_w, session.rs:85). */
typedef struct retto_b200_det_post_desc {
    const float* d_prob; int32_t h, w;
    int32_t ori_h, ori_w;
} retto_b200_det_post_desc;
/* DetProcessorInnerResult (det_processor.rs:104-108): corners tl,tr,br,bl as (x,y) integer-valued f32 */
typedef struct retto_b200_box {
    float xy[8];
    float score;
} retto_b200_box;
/* Runs the whole batch; results stay on the device and are also copied to the host:
 *   h_page_status[n] : per page RETTO_B200_OK / ERR_DEGENERATE_QUAD / ERR_CAPACITY
 *   h_box_offsets[n+1], h_boxes[max_boxes_total] : boxes of page p are [offsets[p], offsets[p+1])
 * Returns ERR_CAPACITY if max_boxes_total is too small (h_box_offsets is still valid). */
retto_b200_status retto_b200_det_postprocess(retto_b200_ctx* ctx, const retto_b200_det_post_desc* h_descs, int32_t n,
                                             int32_t* h_page_status, int32_t* h_box_offsets,
                                             retto_b200_box* h_boxes, int32_t max_boxes_total);
/* debug/parity taps of the last det_postprocess call (page index p): bitmap u8 [h,w] (0/255) and
 * labels i32 [h,w] (-1 background, else min linear index of the 8-connected component) */
retto_b200_status retto_b200_det_post_fetch_bitmap(retto_b200_ctx* ctx, int32_t page, uint8_t* h_bitmap);
retto_b200_status retto_b200_det_post_fetch_labels(retto_b200_ctx* ctx, int32_t page, int32_t* h_labels);

/* parity tap: per connected component (in raster order of its first pixel) of page `page` of the last
 * det_postprocess call: discovery key, status (0 kept, 1 sside<min, 2 score<box_thresh, 3 sside2<min+2,
 * 4 size filter, 5 empty unclip, 6 never discovered by find_contours, -1 reference panic), the first
 * min_area_rect (8 ints), its sside and box score (NaN when not evaluated).  n_holes = number of hole
 * borders find_contours would additionally report (#components - Euler number). */
retto_b200_status retto_b200_det_post_enable_trace(retto_b200_ctx* ctx, int32_t on);
retto_b200_status retto_b200_det_post_fetch_trace(retto_b200_ctx* ctx, int32_t page, int32_t* n_components, int32_t* n_holes,
                                                  int32_t* h_key, int32_t* h_status, int32_t* h_rect1, float* h_sside1,
                                                  float* h_score, int32_t max_components);

/* PointBox::scale_and_clip (points.rs:179-194) on host box arrays — used for session.rs:94-97 */
retto_b200_status retto_b200_scale_and_clip(retto_b200_ctx* ctx, retto_b200_box* h_boxes, int32_t n,
                                            double bitmap_w, double bitmap_h, double ori_w, double ori_h);

/* ---- rotate-crop -------------------------------------------------------------------------------- */
/* ImageHelper::get_crop_img (image_helper.rs:223-249): perspective bicubic warp of each box to a
 * (W as u32) x (H as u32) crop, rotate270 when h/w >= 1.5.  Crops live in a context-owned arena;
 * a crop set handle stays valid until the next retto_b200_crop_boxes call or destroy. */
typedef struct retto_b200_crop_job {
    const uint8_t* d_page; int32_t page_h, page_w;  /* page after resize_both */
    retto_b200_box box;                              /* in that page's coordinates */
} retto_b200_crop_job;
typedef struct retto_b200_crop_info {
    int32_t w, h;          /* crop dims (after the optional rotate270) == ImageHelper::ori_w / ori_h */
    int32_t rotated270;
    int32_t status;        /* OK or ERR_DEGENERATE_QUAD */
    uint64_t offset;       /* byte offset of the HWC u8 crop inside the crop arena */
} retto_b200_crop_info;
retto_b200_status retto_b200_crop_boxes(retto_b200_ctx* ctx, const retto_b200_crop_job* h_jobs, int32_t n,
                                        retto_b200_crop_info* h_infos);
/* copy crop i to the host (HWC u8), applying the cls 180-degree flip if it has been flagged */
retto_b200_status retto_b200_crop_fetch(retto_b200_ctx* ctx, int32_t i, uint8_t* h_out);

/* ---- cls / rec batch build ---------------------------------------------------------------------- */
/* ImageHelper::resize_norm_image + concatenate (image_helper.rs:176-209, cls_processor.rs:142-151,
 * rec_processor.rs:239-251).  A batch plan lists, per line, the crop index, its slot in a batch
 * tensor, and that tensor's img_w; the host-side ordering rules (sort by Reverse(h/w), chunks of
 * batch_num, running max_wh_ratio: cls_processor.rs:137-140, rec_processor.rs:225-238) are
 * computed by retto_b200_plan_batches. */
typedef struct retto_b200_batch {
    int32_t first_line;    /* index into the plan's line list */
    int32_t n;             /* lines in this batch (<= batch_num) */
    int32_t img_w;         /* tensor is [n,3,img_h,img_w] */
    float   max_wh_ratio;  /* rec only: running maximum after this batch */
    uint64_t offset;       /* float offset of the tensor inside the batch arena */
} retto_b200_batch;
typedef struct retto_b200_line_job {
    int32_t crop;          /* index into the current crop set */
    int32_t img_w;         /* padded width of the destination tensor */
    int32_t resized_w;     /* min(img_w, ceil(img_h * w / h)) */
    uint64_t dst_offset;   /* float offset of this line's [3,img_h,img_w] block in the batch arena */
} retto_b200_line_job;
/* crops: per-crop (w,h) of ONE page in detection order.  kind 0 = cls, 1 = rec.
 * h_lines[n_crops] receives the planned lines in batch order (crop = index into `crops`, offsets
 * relative to this page's first tensor); h_batches[ceil(n/batch_num)]. */
retto_b200_status retto_b200_plan_batches(const retto_b200_config* cfg, int32_t kind, const retto_b200_crop_info* h_crops,
                                          int32_t n_crops, retto_b200_line_job* h_lines, retto_b200_batch* h_batches,
                                          int32_t* n_batches, uint64_t* total_floats);
/* kind 0 = cls (ignores flip flags), 1 = rec (reads crops through their flip flags).
 * The batch arena is context-owned; *d_base receives its device address. */
retto_b200_status retto_b200_build_batches(retto_b200_ctx* ctx, int32_t kind, const retto_b200_line_job* h_lines, int32_t n_lines,
                                           uint64_t total_floats, float** d_base);

/* ---- cls postprocess + flip ---------------------------------------------------------------------- */
/* ClsProcessor::postprocess + the rotate_180_in_place rule (cls_processor.rs:108-121,164-166).
 * d_logits: [n, 2] f32 rows for the lines h_crop_index[0..n).  Sets the flip flag of a crop when
 * label == 180 && score >= cls_thresh (consumed by build_batches(kind=1) and crop_fetch). */
typedef struct retto_b200_cls_result {
    int32_t label;  /* cls_label[argmax] */
    float score;
} retto_b200_cls_result;
retto_b200_status retto_b200_cls_postprocess(retto_b200_ctx* ctx, const float* d_logits, int32_t n, const int32_t* h_crop_index,
                                             retto_b200_cls_result* h_results);

/* ---- CTC greedy decode --------------------------------------------------------------------------- */
/* RecCharacter::new (rec_processor.rs:29-46): utf8 = the dictionary file; lines are trimmed, "blank"
 * is prepended and " " appended. */
retto_b200_status retto_b200_dict_load(retto_b200_ctx* ctx, const char* utf8, size_t len);
int32_t retto_b200_dict_size(const retto_b200_ctx* ctx);
/* RecProcessor::postprocess + RecCharacter::decode (rec_processor.rs:190-208, 48-97) for a list of
 * logits tensors: tensor k is [n_k, t_k, C] f32 (C must equal the dictionary size).
 * Outputs, line-major over all tensors in order (total = sum n_k):
 *   h_text_offsets[total+1] into h_text (UTF-8, not NUL terminated), h_scores[total]
 *   (NaN for an empty decode, rec_processor.rs:94).  Optional h_tokens / h_token_counts: kept class
 *   ids per line, max_t ids per line (pass NULL to skip). */
typedef struct retto_b200_logits_desc {
    const float* d_logits; int32_t n, t;
} retto_b200_logits_desc;
retto_b200_status retto_b200_ctc_decode(retto_b200_ctx* ctx, const retto_b200_logits_desc* h_descs, int32_t n_descs, int32_t num_classes,
                                        uint32_t* h_text_offsets, char* h_text, size_t text_capacity, float* h_scores,
                                        int32_t* h_tokens, int32_t* h_token_counts, int32_t max_t);
/* the argmax/max stage alone (rec_processor.rs:198-199), device-resident outputs [sum n_k * t_k] */
retto_b200_status retto_b200_ctc_argmax(retto_b200_ctx* ctx, const retto_b200_logits_desc* h_descs, int32_t n_descs, int32_t num_classes,
                                        int32_t* d_idx, float* d_prob);

/* ---- session: RettoSession::run over a batch of pages ---------------------------------------------- */
/* Forward seam == RettoInnerWorker::{det,cls,rec} (worker.rs:69-73) with device-resident tensors:
 * the callback receives n input tensors (device pointers, NCHW f32) and must fill n output tensors
 * (device pointers owned by the worker, valid until its next call for the same stage), all ordered
 * on `stream`.  stage: 0 det ([1,3,H,W] -> [1,1,H,W]), 1 cls ([n,3,48,192] -> [n,2]),
 * 2 rec ([n,3,48,W] -> [n,T,C]).  Return 0 on success. */
typedef struct retto_b200_tensor {
    float* d_data;
    int64_t shape[4];
    int32_t ndim;
} retto_b200_tensor;
typedef int32_t (*retto_b200_forward_fn)(void* user, int32_t stage, int32_t n, const retto_b200_tensor* inputs,
                                         retto_b200_tensor* outputs, void* stream);

#define RETTO_B200_PAGE_HOST_RGB 0      /* rgb = HWC u8 RGB in host memory (copied inside the call) */
#define RETTO_B200_PAGE_DEVICE_RGB 1    /* rgb = HWC u8 RGB device pointer */
#define RETTO_B200_PAGE_HOST_ENCODED 2  /* rgb = the image FILE bytes in host memory (n_bytes of them): RettoSession::run's own input
                                           (session.rs:75-79, main.rs:83-84); decoded on the device (see "image decode" above);
                                           h, w are ignored on input */
typedef struct retto_b200_page {
    const uint8_t* rgb;
    int32_t h, w;
    int32_t on_device;     /* one of RETTO_B200_PAGE_* */
    uint64_t n_bytes;      /* HOST_ENCODED only */
} retto_b200_page;

/* RettoWorkerResult (session.rs:42-48) for a batch: box i of page p <-> cls i <-> rec i */
typedef struct retto_b200_page_result {
    int32_t status;
    int32_t first_line;    /* index of this page's first box / line in the flat arrays */
    int32_t n_lines;
} retto_b200_page_result;
typedef struct retto_b200_results {
    int32_t n_pages;
    const retto_b200_page_result* pages;
    int32_t n_lines;
    const retto_b200_box* boxes;            /* original-image coordinates (session.rs:94-97) */
    const retto_b200_cls_result* cls;
    const uint32_t* text_offsets;           /* n_lines + 1 */
    const char* text;
    const float* rec_scores;
} retto_b200_results;
/* RettoSession::process_pipeline (session.rs:75-106) for n pages at once.  The returned pointers are owned by
 * the context and valid until the next run.
 * Host-resident batches are cut into units of pages whose upload overlaps the kernels of earlier units; with
 * retto_b200_set_pipeline(ctx, 2, unit) two units are kept in flight on two internal lanes, software-pipelined
 * from the calling thread.  Per unit the stage callbacks fire in the reference's
 * order (Det, then Cls, then Rec — session.rs:98,101,104), always on the calling thread and never concurrently, but
 * the calls of consecutive units interleave (Det(u0), Det(u1), Cls(u0), Rec(u0), ...) and each call names the CUDA
 * stream its tensors are ordered on (`stream` argument of retto_b200_forward_fn) — the worker must enqueue on that
 * stream.  Units cover consecutive page ranges in page order; results do not depend on the cut. */
retto_b200_status retto_b200_run_pages(retto_b200_ctx* ctx, const retto_b200_page* h_pages, int32_t n_pages,
                                       retto_b200_forward_fn forward, void* user, retto_b200_results* out);

/* Per-stage result delivery == the `callback(RettoWorkerStageResult::{Det,Cls,Rec})` of process_pipeline (session.rs:98,101,104),
 * which RettoSession::run_stream forwards to its mpsc::Sender (session.rs:133-143; consumer retto-wasm/src/wasm_lib.rs:142-189).
 * When a callback is set, run_pages calls it three times per unit of pages, on the calling thread, in the reference's order:
 *   stage 0 (Det) after the boxes are final and rescaled to original-image coordinates, BEFORE the cls forward of that unit;
 *   stage 1 (Cls) after the cls postprocess (flip flags set), BEFORE the rec forward of that unit;
 *   stage 2 (Rec) when the unit's strings are on the host.
 * A unit without detections still reports its three (empty) stages.  The arrays belong to the context and are valid during the
 * call only.  pages[i].first_line indexes the unit's own arrays (box i <-> cls i <-> rec i, as in retto_b200_results).
 * Cost: two extra stream synchronisations per unit (the host must hold the Det / Cls results before it may continue), which is why
 * the batch path leaves the callback unset.  Pass callback = NULL to clear. */
typedef struct retto_b200_stage_result {
    int32_t stage;                          /* 0 Det, 1 Cls, 2 Rec */
    int32_t first_page;                     /* index of the unit's first page in the h_pages array of the run_pages call */
    int32_t n_pages;
    const retto_b200_page_result* pages;    /* n_pages entries */
    int32_t n_lines;
    const retto_b200_box* boxes;            /* stage 0, else NULL */
    const retto_b200_cls_result* cls;       /* stage 1, else NULL */
    const uint32_t* text_offsets;           /* stage 2, else NULL: n_lines + 1 */
    const char* text;                       /* stage 2 */
    const float* rec_scores;                /* stage 2 */
} retto_b200_stage_result;
typedef void (*retto_b200_stage_fn)(void* user, const retto_b200_stage_result* result);
retto_b200_status retto_b200_set_stage_callback(retto_b200_ctx* ctx, retto_b200_stage_fn callback, void* user);

/* run_pages pipeline: lanes = 1 (default: units run back to back on the context's own stream) or 2;
 * unit_pages = pages per unit (0 = default: 64 device-resident / 32 host-resident pages).  0 keeps the default. */
retto_b200_status retto_b200_set_pipeline(retto_b200_ctx* ctx, int32_t lanes, int32_t unit_pages);

/* The same for one batch spread over N contexts — one per GPU of the box (several contexts on one GPU are allowed): pages are
 * independent (session.rs:75-106), so they are sharded per image with no collective.  One host thread per context
 * pulls chunks of `chunk_pages` pages (0 = default) from a shared cursor over the pages sorted by descending H*W and runs
 * retto_b200_run_pages on its context; `forward` is therefore called from N threads concurrently, with users[i] for ctxs[i] (users
 * may be NULL).  Pages must be host-resident.  The result, in page order (== the sequential CLI's order), is owned by ctxs[0]
 * and valid until its next run. */
retto_b200_status retto_b200_run_pages_multi(retto_b200_ctx* const* ctxs, int32_t n_ctx, const retto_b200_page* h_pages, int32_t n_pages,
                                             int32_t chunk_pages, retto_b200_forward_fn forward, void* const* users,
                                             retto_b200_results* out);

/* sizes of the last retto_b200_run_pages call: out8 = {pages, lines, det tensor pixels, crop pixels, cls batch floats,
 * rec batch floats, rec logit rows (sum n*img_w/8), 0} — used by bench.py for algorithmic byte counts */
retto_b200_status retto_b200_last_run_stats(const retto_b200_ctx* ctx, uint64_t* out8);

#ifdef __cplusplus
}
#endif
#endif /* RETTO_B200_H */
