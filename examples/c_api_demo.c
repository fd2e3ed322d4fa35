/* The drop-in boundary from plain C99: one synthetic page through retto_b200_run_pages with a stand-in forward callback.
 * Nothing but include/retto_b200.h and libretto_b200.so (the library's own memory helpers: no CUDA headers, no torch).
 *   gcc -std=c99 -Iinclude examples/c_api_demo.c -Lretto_b200 -lretto_b200 -Wl,-rpath,$PWD/retto_b200 -o c_api_demo
 * The callback is where RettoInnerWorker::{det,cls,rec} (retto-core/src/worker.rs:69-73) plugs in: device tensors in, device
 * tensors out, enqueued on the stream the call names.  Here: det = a fixed probability map with three text-line blobs,
 * cls = "upright", rec = the dictionary's classes 1, 2, 3, ... one per time step. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "retto_b200.h"

#define CHECK(call)                                                                                   \
    do {                                                                                              \
        retto_b200_status s_ = (call);                                                                \
        if (s_ != RETTO_B200_OK) {                                                                    \
            fprintf(stderr, "%s -> status %d: %s\n", #call, (int)s_, retto_b200_last_error(g.ctx));   \
            return 1;                                                                                 \
        }                                                                                             \
    } while (0)

static struct {
    retto_b200_ctx* ctx;
    float* d_out[3];      /* worker-owned output buffers, one per stage (grown on demand) */
    size_t cap[3];
    int n_classes;
    int calls[3];
} g;

static int ensure(int stage, size_t bytes) {
    if (bytes <= g.cap[stage]) return 0;
    if (g.d_out[stage]) retto_b200_dev_free(g.ctx, g.d_out[stage]);
    void* p = NULL;
    if (retto_b200_dev_alloc(g.ctx, bytes, &p) != RETTO_B200_OK) return 1;
    g.d_out[stage] = (float*)p;
    g.cap[stage] = bytes;
    return 0;
}

static int32_t forward(void* user, int32_t stage, int32_t n, const retto_b200_tensor* in, retto_b200_tensor* out, void* stream) {
    (void)user; (void)stream;   /* one lane: `stream` is the context's own stream, the one retto_b200_h2d enqueues on */
    g.calls[stage]++;
    if (stage == 0) {           /* det: [1,3,H,W] -> [1,1,H,W] probabilities */
        const int64_t H = in[0].shape[2], W = in[0].shape[3];
        float* h = (float*)malloc((size_t)(H * W) * sizeof(float));
        if (!h || n != 1 || ensure(0, (size_t)(H * W) * sizeof(float))) return 1;
        for (int64_t i = 0; i < H * W; ++i) h[i] = 0.02f;
        for (int k = 0; k < 3; ++k)                          /* three "text lines" */
            for (int64_t y = 60 + 90 * k; y < 92 + 90 * k && y < H; ++y)
                for (int64_t x = 40; x < 40 + 160 * (k + 1) && x < W; ++x) h[y * W + x] = 0.93f;
        if (retto_b200_h2d(g.ctx, g.d_out[0], h, (size_t)(H * W) * sizeof(float)) != RETTO_B200_OK) return 1;
        if (retto_b200_sync(g.ctx) != RETTO_B200_OK) return 1;   /* h is freed below */
        free(h);
        out[0].d_data = g.d_out[0]; out[0].ndim = 4;
        out[0].shape[0] = 1; out[0].shape[1] = 1; out[0].shape[2] = H; out[0].shape[3] = W;
        return 0;
    }
    /* cls: k tensors [m,3,48,192] -> [m,2]; rec: k tensors [m,3,48,W] -> [m,W/8,C] */
    size_t total = 0;
    for (int i = 0; i < n; ++i) total += (size_t)in[i].shape[0] * (stage == 1 ? 2 : (size_t)(in[i].shape[3] / 8) * (size_t)g.n_classes);
    float* h = (float*)calloc(total, sizeof(float));
    if (!h || ensure(stage, total * sizeof(float))) return 1;
    size_t off = 0;
    for (int i = 0; i < n; ++i) {
        const int64_t m = in[i].shape[0];
        out[i].d_data = g.d_out[stage] + off;
        if (stage == 1) {
            for (int64_t r = 0; r < m; ++r) { h[off + 2 * r] = 0.95f; h[off + 2 * r + 1] = 0.05f; }
            out[i].ndim = 2; out[i].shape[0] = m; out[i].shape[1] = 2;
            off += (size_t)m * 2;
        } else {
            const int64_t T = in[i].shape[3] / 8;
            for (int64_t r = 0; r < m; ++r)
                for (int64_t t = 0; t < T; ++t) h[off + (size_t)((r * T + t) * g.n_classes) + (size_t)(t < 6 ? 1 + t : 0)] = 0.9f;
            out[i].ndim = 3; out[i].shape[0] = m; out[i].shape[1] = T; out[i].shape[2] = g.n_classes;
            off += (size_t)(m * T) * (size_t)g.n_classes;
        }
    }
    if (retto_b200_h2d(g.ctx, g.d_out[stage], h, total * sizeof(float)) != RETTO_B200_OK) return 1;
    if (retto_b200_sync(g.ctx) != RETTO_B200_OK) return 1;
    free(h);
    return 0;
}

int main(void) {
    memset(&g, 0, sizeof(g));
    if (retto_b200_abi_version() != RETTO_B200_ABI_VERSION) { fprintf(stderr, "ABI mismatch\n"); return 1; }
    retto_b200_config cfg;
    retto_b200_config_default(&cfg);
    {
        retto_b200_status s = retto_b200_create(0, &cfg, &g.ctx);
        if (s != RETTO_B200_OK) { fprintf(stderr, "retto_b200_create -> %d (no CUDA device?)\n", (int)s); return 2; }
    }
    /* dictionary: 40 one-letter entries (RecCharacter::new adds "blank" in front and " " behind) */
    char dict[256];
    size_t dl = 0;
    for (int i = 0; i < 40; ++i) { dict[dl++] = (char)(i < 26 ? 'a' + i : '0' + (i - 26) % 10); dict[dl++] = '\n'; }
    CHECK(retto_b200_dict_load(g.ctx, dict, dl));
    g.n_classes = retto_b200_dict_size(g.ctx);
    /* one white 736 x 992 page */
    const int H = 736, W = 992;
    uint8_t* rgb = (uint8_t*)malloc((size_t)H * W * 3);
    memset(rgb, 255, (size_t)H * W * 3);
    retto_b200_page page;
    memset(&page, 0, sizeof(page));
    page.rgb = rgb; page.h = H; page.w = W; page.on_device = RETTO_B200_PAGE_HOST_RGB;
    retto_b200_results res;
    CHECK(retto_b200_run_pages(g.ctx, &page, 1, forward, NULL, &res));
    printf("pages %d  lines %d  forward calls det/cls/rec %d/%d/%d  kernels launched %llu\n", res.n_pages, res.n_lines, g.calls[0], g.calls[1], g.calls[2],
           (unsigned long long)retto_b200_launch_count(g.ctx));
    for (int i = 0; i < res.n_lines; ++i) {
        const retto_b200_box* b = &res.boxes[i];
        printf("line %d: box (%.0f,%.0f) (%.0f,%.0f) (%.0f,%.0f) (%.0f,%.0f) score %.3f  cls %d (%.2f)  text \"%.*s\" (%.3f)\n", i, b->xy[0], b->xy[1], b->xy[2],
               b->xy[3], b->xy[4], b->xy[5], b->xy[6], b->xy[7], b->score, res.cls[i].label, res.cls[i].score,
               (int)(res.text_offsets[i + 1] - res.text_offsets[i]), res.text + res.text_offsets[i], res.rec_scores[i]);
    }
    const int ok = res.n_pages == 1 && res.n_lines == 3 && res.pages[0].status == RETTO_B200_OK &&
                   res.text_offsets[1] - res.text_offsets[0] == 6 && memcmp(res.text + res.text_offsets[0], "abcdef", 6) == 0;
    for (int s = 0; s < 3; ++s) if (g.d_out[s]) retto_b200_dev_free(g.ctx, g.d_out[s]);
    retto_b200_destroy(g.ctx);
    free(rgb);
    printf(ok ? "c_api_demo ok\n" : "c_api_demo: unexpected result\n");
    return ok ? 0 : 3;
}
