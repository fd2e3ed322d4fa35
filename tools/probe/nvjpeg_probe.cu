// Probe: which nvJPEG backends exist on this box and how fast is batched decode of page-sized JPEGs.
// Usage: nvjpeg_probe blob.bin out_prefix   (blob = u32 n, u32 len[n], bytes...)
#include <cuda_runtime.h>
#include <nvjpeg.h>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("cuda error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)
static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
int main(int argc, char** argv) {
    FILE* f = fopen(argv[1], "rb");
    unsigned n; fread(&n, 4, 1, f);
    std::vector<unsigned> len(n); fread(len.data(), 4, n, f);
    std::vector<std::vector<unsigned char>> img(n);
    for (unsigned i = 0; i < n; i++) { img[i].resize(len[i]); fread(img[i].data(), 1, len[i], f); }
    fclose(f);
    printf("blob: %u images, first %u bytes\n", n, len[0]);
    int backends[] = {3, 2, 0, 1, 5, 4};
    const char* names[] = {"HARDWARE", "GPU_HYBRID", "DEFAULT", "HYBRID", "HARDWARE_DEVICE", "GPU_HYBRID_DEVICE"};
    cudaStream_t st; CK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    for (int bi = 0; bi < 6; bi++) {
        nvjpegHandle_t h; nvjpegStatus_t s = nvjpegCreateEx((nvjpegBackend_t)backends[bi], nullptr, nullptr, 0, &h);
        printf("backend %s: create status %d\n", names[bi], (int)s);
        if (s != NVJPEG_STATUS_SUCCESS) continue;
        nvjpegJpegState_t state; s = nvjpegJpegStateCreate(h, &state);
        if (s) { printf("  state create %d\n", (int)s); nvjpegDestroy(h); continue; }
        int wd[4], ht[4], nc; nvjpegChromaSubsampling_t ss;
        nvjpegGetImageInfo(h, img[0].data(), len[0], &nc, &ss, wd, ht);
        printf("  image info: %dx%d comps %d subsampling %d\n", wd[0], ht[0], nc, (int)ss);
        nvjpegJpegStream_t js; nvjpegJpegStreamCreate(h, &js);
        nvjpegJpegStreamParse(h, img[0].data(), len[0], 0, 0, js);
        int sup = -1; s = nvjpegDecodeBatchedSupported(h, js, &sup);
        printf("  batched supported: status %d is_supported(0=yes) %d\n", (int)s, sup);
        bool device_in = backends[bi] >= 4;
        for (int B : {32, 128}) {
            if ((unsigned)B > n) continue;
            s = nvjpegDecodeBatchedInitialize(h, state, B, 8, NVJPEG_OUTPUT_RGBI);
            if (s) { printf("  batched init B=%d status %d\n", B, (int)s); continue; }
            std::vector<const unsigned char*> ptr(B); std::vector<size_t> ln(B); std::vector<nvjpegImage_t> dst(B);
            std::vector<unsigned char*> pin(B), dbs(B);
            for (int i = 0; i < B; i++) {
                CK(cudaMallocHost(&pin[i], len[i])); memcpy(pin[i], img[i].data(), len[i]);
                if (device_in) { CK(cudaMalloc(&dbs[i], len[i])); CK(cudaMemcpy(dbs[i], pin[i], len[i], cudaMemcpyHostToDevice)); ptr[i] = dbs[i]; }
                else ptr[i] = pin[i];
                ln[i] = len[i];
                memset(&dst[i], 0, sizeof(nvjpegImage_t));
                CK(cudaMalloc(&dst[i].channel[0], (size_t)wd[0] * ht[0] * 3)); dst[i].pitch[0] = wd[0] * 3;
            }
            double best = 1e9; int ok = 1;
            for (int it = 0; it < 6; it++) {
                CK(cudaStreamSynchronize(st));
                double t0 = now();
                s = nvjpegDecodeBatched(h, state, ptr.data(), ln.data(), dst.data(), st);
                if (s) { printf("  decode B=%d status %d\n", B, (int)s); ok = 0; break; }
                double t1 = now();
                CK(cudaStreamSynchronize(st));
                double t2 = now();
                if (it > 0 && t2 - t0 < best) best = t2 - t0;
                if (it == 5) printf("  B=%d: call %.2f ms, +sync %.2f ms\n", B, (t1 - t0) * 1e3, (t2 - t0) * 1e3);
            }
            if (ok) {
                printf("  B=%d best %.2f ms -> %.0f pages/s, %.2f GB/s of RGB out\n", B, best * 1e3, B / best, B / best * wd[0] * ht[0] * 3 / 1e9);
                if (B == 32) {
                    std::vector<unsigned char> out((size_t)wd[0] * ht[0] * 3);
                    CK(cudaMemcpy(out.data(), dst[0].channel[0], out.size(), cudaMemcpyDeviceToHost));
                    char nm[512]; snprintf(nm, 512, "%s_%s.rgb", argv[2], names[bi]);
                    FILE* g = fopen(nm, "wb"); fwrite(out.data(), 1, out.size(), g); fclose(g);
                }
            }
            for (int i = 0; i < B; i++) { cudaFreeHost(pin[i]); if (device_in) cudaFree(dbs[i]); cudaFree(dst[i].channel[0]); }
        }
        nvjpegJpegStreamDestroy(js); nvjpegJpegStateDestroy(state); nvjpegDestroy(h);
    }
    return 0;
}
