// jpeg_decode.cu — image decode on the device: ImageHelper::new_from_raw_img_flow
// (retto-core/src/image_helper.rs:34-44: image::load_from_memory(..).to_rgb8()) for baseline JPEG files, the wire format of
// retto-cli (retto-cli/src/main.rs:83-84 reads the file bytes and hands them to RettoSession::run).
//
// Why: a decoded 1280x1280 page is 4.9 MB, its JPEG ~0.2 MB.  With raw RGB pages the end-to-end path is PCIe-bound at ~10 k pages/s
// per GPU and does not scale on a shared host; with file bytes on the wire the upload shrinks 20x and the pixels are
// produced in HBM.  nvJPEG on this box: the hardware backend is unavailable (nvjpegCreateEx(HARDWARE) -> status 7) and the
// GPU-hybrid backend decodes 3.0 k 1280^2 pages/s (profiles/r02_nvjpeg_probe.txt) — below the raw-RGB PCIe path — so the decoder
// is hand-written:
//   host  : marker parse only (tables, frame, scan header: a few hundred bytes per file) -> JpegInfo
//   K-J1  jpeg_scan_kernel   : one block per file turns the entropy-coded segment into a CLEAN bit stream (stuffed zeros and RSTn
//                              markers removed by an ordered block-wide compaction) and records where every restart interval starts
//   K-J2  jpeg_huff_kernel   : one THREAD per restart interval — intervals are independently decodable (DC prediction resets,
//                              byte-aligned) — a 64-bit window of two big-endian words read with one funnel shift per symbol,
//                              words loaded one refill ahead; 10-bit Huffman look-up tables in shared memory whose entries carry
//                              the EXTENDed value when code + magnitude bits fit; non-zero coefficients are scattered into a
//                              pre-zeroed int16 coefficient plane (a file without DRI is one interval = one thread).  Per block the
//                              intervals are sorted by length and spread in tiers: the longest decode alone in their warp (nested
//                              loops), the shortest fill whole warps (see jh_tier_rank)
//   K-J3  jpeg_idct_kernel   : 8 threads per 8x8 block: dequantise + libjpeg's jidctint "islow" integer IDCT (columns, then rows);
//                              DC-only blocks skip both passes
//   K-J4  jpeg_color_kernel  : libjpeg's fancy (triangle) chroma up-sampling h2v1 / h2v2 / h1v2 + YCbCr->RGB fixed point -> HWC u8;
//         jpeg_color420_kernel: the same for batches of 4:2:0 files of one size, two output rows per thread
// Pixel policy: bit-exact with libjpeg-turbo's default decode (what Pillow / OpenCV produce); the oracle for this row is
// oracle/jpeg_oracle.cpp, itself pinned against both libraries (tests/test_cpu_jpeg.py).  The reference's own decoder (zune-jpeg via
// `image 0.25.6`) is an un-vendored dependency; JPEG decoders agree to +-1 LSB, not bit for bit — stated in DESIGN.md.
// Unsupported (status RETTO_B200_ERR_UNSUPPORTED, never a silent fallback): progressive / arithmetic / lossless / 12-bit JPEG,
// CMYK / RGB colour spaces, non-interleaved scans, sampling factors other than 4:4:4 / 4:2:2 / 4:2:0 / 4:4:0 / grayscale, PNG & co.
#include "common.cuh"

#define JPEG_ZZ {0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6, 7, 14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63}
static const uint8_t H_ZIGZAG[64] = JPEG_ZZ;
__constant__ uint8_t D_ZIGZAG[64] = JPEG_ZZ;

// ---- host: marker parse (jdmarker.c) ----------------------------------------------------------------------------------------
retto_b200_status rt_jpeg_parse(const uint8_t* d, size_t n, JpegInfo* out) {
    JpegInfo& J = *out;
    memset(&J, 0, sizeof(J));
    J.status = RETTO_B200_ERR_DECODE;
    if (!d || n < 4 || d[0] != 0xFF || d[1] != 0xD8) { J.status = (d && n >= 8 && d[0] == 0x89 && d[1] == 'P') ? RETTO_B200_ERR_UNSUPPORTED : RETTO_B200_ERR_DECODE; return (retto_b200_status)J.status; }
    bool sof = false, jfif = false, adobe = false;
    int adobe_transform = -1;
    int ids[3] = {0, 0, 0};
    size_t pos = 2;
    for (;;) {
        if (pos + 4 > n || d[pos] != 0xFF) return RETTO_B200_ERR_DECODE;
        while (pos < n && d[pos] == 0xFF) ++pos;
        if (pos >= n) return RETTO_B200_ERR_DECODE;
        const int m = d[pos++];
        if (m == 0xD8 || m == 0x01 || (m >= 0xD0 && m <= 0xD7)) continue;
        if (m == 0xD9 || pos + 2 > n) return RETTO_B200_ERR_DECODE;
        const size_t L = ((size_t)d[pos] << 8) | d[pos + 1];
        if (L < 2 || pos + L > n) return RETTO_B200_ERR_DECODE;
        const uint8_t* s = d + pos + 2;
        const size_t sl = L - 2;
        if (m == 0xDB) {
            size_t i = 0;
            while (i < sl) {
                const int pq = s[i] >> 4, tq = s[i] & 15;
                ++i;
                if (tq > 3 || pq > 1 || i + (pq ? 128 : 64) > sl) return RETTO_B200_ERR_DECODE;
                for (int k = 0; k < 64; ++k) {
                    J.qt[tq][H_ZIGZAG[k]] = (uint16_t)(pq ? ((s[i] << 8) | s[i + 1]) : s[i]);
                    i += pq ? 2 : 1;
                }
                J.qt_present[tq] = 1;
            }
        } else if (m == 0xC0 || m == 0xC1) {
            if (sl < 6) return RETTO_B200_ERR_DECODE;
            if (s[0] != 8) { J.status = RETTO_B200_ERR_UNSUPPORTED; return RETTO_B200_ERR_UNSUPPORTED; }
            J.Y = (s[1] << 8) | s[2]; J.X = (s[3] << 8) | s[4];
            J.nc = s[5];
            if (J.X <= 0 || J.Y <= 0) return RETTO_B200_ERR_DECODE;
            if (J.nc != 1 && J.nc != 3) { J.status = RETTO_B200_ERR_UNSUPPORTED; return RETTO_B200_ERR_UNSUPPORTED; }
            if (sl < (size_t)6 + 3 * J.nc) return RETTO_B200_ERR_DECODE;
            for (int i = 0; i < J.nc; ++i) {
                ids[i] = s[6 + 3 * i]; J.h[i] = s[7 + 3 * i] >> 4; J.v[i] = s[7 + 3 * i] & 15; J.tq[i] = s[8 + 3 * i];
                if (J.h[i] < 1 || J.h[i] > 4 || J.v[i] < 1 || J.v[i] > 4 || J.tq[i] > 3) return RETTO_B200_ERR_DECODE;
            }
            sof = true;
        } else if (m == 0xC4) {
            size_t i = 0;
            while (i < sl) {
                const int tc = s[i] >> 4, th = s[i] & 15;
                ++i;
                if (tc > 1 || th > 3 || i + 16 > sl) return RETTO_B200_ERR_DECODE;
                JpegInfo::Huff& t = tc ? J.ac[th] : J.dc[th];
                int cnt = 0;
                t.bits[0] = 0;
                for (int k = 1; k <= 16; ++k) { t.bits[k] = s[i + k - 1]; cnt += t.bits[k]; }
                i += 16;
                if (cnt > 256 || i + cnt > sl) return RETTO_B200_ERR_DECODE;
                memset(t.vals, 0, sizeof(t.vals));
                memcpy(t.vals, s + i, cnt);
                i += cnt;
                t.present = 1;
            }
        } else if (m >= 0xC2 && m <= 0xCF) {   // progressive, lossless, differential, arithmetic (C4 / C0 / C1 handled above)
            J.status = RETTO_B200_ERR_UNSUPPORTED;
            return RETTO_B200_ERR_UNSUPPORTED;
        } else if (m == 0xDD) {
            if (sl < 2) return RETTO_B200_ERR_DECODE;
            J.ri = (s[0] << 8) | s[1];
        } else if (m == 0xE0) {
            if (sl >= 5 && s[0] == 'J' && s[1] == 'F' && s[2] == 'I' && s[3] == 'F' && s[4] == 0) jfif = true;
        } else if (m == 0xEE) {
            if (sl >= 12 && s[0] == 'A' && s[1] == 'd' && s[2] == 'o' && s[3] == 'b' && s[4] == 'e') { adobe = true; adobe_transform = s[11]; }
        } else if (m == 0xDA) {
            if (!sof || sl < 1) return RETTO_B200_ERR_DECODE;
            const int ns = s[0];
            if (ns != J.nc) { J.status = RETTO_B200_ERR_UNSUPPORTED; return RETTO_B200_ERR_UNSUPPORTED; }   // non-interleaved scans
            if (sl < (size_t)1 + 2 * ns + 3) return RETTO_B200_ERR_DECODE;
            for (int i = 0; i < ns; ++i) {
                if (s[1 + 2 * i] != ids[i]) { J.status = RETTO_B200_ERR_UNSUPPORTED; return RETTO_B200_ERR_UNSUPPORTED; }
                J.td[i] = s[2 + 2 * i] >> 4; J.ta[i] = s[2 + 2 * i] & 15;
                if (J.td[i] > 3 || J.ta[i] > 3) return RETTO_B200_ERR_DECODE;
            }
            pos += L;
            break;
        }
        pos += L;
    }
    J.ecs_off = pos;
    J.ecs_len = n - pos;
    if (J.nc == 3) {   // jdapimin.c default_decompress_parms: JFIF -> YCbCr; Adobe transform 0 -> RGB; ids 'R','G','B' -> RGB
        bool ycc = true;
        if (!jfif) { if (adobe) ycc = adobe_transform != 0; else if (ids[0] == 'R' && ids[1] == 'G' && ids[2] == 'B') ycc = false; }
        if (!ycc) { J.status = RETTO_B200_ERR_UNSUPPORTED; return RETTO_B200_ERR_UNSUPPORTED; }
        if (J.h[1] != 1 || J.v[1] != 1 || J.h[2] != 1 || J.v[2] != 1 || J.h[0] > 2 || J.v[0] > 2) { J.status = RETTO_B200_ERR_UNSUPPORTED; return RETTO_B200_ERR_UNSUPPORTED; }
    } else { J.h[0] = J.v[0] = 1; }   // a single-component scan is non-interleaved: one block per MCU whatever the frame header says
    J.max_h = J.h[0]; J.max_v = J.v[0];
    J.mcux = (J.X + 8 * J.max_h - 1) / (8 * J.max_h);
    J.mcuy = (J.Y + 8 * J.max_v - 1) / (8 * J.max_v);
    for (int i = 0; i < J.nc; ++i)
        if (!J.qt_present[J.tq[i]] || !J.dc[J.td[i]].present || !J.ac[J.ta[i]].present) return RETTO_B200_ERR_DECODE;
    const long long n_mcu = (long long)J.mcux * J.mcuy;
    J.n_seg = J.ri > 0 ? (int)((n_mcu + J.ri - 1) / J.ri) : 1;
    J.status = RETTO_B200_OK;
    return RETTO_B200_OK;
}

extern "C" retto_b200_status retto_b200_image_info(const uint8_t* bytes, uint64_t n_bytes, retto_b200_image_info_t* out) {
    if (!bytes || !out) return RETTO_B200_ERR_INVALID_ARG;
    JpegInfo J;
    const retto_b200_status s = rt_jpeg_parse(bytes, (size_t)n_bytes, &J);
    memset(out, 0, sizeof(*out));
    out->status = s;
    if (s != RETTO_B200_OK) return s;
    out->h = J.Y; out->w = J.X; out->format = 1; out->components = J.nc; out->restart_interval = J.ri;
    out->subsampling = J.nc == 1 ? 3 : (J.max_h == 1 ? (J.max_v == 1 ? 0 : 4) : (J.max_v == 1 ? 1 : 2));
    return RETTO_B200_OK;
}

// ---- device descriptors --------------------------------------------------------------------------------------------------------
struct JpegComp {
    int h, v, tq, td, ta;
    int wb, hb;                   // blocks per row / column (whole MCUs)
    int ds_w, ds_h;               // real (down-sampled) samples
    unsigned coef_base;           // first 8x8 block of this component in the batch's coefficient array
    unsigned long long plane_off; // byte offset of the component's sample plane (stride wb*8)
};
struct JpegDev {
    const uint8_t* ecs;           // device pointer to the entropy-coded data
    unsigned ecs_len;
    int X, Y, nc, max_h, max_v, mcux, mcuy, ri, n_seg;
    int seg_base;                 // first entry of this file in the segment-start table (n_seg entries)
    int thread_base;              // first decode thread of this file (== seg_base)
    unsigned block_base, n_blocks;   // 8x8 blocks of all components (IDCT work list)
    JpegComp c[3];
    unsigned long long clean_off; // byte offset of this file's clean bit stream (K-J1) in the clean arena, 16-byte aligned
    int status_slot;
};
struct JpegHuffRaw { uint8_t bits[17]; uint8_t pad[3]; uint8_t vals[256]; };   // 276 B
struct JpegTables { uint16_t qt[4][64]; JpegHuffRaw dc[4], ac[4]; };

// ---- K-J1: clean bit stream + restart intervals ------------------------------------------------------------------------------
// Inside the entropy-coded data 0xFF is always followed by 0x00 (stuffing) or by a marker, so every "FF D0..D7" pair is an RSTn
// marker.  The kernel copies the data bytes to clean + f.clean_off, dropping stuffed zeros and markers (everything from the EOI
// marker on is dropped too), seg[seg_base + j] = clean byte offset at which restart interval j starts, clean_len[file] = length.
__global__ void __launch_bounds__(256) jpeg_scan_kernel(const JpegDev* __restrict__ files, unsigned* __restrict__ seg, int* __restrict__ status,
                                                        unsigned char* __restrict__ clean, unsigned* __restrict__ clean_len) {
    const JpegDev& f = files[blockIdx.x];
    unsigned* sg = seg + f.seg_base;
    const int n_seg = f.n_seg;
    unsigned char* out = clean + f.clean_off;
    __shared__ int s_wm[8], s_wk[8];
    __shared__ int s_base_m;
    __shared__ unsigned s_base_k, s_eoi;
    if (threadIdx.x == 0) { s_base_m = 0; s_base_k = 0; s_eoi = 0xFFFFFFFFu; }
    __syncthreads();
    const uint8_t* p = f.ecs;
    const unsigned len = f.ecs_len;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (unsigned base = 0; base < len; base += 256 * 16) {
        const unsigned o = base + threadIdx.x * 16;
        // 18 bytes: the last of the previous thread, 16 of this thread, the first of the next
        unsigned char b[18];
#pragma unroll
        for (int k = 0; k < 18; ++k) b[k] = (o + k >= 1 && o + k - 1 < len) ? __ldg(p + o + k - 1) : 0;
        unsigned keep = 0, mark = 0;
        unsigned eoi_at = 0xFFFFFFFFu;
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            if (o + k >= len) break;
            const unsigned c = b[k + 1], prev = b[k], next = b[k + 2];
            const bool rst_ff = c == 0xFF && next >= 0xD0 && next <= 0xD7;
            const bool rst_code = prev == 0xFF && c >= 0xD0 && c <= 0xD7;
            const bool stuffed = prev == 0xFF && c == 0x00;
            if (c == 0xFF && next == 0xD9 && eoi_at == 0xFFFFFFFFu) eoi_at = o + k;
            if (!(rst_ff || rst_code || stuffed)) keep |= 1u << k;
            if (rst_ff) mark |= 1u << k;
        }
        if (eoi_at != 0xFFFFFFFFu) atomicMin(&s_eoi, eoi_at);
        __syncthreads();
        const unsigned eoi = s_eoi;          // first EOI seen so far (this tile or an earlier one): nothing at or after it is data
        if (eoi != 0xFFFFFFFFu) {
#pragma unroll
            for (int k = 0; k < 16; ++k) if (o + k >= eoi) { keep &= ~(1u << k); mark &= ~(1u << k); }
        }
        const int cm = __popc(mark), ck = __popc(keep);
        int im = cm, ik = ck;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int vm = __shfl_up_sync(0xffffffffu, im, d), vk = __shfl_up_sync(0xffffffffu, ik, d);
            if (lane >= d) { im += vm; ik += vk; }
        }
        if (lane == 31) { s_wm[w] = im; s_wk[w] = ik; }
        __syncthreads();
        int bm = s_base_m;
        unsigned bk = s_base_k;
        for (int k = 0; k < w; ++k) { bm += s_wm[k]; bk += (unsigned)s_wk[k]; }
        int idx = bm + im - cm;               // markers before this thread
        unsigned dst = bk + (unsigned)(ik - ck);   // clean bytes before this thread
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            if (keep & (1u << k)) out[dst++] = b[k + 1];
            if (mark & (1u << k)) { ++idx; if (idx < n_seg) sg[idx] = dst; }   // interval idx starts at the current clean position
        }
        __syncthreads();
        if (threadIdx.x == 255) { s_base_m = bm + im; s_base_k = bk + (unsigned)ik; }
        __syncthreads();
    }
    const unsigned total = s_base_k;
    for (int j = threadIdx.x; j < n_seg; j += blockDim.x) {
        if (j == 0) sg[0] = 0;
        else if (j > s_base_m) sg[j] = total;      // missing markers: empty intervals (zeros)
    }
    if (threadIdx.x < 16) out[total + threadIdx.x] = 0;   // zero padding: the bit reader runs a few words past the end
    if (threadIdx.x == 0) {
        clean_len[blockIdx.x] = total;
        if (s_base_m != n_seg - 1) status[f.status_slot] = RETTO_B200_ERR_DECODE;   // marker count does not match DRI
    }
}

// ---- K-J2: Huffman decode, one thread per restart interval ------------------------------------------------------------------------
// The kernel is a set of long serial chains (one per interval; ~1 warp per scheduler), so what counts is the dependent latency and
// the instruction count PER SYMBOL (ncu, first version: 10 cycles per issued instruction, 4.5 active lanes per instruction):
//  * one flat loop over symbols — the lanes of a warp decode different intervals and must not sit in different loop nests;
//  * 10-bit look-up tables whose entries also carry the EXTENDed coefficient when code + magnitude bits fit in the window
//    (entry = total bits | run << 5 | size << 9 | fast << 13 | value << 16), so most symbols cost one shared-memory load;
//  * the bit buffer is refilled 32 bits at a time from words loaded one refill AHEAD (the load latency is off the chain); words
//    are read through an aligned-pair funnel shift so the stream may start at any byte; a word holding 0xFF (stuffing or the
//    marker that ends the interval) takes a byte-wise slow path.
#define JH_LUT_BITS 10
#define JH_SUB_SLOTS 16
struct HuffDev {               // shared memory, one per distinct table of the file
    unsigned lut[1 << JH_LUT_BITS];
    unsigned sub[JH_SUB_SLOTS * 64];   // second level: one 64-entry table (the 6 bits after the window) per 10-bit prefix of the long codes
    int nsub;
    int maxcode[17];           // canonical code ranges (table build, and the fallback when a table has more than JH_SUB_SLOTS long prefixes)
    int valoff[17];            // valptr - mincode
    unsigned char vals[256];
};
#define JH_THREADS 128          // dense layout: every lane decodes
#define JH_T_THREADS 512        // tiered layout: 16 warps per block
#define JH_T_SOLO 8             // ... the first 8 decode ONE restart interval each (lane 0; nested loops)
#define JH_T_IPB 112            // ... intervals per block: 8 x 1 lane, 4 x 4 lanes, 2 x 12 lanes, 2 x 32 lanes
#define JH_IPB_MAX 128
// Tiered layout, rank r (0 = longest interval of the block) of thread (warp w, lane l), -1 = idle lane.  The kernel lasts as long as
// its longest interval's chain, and a warp runs the union of its lanes' paths for as long as its slowest lane: the longest intervals
// get a warp to themselves (no divergence, shortest iteration), the shortest — blank rows of a page: DC + EOB per block — fill whole warps.
__device__ __forceinline__ int jh_tier_rank(int w, int l) {
    if (w < JH_T_SOLO) return l == 0 ? w : -1;
    if (w < 12) return l < 4 ? 8 + (w - 8) * 4 + l : -1;
    if (w < 14) return l < 12 ? 24 + (w - 12) * 12 + l : -1;
    return 48 + (w - 14) * 32 + l;
}
__device__ __forceinline__ int jh_extend(int v, int s) { return v < (1 << (s - 1)) ? v - (1 << s) + 1 : v; }
__device__ __forceinline__ unsigned jh_has_ff(unsigned w) { return __vcmpeq4(w, 0xFFFFFFFFu); }

// Shared-memory decode tables of one file (canonical codes, jdhuff.c jpeg_make_d_derived_tbl): s_tab[2 * ci] = DC table of component
// ci, [2 * ci + 1] = its AC table (s_slot maps tables shared by several components onto one copy).  Whole block; ends with a barrier.
__device__ __forceinline__ void jh_build_tables(const JpegDev& f, const JpegTables& T, HuffDev* s_tab, unsigned char* s_zz, int* s_slot) {
    if (threadIdx.x < 64) s_zz[threadIdx.x] = D_ZIGZAG[threadIdx.x];
    // distinct tables -> shared memory (canonical codes, jdhuff.c jpeg_make_d_derived_tbl)
    for (int t = 0; t < 2 * f.nc; ++t) {
        const int ci = t >> 1, is_ac = t & 1;
        const int id = is_ac ? f.c[ci].ta : f.c[ci].td;
        int alias = -1;
        for (int c2 = 0; c2 < ci; ++c2) if ((is_ac ? f.c[c2].ta : f.c[c2].td) == id) { alias = 2 * c2 + is_ac; break; }
        if (threadIdx.x == 0) s_slot[t] = alias >= 0 ? alias : t;
        if (alias >= 0) continue;   // block-uniform
        const JpegHuffRaw& raw = is_ac ? T.ac[id] : T.dc[id];
        HuffDev& H = s_tab[t];
        if (threadIdx.x == 0) {
            int code = 0, k = 0;
            for (int l = 1; l <= 16; ++l) {
                H.valoff[l] = k - code;
                code += raw.bits[l];
                k += raw.bits[l];
                H.maxcode[l] = raw.bits[l] ? code - 1 : -1;
                code <<= 1;
            }
            H.maxcode[0] = -1; H.valoff[0] = 0; H.nsub = 0;
        }
        for (int i = threadIdx.x; i < 256; i += blockDim.x) H.vals[i] = raw.vals[i];
        __syncthreads();
        for (int i = threadIdx.x; i < (1 << JH_LUT_BITS); i += blockDim.x) {
            unsigned e = 0;
            for (int l = 1; l <= JH_LUT_BITS; ++l) {
                const int code = i >> (JH_LUT_BITS - l);
                if (code <= H.maxcode[l]) {
                    const unsigned sym = H.vals[(H.valoff[l] + code) & 255];
                    const unsigned r = is_ac ? sym >> 4 : 0u, sz = sym & 15u;
                    if (l + (int)sz <= JH_LUT_BITS) {
                        const int bits = (i >> (JH_LUT_BITS - l - (int)sz)) & ((1 << sz) - 1);
                        const int v = sz ? jh_extend(bits, (int)sz) : 0;
                        e = (unsigned)(l + sz) | (r << 5) | (sz << 9) | (1u << 13) | ((unsigned)(v & 0xFFFF) << 16);
                    } else e = (unsigned)l | (r << 5) | (sz << 9);
                    break;
                }
            }
            if (e == 0) {   // prefix of codes longer than the window: second-level table over the next 6 bits
                const int slot = atomicAdd(&H.nsub, 1);
                e = (unsigned)(slot < JH_SUB_SLOTS ? slot : 0xFFFF) << 16;
                if (slot < JH_SUB_SLOTS)
                    for (int t2 = 0; t2 < 64; ++t2) {
                        const int code16 = (i << 6) | t2;
                        unsigned e2 = 16u;   // no such code: skip 16 bits as symbol 0 (jdhuff.c: corrupt data)
                        for (int l = JH_LUT_BITS + 1; l <= 16; ++l) {
                            const int code = code16 >> (16 - l);
                            if (code <= H.maxcode[l]) {
                                const unsigned sym = H.vals[(H.valoff[l] + code) & 255];
                                e2 = (unsigned)l | ((is_ac ? sym >> 4 : 0u) << 5) | ((sym & 15u) << 9);
                                break;
                            }
                        }
                        H.sub[slot * 64 + t2] = e2;
                    }
            }
            H.lut[i] = e;
        }
    }
    __syncthreads();
}

// Bit reader over the CLEAN stream (K-J1): w0:w1 = the next 64 bits (big-endian words), o = bits of w0 already consumed;
// the 32-bit window at the read position is one funnel shift; `nxt` is the raw word after w1, loaded one refill ahead.
__global__ void __launch_bounds__(JH_T_THREADS) jpeg_huff_kernel(const JpegDev* __restrict__ files, const int* __restrict__ block_file,
                                                              const int* __restrict__ block_first, const JpegTables* __restrict__ tables,
                                                              const unsigned* __restrict__ seg, const unsigned char* __restrict__ clean,
                                                              const unsigned* __restrict__ clean_len, short* __restrict__ coef, int tiered, const unsigned* __restrict__ wend) {
    extern __shared__ __align__(16) unsigned char jh_smem[];
    HuffDev* s_tab = reinterpret_cast<HuffDev*>(jh_smem);   // [2 * ci] = DC table of component ci, [2 * ci + 1] = its AC table (aliased when shared)
    __shared__ unsigned char s_zz[64];
    __shared__ int s_slot[6];
    const int fi = block_file[blockIdx.x];
    const JpegDev& f = files[fi];
    const JpegTables& T = tables[fi];
    jh_build_tables(f, T, s_tab, s_zz, s_slot);
    // The block's restart intervals, longest first (their lengths in the clean stream are the work: ncu showed the kernel at 0.14 IPC,
    // one warp per scheduler, 100 instructions per symbol iteration — the union of 32 lanes' paths — and as long as the longest
    // interval of the batch, 4.4x the mean on text pages).  Lanes of a warp then hold intervals of similar length, and in the tiered
    // layout the longest ones decode alone in their warp.
    __shared__ unsigned s_len[JH_IPB_MAX];
    __shared__ unsigned char s_order[JH_IPB_MAX];
    const unsigned clen = clean_len[fi];
    const int j0 = block_first[blockIdx.x];
    const int cnt = min(tiered ? JH_T_IPB : JH_THREADS, f.n_seg - j0);
    if ((int)threadIdx.x < cnt) {
        const int jj = j0 + threadIdx.x;
        const unsigned a0 = min(seg[f.seg_base + jj], clen), a1 = jj + 1 < f.n_seg ? min(seg[f.seg_base + jj + 1], clen) : clen;
        s_len[threadIdx.x] = a1 > a0 ? a1 - a0 : 0u;
    }
    __syncthreads();
    if ((int)threadIdx.x < cnt) {
        const unsigned me = s_len[threadIdx.x];
        int rank = 0;
        for (int u = 0; u < cnt; ++u) { const unsigned o = s_len[u]; rank += (o > me || (o == me && u < (int)threadIdx.x)) ? 1 : 0; }
        s_order[rank] = (unsigned char)threadIdx.x;
    }
    __syncthreads();
    const int rk = tiered ? jh_tier_rank(threadIdx.x >> 5, threadIdx.x & 31) : (int)threadIdx.x;
    if (rk < 0 || rk >= cnt) return;
    const int j = j0 + s_order[rk];   // restart interval of this thread
    const unsigned char* base = clean + f.clean_off;             // 16-byte aligned
    const unsigned start = min(seg[f.seg_base + j], clen);
    // word index into the file's clean stream; wmax = the last readable word of the arena: the look-ahead index is clamped to it, so a
    // truncated / damaged stream that runs on into whatever follows never reads past the allocation (32-bit min: a clamp of the
    // 64-bit pointer cost the parse kernels a factor of two)
    const unsigned* wb = reinterpret_cast<const unsigned*>(base);
    const unsigned wmax = (unsigned)(wend - wb);
    unsigned wi = start >> 2;
    unsigned w0 = __byte_perm(__ldg(wb + wi), 0, 0x0123), w1 = __byte_perm(__ldg(wb + wi + 1), 0, 0x0123);
    unsigned nxt = __ldg(wb + wi + 2);
    wi += 3;
    unsigned o = 8u * (start & 3u);
    const long long n_mcu = (long long)f.mcux * f.mcuy;
    long long m = f.ri > 0 ? (long long)j * f.ri : 0;
    const long long m1 = f.ri > 0 ? min(m + f.ri, n_mcu) : n_mcu;
    int my = (int)(m / f.mcux), mx = (int)(m - (long long)my * f.mcux);
    const int nb0 = f.c[0].h * f.c[0].v, nb = f.nc == 1 ? 1 : nb0 + 2;   // blocks per MCU: luma blocks, then Cb, Cr
    // per-component constants in registers
    const int h0 = f.c[0].h, v0 = f.c[0].v;
    // (three-way selects instead of arrays indexed by ci: a dynamically indexed array would live in local memory)
    const int c1 = f.nc > 1 ? 1 : 0, c2 = f.nc > 2 ? 2 : 0;
    const HuffDev* tdc0 = &s_tab[s_slot[0]];
    const HuffDev* tac0 = &s_tab[s_slot[1]];
    const HuffDev* tdc1 = &s_tab[s_slot[2 * c1]];
    const HuffDev* tac1 = &s_tab[s_slot[2 * c1 + 1]];
    const HuffDev* tdc2 = &s_tab[s_slot[2 * c2]];
    const HuffDev* tac2 = &s_tab[s_slot[2 * c2 + 1]];
    const unsigned cbase0 = f.c[0].coef_base, cbase1 = f.c[c1].coef_base, cbase2 = f.c[c2].coef_base;
    const int cwb0 = f.c[0].wb, cwb1 = f.c[c1].wb, cwb2 = f.c[c2].wb;
    int pred0 = 0, pred1 = 0, pred2 = 0;
    int b = 0, k = 0, ci = 0;
    short* blk = coef + ((size_t)cbase0 + (size_t)(my * v0) * cwb0 + mx * h0) * 64;
    const HuffDev* tab = tdc0;
    if (tiered && (threadIdx.x >> 5) < JH_T_SOLO) {
        // A warp with ONE decoding lane (the longest intervals of the block): nothing to keep in step with, so the decoder is
        // written as libjpeg's nested loops — MCU, block, one DC symbol, the AC symbols — instead of the flat one-symbol-per-iteration
        // loop below, whose block / table switches and three-way symbol branch sit inside every iteration.
        // (Measured and dropped: all 32 lanes of such a warp decoding the 32 possible next bit offsets, the chain of real symbol
        // starts walked with one shuffle per symbol — bit-exact, but 640 cycles per symbol against 310 here: rounds end at every
        // block boundary (the next code is a DC code of another table), text pages average a handful of symbols per block, and the
        // per-round set-up is serial code at 0.2 IPC.)
        auto symbol = [&](const HuffDev* t, bool is_dc, int& r, int& sz, int& v) {
            const unsigned win = __funnelshift_l(w1, w0, o);
            unsigned e = t->lut[win >> (32 - JH_LUT_BITS)];
            if ((e & 31u) == 0) {
                const unsigned slot = e >> 16;
                if (slot < JH_SUB_SLOTS) e = t->sub[slot * 64 + ((win >> 16) & 63u)];
                else {
                    const unsigned top = win >> 16;
                    unsigned sym = 0, l = 16;
#pragma unroll 1
                    for (int q = JH_LUT_BITS + 1; q <= 16; ++q) {
                        const int code = (int)(top >> (16 - q));
                        if (code <= t->maxcode[q]) { sym = t->vals[(t->valoff[q] + code) & 255]; l = q; break; }
                    }
                    e = is_dc ? (l | ((sym & 15u) << 9)) : (l | ((sym >> 4) << 5) | ((sym & 15u) << 9));
                }
            }
            unsigned L = e & 31u;
            r = (int)((e >> 5) & 15u); sz = (int)((e >> 9) & 15u);
            v = (int)e >> 16;
            if (!(e & (1u << 13)) && sz) {
                v = jh_extend((int)((win << L) >> (32 - sz)), sz);
                L += (unsigned)sz;
            }
            o += L;
            if (o >= 32u) {
                o -= 32u;
                w0 = w1;
                w1 = __byte_perm(nxt, 0, 0x0123);
                nxt = __ldg(wb + wi);
                wi = min(wi + 1u, wmax);
            }
        };
        for (; m < m1; ++m) {
            for (int bb = 0; bb < nb; ++bb) {
                const int cc = max(0, bb - nb0 + 1);
                const int bi = cc == 0 ? bb : 0;
                const int dy = h0 == 2 ? (bi >> 1) : bi, dx = h0 == 2 ? (bi & 1) : 0;
                const int hh = cc == 0 ? h0 : 1, vv = cc == 0 ? v0 : 1;
                const unsigned cb_ = cc == 0 ? cbase0 : (cc == 1 ? cbase1 : cbase2);
                const int wb_ = cc == 0 ? cwb0 : (cc == 1 ? cwb1 : cwb2);
                short* bp = coef + ((size_t)cb_ + (size_t)(my * vv + dy) * wb_ + (mx * hh + dx)) * 64;
                const HuffDev* tdc = cc == 0 ? tdc0 : (cc == 1 ? tdc1 : tdc2);
                const HuffDev* tac = cc == 0 ? tac0 : (cc == 1 ? tac1 : tac2);
                int r, sz, v;
                symbol(tdc, true, r, sz, v);
                int pv;
                if (cc == 0) { pred0 += v; pv = pred0; } else if (cc == 1) { pred1 += v; pv = pred1; } else { pred2 += v; pv = pred2; }
                if (pv) bp[0] = (short)pv;
                int kk = 1;
                while (kk < 64) {
                    symbol(tac, false, r, sz, v);
                    if (sz) {
                        kk += r;
                        if (kk < 64) bp[s_zz[kk]] = (short)v;
                        ++kk;
                    } else kk = r == 15 ? kk + 16 : 64;
                }
            }
            if (++mx == f.mcux) { mx = 0; ++my; }
        }
        return;
    }
    while (m < m1) {
        const unsigned win = __funnelshift_l(w1, w0, o);          // the 32 bits at the read position
        unsigned e = tab->lut[win >> (32 - JH_LUT_BITS)];
        if ((e & 31u) == 0) {   // a code longer than the window
            const unsigned slot = e >> 16;
            if (slot < JH_SUB_SLOTS) e = tab->sub[slot * 64 + ((win >> 16) & 63u)];
            else {              // more long prefixes than sub-tables (never with the standard tables): canonical search
                const unsigned top = win >> 16;
                unsigned sym = 0, l = 16;
#pragma unroll 1
                for (int q = JH_LUT_BITS + 1; q <= 16; ++q) {
                    const int code = (int)(top >> (16 - q));
                    if (code <= tab->maxcode[q]) { sym = tab->vals[(tab->valoff[q] + code) & 255]; l = q; break; }
                }
                e = k == 0 ? (l | ((sym & 15u) << 9)) : (l | ((sym >> 4) << 5) | ((sym & 15u) << 9));
            }
        }
        unsigned L = e & 31u;
        const int r = (int)((e >> 5) & 15u), sz = (int)((e >> 9) & 15u);
        int v = (int)e >> 16;
        if (!(e & (1u << 13)) && sz) {   // magnitude bits outside the table window: still inside `win` (code <= 16, size <= 15 bits)
            v = jh_extend((int)((win << L) >> (32 - sz)), sz);
            L += (unsigned)sz;
        }
        o += L;
        if (o >= 32u) {
            o -= 32u;
            w0 = w1;
            w1 = __byte_perm(nxt, 0, 0x0123);
            nxt = __ldg(wb + wi);     // unconditional: a select on the VALUE made the compiler consume the load at once; the INDEX is clamped instead
            wi = min(wi + 1u, wmax);
        }
        if (k == 0) {
            int pv;
            if (ci == 0) { pred0 += v; pv = pred0; } else if (ci == 1) { pred1 += v; pv = pred1; } else { pred2 += v; pv = pred2; }
            if (pv) blk[0] = (short)pv;
            k = 1;
            tab = ci == 0 ? tac0 : (ci == 1 ? tac1 : tac2);
        } else if (sz) {
            k += r;
            if (k < 64) blk[s_zz[k]] = (short)v;
            ++k;
        } else {
            k = r == 15 ? k + 16 : 64;
        }
        if (k >= 64) {   // next block of the MCU / next MCU
            k = 0;
            if (++b == nb) { b = 0; ++m; if (++mx == f.mcux) { mx = 0; ++my; } }
            ci = max(0, b - nb0 + 1);
            const int bi = ci == 0 ? b : 0;
            const int dy = h0 == 2 ? (bi >> 1) : bi, dx = h0 == 2 ? (bi & 1) : 0;
            const int hh = ci == 0 ? h0 : 1, vv = ci == 0 ? v0 : 1;
            const unsigned cb_ = ci == 0 ? cbase0 : (ci == 1 ? cbase1 : cbase2);
            const int wb_ = ci == 0 ? cwb0 : (ci == 1 ? cwb1 : cwb2);
            blk = coef + ((size_t)cb_ + (size_t)(my * vv + dy) * wb_ + (mx * hh + dx)) * 64;
            tab = ci == 0 ? tdc0 : (ci == 1 ? tdc1 : tdc2);
        }
    }
}

// ---- K-J2s: Huffman decode of SUB-SEQUENCES of a restart interval (self-synchronising parse) ------------------------------------
// A restart interval is one serial chain, and a GPU thread walks it at ~310 cycles per symbol: a file without restart markers is ONE
// chain (42 ms for a 1280x1280 page), and with one interval per MCU row the longest row of a text page is the kernel's duration.
// Huffman streams re-synchronise: a decoder started at an arbitrary bit, in an arbitrary state, soon falls in step with the true
// parse (Klein & Wiseman; Weissenberger & Schmidt, "Massively parallel Huffman decoding on GPUs", for JPEG: the state also holds
// the position inside the MCU).  So every interval longer than JS_SUB_BYTES is cut into sub-sequences of JS_SUB_BYTES, one THREAD each:
//   jpeg_sub_kernel    slot table: sub-sequence k of interval j of file f (the counts are only known on the device, after K-J1)
//   jpeg_sync_kernel<1> every slot parses its own region from its first bit AS IF an MCU started there (lengths only: no values, no
//                      stores) and records, in a bitmap over its region, the bit positions at which ITS parse starts an MCU, plus a
//                      running count per bitmap word
//   jpeg_sync_kernel<2> every slot keeps parsing into the following regions until it stands at an MCU start that the slot owning that
//                      region ALSO recorded: from that bit on the two parses are identical.  Slot 0's parse is the true one from
//                      its first bit, so by induction the chain of matches gives every slot a true start
//   jpeg_resolve_kernel one thread per interval walks that chain: true start bit and MCU index of every slot (a slot whose region the
//                      true parse crossed without a common MCU start takes over at the first MCU start the true parse made inside
//                      its region — the look-ahead's check points; an interval whose chain breaks is decoded by its first slot
//                      alone, from bit 0 — correct, just serial)
//   jpeg_huff_sub_kernel the real decode (values, coefficient stores) of every slot's MCU range, DC predictors starting at 0
//   jpeg_dcfix_kernel  adds to the DC terms of a slot's blocks the predictors carried in from the slots before it
#define JS_SUB_BYTES 1024
#define JS_SUB_BITS (JS_SUB_BYTES * 8)
#define JS_WORDS (JS_SUB_BITS / 32)
#define JS_MIN_MEAN_INTERVAL 4096   // a file goes through K-J2s when its mean restart interval is longer than this many bytes
#define JS_CK 8                    // check points a slot keeps (one per region its look-ahead crosses without meeting that region's parse)
#define JS_MAX_OVERLAP 4096     // regions a parse may cross while looking for a common MCU start before the interval falls back to one serial chain
struct JsSlot { int j, k, nsub, pad; };                 // interval (index inside the file), sub-sequence, sub-sequences of the interval
struct JsState { unsigned p; int b, kk, marked; };      // after pass 1: bit position (relative to the interval), block in MCU, zig-zag index, MCU starts recorded
struct JsSync { int t; unsigned X; int C, pad; };       // written by slot s: first later slot t of the interval it met (-1: none, -2: gave up), at bit X; C = MCU starts of s's parse before X
struct JsCk { unsigned bit; int count; };                // where a slot's look-ahead parse first starts an MCU inside a later region, and its MCU-start count before that bit
struct JsStart { unsigned bit; int mcu, n_mcu, pad; };  // true start of a slot: bit (relative to the interval), first MCU (relative to the interval), MCUs to decode

__global__ void __launch_bounds__(256) jpeg_sub_kernel(const JpegDev* __restrict__ files, const unsigned* __restrict__ seg, const unsigned* __restrict__ clean_len,
                                                       const int* __restrict__ sub_base, int* __restrict__ isub, JsSlot* __restrict__ slots) {
    const int fi = blockIdx.x;
    const JpegDev& f = files[fi];
    const unsigned clen = clean_len[fi];
    const int first = sub_base[fi], cap = sub_base[fi + 1] - first;
    if (cap == 0) return;   // a file K-J2 decodes
    __shared__ int s_scan[256];
    __shared__ int s_run;
    if (threadIdx.x == 0) s_run = 0;
    __syncthreads();
    for (int j0 = 0; j0 < f.n_seg; j0 += 256) {
        const int j = j0 + threadIdx.x;
        int ns = 0;
        if (j < f.n_seg) {
            const unsigned a0 = min(seg[f.seg_base + j], clen), a1 = j + 1 < f.n_seg ? min(seg[f.seg_base + j + 1], clen) : clen;
            const unsigned len = a1 > a0 ? a1 - a0 : 0u;
            ns = max(1, (int)((len + JS_SUB_BYTES - 1) / JS_SUB_BYTES));
        }
        s_scan[threadIdx.x] = ns;
        __syncthreads();
        for (int d = 1; d < 256; d <<= 1) {
            const int v = (int)threadIdx.x >= d ? s_scan[threadIdx.x - d] : 0;
            __syncthreads();
            s_scan[threadIdx.x] += v;
            __syncthreads();
        }
        const int at = s_run + s_scan[threadIdx.x] - ns;
        if (j < f.n_seg) {
            isub[f.seg_base + j] = first + at;
            for (int k = 0; k < ns; ++k) if (at + k < cap) slots[first + at + k] = JsSlot{j, k, ns, 0};
        }
        __syncthreads();
        if (threadIdx.x == 255) s_run += s_scan[255];
        __syncthreads();
    }
    for (int i = s_run + threadIdx.x; i < cap; i += 256) slots[first + i] = JsSlot{-1, 0, 0, 0};
}

// One symbol of the parse: total length in bits (code + magnitude), run, size.  t = the table the state selects.
__device__ __forceinline__ void js_symbol(const HuffDev* t, bool is_dc, unsigned win, unsigned& L, int& r, int& sz) {
    unsigned e = t->lut[win >> (32 - JH_LUT_BITS)];
    if ((e & 31u) == 0) {
        const unsigned slot = e >> 16;
        if (slot < JH_SUB_SLOTS) e = t->sub[slot * 64 + ((win >> 16) & 63u)];
        else {
            const unsigned top = win >> 16;
            unsigned sym = 0, l = 16;
#pragma unroll 1
            for (int q = JH_LUT_BITS + 1; q <= 16; ++q) {
                const int code = (int)(top >> (16 - q));
                if (code <= t->maxcode[q]) { sym = t->vals[(t->valoff[q] + code) & 255]; l = q; break; }
            }
            e = is_dc ? (l | ((sym & 15u) << 9)) : (l | ((sym >> 4) << 5) | ((sym & 15u) << 9));
        }
    }
    r = (int)((e >> 5) & 15u); sz = (int)((e >> 9) & 15u);
    L = (e & 31u) + ((e & (1u << 13)) ? 0u : (unsigned)sz);
}

template <int PHASE>
__global__ void __launch_bounds__(128) jpeg_sync_kernel(const JpegDev* __restrict__ files, const int* __restrict__ block_file, const int* __restrict__ block_first,
                                                        const JpegTables* __restrict__ tables, const unsigned* __restrict__ seg,
                                                        const unsigned char* __restrict__ clean, const unsigned* __restrict__ clean_len,
                                                        const JsSlot* __restrict__ slots, unsigned* __restrict__ bitmaps, unsigned short* __restrict__ counts,
                                                        JsState* __restrict__ states, JsSync* __restrict__ syncs, JsCk* __restrict__ cks, const unsigned* __restrict__ wend) {
    extern __shared__ __align__(16) unsigned char jh_smem[];
    HuffDev* s_tab = reinterpret_cast<HuffDev*>(jh_smem);
    __shared__ unsigned char s_zz[64];
    __shared__ int s_slot[6];
    const int fi = block_file[blockIdx.x];
    const JpegDev& f = files[fi];
    jh_build_tables(f, tables[fi], s_tab, s_zz, s_slot);
    const int slot = block_first[blockIdx.x] + threadIdx.x;
    if (slot >= block_first[blockIdx.x + 1]) return;   // block_first holds one more entry: the end of the last block's file
    const JsSlot sl = slots[slot];
    if (sl.j < 0 || sl.nsub <= 1) return;               // single-slot intervals need no parse
    if (PHASE == 2 && sl.k == sl.nsub - 1) { syncs[slot] = JsSync{-1, 0u, 0, 0}; return; }
    const unsigned clen = clean_len[fi];
    const unsigned a0 = min(seg[f.seg_base + sl.j], clen), a1 = sl.j + 1 < f.n_seg ? min(seg[f.seg_base + sl.j + 1], clen) : clen;
    const unsigned len_bits = 8u * (a1 > a0 ? a1 - a0 : 0u);
    const unsigned r0 = (unsigned)sl.k * JS_SUB_BITS, r1 = min(r0 + JS_SUB_BITS, len_bits);
    const int nb0 = f.c[0].h * f.c[0].v, nb = f.nc == 1 ? 1 : nb0 + 2;
    const int c1 = f.nc > 1 ? 1 : 0, c2 = f.nc > 2 ? 2 : 0;
    const HuffDev* tdc0 = &s_tab[s_slot[0]];
    const HuffDev* tac0 = &s_tab[s_slot[1]];
    const HuffDev* tdc1 = &s_tab[s_slot[2 * c1]];
    const HuffDev* tac1 = &s_tab[s_slot[2 * c1 + 1]];
    const HuffDev* tdc2 = &s_tab[s_slot[2 * c2]];
    const HuffDev* tac2 = &s_tab[s_slot[2 * c2 + 1]];
    unsigned p;
    int b, kk;
    if (PHASE == 1) { p = r0; b = 0; kk = 0; }
    else { const JsState st = states[slot]; p = st.p; b = st.b; kk = st.kk; }
    // bit reader at absolute bit 8 * a0 + p of the file's clean stream (16-byte aligned base)
    const unsigned char* base = clean + f.clean_off;
    const unsigned long long abit = 8ull * a0 + p;
    const unsigned* wb = reinterpret_cast<const unsigned*>(base);
    const unsigned wmax = (unsigned)(wend - wb);
    unsigned wi = (unsigned)(abit >> 5);
    unsigned w0 = __byte_perm(__ldg(wb + wi), 0, 0x0123), w1 = __byte_perm(__ldg(wb + wi + 1), 0, 0x0123);
    unsigned nxt = __ldg(wb + wi + 2);
    wi += 3;
    unsigned o = (unsigned)(abit & 31ull);
    // PHASE 1: bitmap of this slot's MCU starts + running count per word
    unsigned* bm = bitmaps + (size_t)slot * JS_WORDS;
    unsigned short* cnt = counts + (size_t)slot * JS_WORDS;
    int cw = 0, run = 0;
    unsigned cur = 0;
    int extra = 0;                       // PHASE 2: MCU starts of this parse at or after r1 that were not common
    int last_ck_t = sl.k;                // PHASE 2: last region a check point was stored for
    const int marked0 = PHASE == 2 ? states[slot].marked : 0;
    if (PHASE == 2) {
#pragma unroll
        for (int i = 0; i < JS_CK; ++i) cks[(size_t)slot * JS_CK + i] = JsCk{0xFFFFFFFFu, 0};
    }
    int found_t = -1;
    unsigned found_x = 0;
    bool gave_up = false;
    for (;;) {
        if (b == 0 && kk == 0) {         // an MCU starts at p
            if (PHASE == 1) {
                if (p >= r1) break;
                const int d = (int)(p - r0), wd = d >> 5;
                while (cw < wd) { bm[cw] = cur; cnt[cw] = (unsigned short)run; run += __popc(cur); cur = 0; ++cw; }
                cur |= 1u << (d & 31);
            } else if (p >= r1) {
                if (p >= len_bits) break;                                   // the interval ends: this parse runs to its end
                const int t = (int)(p / JS_SUB_BITS);                     // the region p lies in
                if (t - sl.k > JS_MAX_OVERLAP) { gave_up = true; break; }
                const int d = (int)(p - (unsigned)t * JS_SUB_BITS);
                if ((bitmaps[(size_t)(slot + (t - sl.k)) * JS_WORDS + (d >> 5)] >> (d & 31)) & 1u) { found_t = t; found_x = p; break; }
                // region t's own parse is not in step here.  If THIS parse turns out to be the true one (jpeg_resolve_kernel), slot t can
                // still take over at the first MCU start inside its region: remember that bit and the MCU count before it
                if (t > last_ck_t) {
                    if (t - sl.k - 1 < JS_CK) cks[(size_t)slot * JS_CK + (t - sl.k - 1)] = JsCk{p, marked0 + extra};
                    last_ck_t = t;
                }
                ++extra;
            }
        } else if (PHASE == 1 && p >= r1) break;   // mid-MCU at the end of the region: pass 2 continues from this state
        if (p >= len_bits) break;                  // out of data
        const unsigned win = __funnelshift_l(w1, w0, o);
        const int ci = max(0, b - nb0 + 1);
        const HuffDev* t = kk == 0 ? (ci == 0 ? tdc0 : (ci == 1 ? tdc1 : tdc2)) : (ci == 0 ? tac0 : (ci == 1 ? tac1 : tac2));
        unsigned L;
        int r, sz;
        js_symbol(t, kk == 0, win, L, r, sz);
        p += L;
        o += L;
        if (o >= 32u) {
            o -= 32u;
            w0 = w1;
            w1 = __byte_perm(nxt, 0, 0x0123);
            nxt = __ldg(wb + wi);
            wi = min(wi + 1u, wmax);
        }
        if (kk == 0) kk = 1;
        else if (sz) kk += r + 1;
        else kk = r == 15 ? kk + 16 : 64;
        if (kk >= 64) { kk = 0; if (++b == nb) b = 0; }
    }
    if (PHASE == 1) {
        while (cw < JS_WORDS) { bm[cw] = cur; cnt[cw] = (unsigned short)run; run += __popc(cur); cur = 0; ++cw; }
        states[slot] = JsState{p, b, kk, run};
    } else {
        syncs[slot] = JsSync{gave_up ? -2 : found_t, found_x, marked0 + extra, 0};
    }
}

// One thread per restart interval with more than one slot: the chain of common MCU starts -> true start of every slot.
__global__ void jpeg_resolve_kernel(const JpegDev* __restrict__ files, const int* __restrict__ seg_file, int total_seg, const int* __restrict__ sub_base,
                                    const int* __restrict__ isub,
                                    const JsSlot* __restrict__ slots, const unsigned* __restrict__ bitmaps, const unsigned short* __restrict__ counts,
                                    const JsSync* __restrict__ syncs, const JsCk* __restrict__ cks, JsStart* __restrict__ starts) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;   // global interval index
    if (g >= total_seg) return;
    const int fi = seg_file[g];
    if (sub_base[fi + 1] == sub_base[fi]) return;   // a file K-J2 decodes
    const JpegDev& f = files[fi];
    const int j = g - f.seg_base;
    const int first = isub[g];
    const int nsub = slots[first].nsub;
    const long long n_mcu_file = (long long)f.mcux * f.mcuy;
    const long long m0 = f.ri > 0 ? (long long)j * f.ri : 0;
    const int total = (int)max(0ll, min(f.ri > 0 ? (long long)f.ri : n_mcu_file, n_mcu_file - m0));   // MCUs of this interval
    if (nsub <= 1) { starts[first] = JsStart{0u, 0, total, 0}; return; }
    for (int s = 0; s < nsub; ++s) starts[first + s] = JsStart{0u, 0, 0, 0};
    int s = 0, m = 0;
    unsigned X = 0;
    bool ok = true;
    while (true) {
        const JsSync sy = syncs[first + s];
        if (sy.t == -2) { ok = false; break; }
        const int d = (int)(X - (unsigned)s * JS_SUB_BITS);                  // X lies in slot s's own region (it was found in its bitmap)
        const int before = (int)counts[(size_t)(first + s) * JS_WORDS + (d >> 5)] + __popc(bitmaps[(size_t)(first + s) * JS_WORDS + (d >> 5)] & ((1u << (d & 31)) - 1u));
        const bool last = sy.t < 0;
        const int n_link = last ? max(total - m, 0) : sy.C - before;        // MCUs between X_s and the next common start (or the interval's end)
        if (n_link < 0 || m + n_link > total || (!last && (sy.t <= s || sy.t >= nsub))) { ok = false; break; }
        // slot s's parse is the true one from X_s on.  The slots whose regions it crossed without a common start take over at the first
        // MCU start inside their region (the check points of slot s), so no slot decodes much more than one region
        int owner = s, prev_cnt = before;
        unsigned prev_bit = X;
        const int u_end = last ? nsub : sy.t;
        for (int u = s + 1; u < u_end && u - s - 1 < JS_CK; ++u) {
            const JsCk ck = cks[(size_t)(first + s) * JS_CK + (u - s - 1)];
            if (ck.bit == 0xFFFFFFFFu) continue;                           // no MCU start of the true parse inside region u: it stays with the owner
            const int at = ck.count - before;                               // MCUs of this link before the check point
            if (at < prev_cnt - before || at > n_link) { ok = false; break; }
            starts[first + owner] = JsStart{prev_bit, m + (prev_cnt - before), ck.count - prev_cnt, 0};
            owner = u; prev_bit = ck.bit; prev_cnt = ck.count;
        }
        if (!ok) break;
        starts[first + owner] = JsStart{prev_bit, m + (prev_cnt - before), n_link - (prev_cnt - before), 0};
        if (last) break;
        m += n_link; s = sy.t; X = sy.X;
    }
    if (!ok) {   // the chain broke (or the data is damaged): the first slot decodes the whole interval
        for (int u = 1; u < nsub; ++u) starts[first + u] = JsStart{0u, 0, 0, 0};
        starts[first] = JsStart{0u, 0, total, 0};
    }
}

// The real decode of one slot's MCU range (the flat symbol loop of jpeg_huff_kernel), DC predictors starting at 0; the predictors
// at the end of the range go to preds[slot] for jpeg_dcfix_kernel.
__global__ void __launch_bounds__(128) jpeg_huff_sub_kernel(const JpegDev* __restrict__ files, const int* __restrict__ block_file, const int* __restrict__ block_first,
                                                            const JpegTables* __restrict__ tables, const unsigned* __restrict__ seg,
                                                            const unsigned char* __restrict__ clean, const unsigned* __restrict__ clean_len,
                                                            const JsSlot* __restrict__ slots, const JsStart* __restrict__ starts, int4* __restrict__ preds,
                                                            short* __restrict__ coef, const unsigned* __restrict__ wend) {
    extern __shared__ __align__(16) unsigned char jh_smem[];
    HuffDev* s_tab = reinterpret_cast<HuffDev*>(jh_smem);
    __shared__ unsigned char s_zz[64];
    __shared__ int s_slot[6];
    const int fi = block_file[blockIdx.x];
    const JpegDev& f = files[fi];
    jh_build_tables(f, tables[fi], s_tab, s_zz, s_slot);
    const int slot = block_first[blockIdx.x] + threadIdx.x;
    if (slot >= block_first[blockIdx.x + 1]) return;
    const JsSlot sl = slots[slot];
    if (sl.j < 0) return;
    const JsStart st = starts[slot];
    if (st.n_mcu <= 0) { preds[slot] = make_int4(0, 0, 0, 0); return; }
    const unsigned clen = clean_len[fi];
    const unsigned a0 = min(seg[f.seg_base + sl.j], clen);
    const unsigned char* base = clean + f.clean_off;
    const unsigned long long abit = 8ull * a0 + st.bit;
    const unsigned* wb = reinterpret_cast<const unsigned*>(base);
    const unsigned wmax = (unsigned)(wend - wb);
    unsigned wi = (unsigned)(abit >> 5);
    unsigned w0 = __byte_perm(__ldg(wb + wi), 0, 0x0123), w1 = __byte_perm(__ldg(wb + wi + 1), 0, 0x0123);
    unsigned nxt = __ldg(wb + wi + 2);
    wi += 3;
    unsigned o = (unsigned)(abit & 31ull);
    long long m = (f.ri > 0 ? (long long)sl.j * f.ri : 0) + st.mcu;
    const long long m1 = m + st.n_mcu;
    int my = (int)(m / f.mcux), mx = (int)(m - (long long)my * f.mcux);
    const int nb0 = f.c[0].h * f.c[0].v, nb = f.nc == 1 ? 1 : nb0 + 2;
    const int h0 = f.c[0].h, v0 = f.c[0].v;
    const int c1 = f.nc > 1 ? 1 : 0, c2 = f.nc > 2 ? 2 : 0;
    const HuffDev* tdc0 = &s_tab[s_slot[0]];
    const HuffDev* tac0 = &s_tab[s_slot[1]];
    const HuffDev* tdc1 = &s_tab[s_slot[2 * c1]];
    const HuffDev* tac1 = &s_tab[s_slot[2 * c1 + 1]];
    const HuffDev* tdc2 = &s_tab[s_slot[2 * c2]];
    const HuffDev* tac2 = &s_tab[s_slot[2 * c2 + 1]];
    const unsigned cbase0 = f.c[0].coef_base, cbase1 = f.c[c1].coef_base, cbase2 = f.c[c2].coef_base;
    const int cwb0 = f.c[0].wb, cwb1 = f.c[c1].wb, cwb2 = f.c[c2].wb;
    int pred0 = 0, pred1 = 0, pred2 = 0;
    int b = 0, k = 0, ci = 0;
    short* blk = coef + ((size_t)cbase0 + (size_t)(my * v0) * cwb0 + mx * h0) * 64;
    const HuffDev* tab = tdc0;
    while (m < m1) {
        const unsigned win = __funnelshift_l(w1, w0, o);
        unsigned e = tab->lut[win >> (32 - JH_LUT_BITS)];
        if ((e & 31u) == 0) {
            const unsigned slot2 = e >> 16;
            if (slot2 < JH_SUB_SLOTS) e = tab->sub[slot2 * 64 + ((win >> 16) & 63u)];
            else {
                const unsigned top = win >> 16;
                unsigned sym = 0, l = 16;
#pragma unroll 1
                for (int q = JH_LUT_BITS + 1; q <= 16; ++q) {
                    const int code = (int)(top >> (16 - q));
                    if (code <= tab->maxcode[q]) { sym = tab->vals[(tab->valoff[q] + code) & 255]; l = q; break; }
                }
                e = k == 0 ? (l | ((sym & 15u) << 9)) : (l | ((sym >> 4) << 5) | ((sym & 15u) << 9));
            }
        }
        unsigned L = e & 31u;
        const int r = (int)((e >> 5) & 15u), sz = (int)((e >> 9) & 15u);
        int v = (int)e >> 16;
        if (!(e & (1u << 13)) && sz) {
            v = jh_extend((int)((win << L) >> (32 - sz)), sz);
            L += (unsigned)sz;
        }
        o += L;
        if (o >= 32u) {
            o -= 32u;
            w0 = w1;
            w1 = __byte_perm(nxt, 0, 0x0123);
            nxt = __ldg(wb + wi);
            wi = min(wi + 1u, wmax);
        }
        if (k == 0) {
            int pv;
            if (ci == 0) { pred0 += v; pv = pred0; } else if (ci == 1) { pred1 += v; pv = pred1; } else { pred2 += v; pv = pred2; }
            if (pv) blk[0] = (short)pv;
            k = 1;
            tab = ci == 0 ? tac0 : (ci == 1 ? tac1 : tac2);
        } else if (sz) {
            k += r;
            if (k < 64) blk[s_zz[k]] = (short)v;
            ++k;
        } else {
            k = r == 15 ? k + 16 : 64;
        }
        if (k >= 64) {
            k = 0;
            if (++b == nb) { b = 0; ++m; if (++mx == f.mcux) { mx = 0; ++my; } }
            ci = max(0, b - nb0 + 1);
            const int bi = ci == 0 ? b : 0;
            const int dy = h0 == 2 ? (bi >> 1) : bi, dx = h0 == 2 ? (bi & 1) : 0;
            const int hh = ci == 0 ? h0 : 1, vv = ci == 0 ? v0 : 1;
            const unsigned cb_ = ci == 0 ? cbase0 : (ci == 1 ? cbase1 : cbase2);
            const int wb_ = ci == 0 ? cwb0 : (ci == 1 ? cwb1 : cwb2);
            blk = coef + ((size_t)cb_ + (size_t)(my * vv + dy) * wb_ + (mx * hh + dx)) * 64;
            tab = ci == 0 ? tdc0 : (ci == 1 ? tdc1 : tdc2);
        }
    }
    preds[slot] = make_int4(pred0, pred1, pred2, 0);
}

// DC terms are coded as differences: a slot that did not start its interval decoded its DC terms from predictors 0, so the
// predictors at the end of all earlier slots of the interval are added to the DC term of every block of its MCU range.
#define JS_FIX_LANES 8   // threads that share the MCU range of one slot
__global__ void __launch_bounds__(128) jpeg_dcfix_kernel(const JpegDev* __restrict__ files, const int* __restrict__ block_file, const int* __restrict__ block_first,
                                                         const JsSlot* __restrict__ slots, const JsStart* __restrict__ starts, const int4* __restrict__ preds,
                                                         short* __restrict__ coef) {
    // grid: JS_FIX_LANES blocks per block of 128 slots; thread t of block q works for slot (q % JS_FIX_LANES) * 16 + t / 8 of that group
    const int sb = blockIdx.x / JS_FIX_LANES, part = blockIdx.x % JS_FIX_LANES;
    const int fi = block_file[sb];
    const JpegDev& f = files[fi];
    const int slot = block_first[sb] + part * (128 / JS_FIX_LANES) + (threadIdx.x / JS_FIX_LANES);
    const int lane8 = threadIdx.x % JS_FIX_LANES;
    if (slot >= block_first[sb + 1]) return;
    const JsSlot sl = slots[slot];
    if (sl.j < 0 || sl.k == 0) return;
    const JsStart st = starts[slot];
    if (st.n_mcu <= 0) return;
    int c0 = 0, c1 = 0, c2 = 0;
    for (int u = 1; u <= sl.k; ++u) { const int4 q = preds[slot - u]; c0 += q.x; c1 += q.y; c2 += q.z; }
    if ((c0 | c1 | c2) == 0) return;
    const int nb0 = f.c[0].h * f.c[0].v, h0 = f.c[0].h, v0 = f.c[0].v;
    const long long mb = (f.ri > 0 ? (long long)sl.j * f.ri : 0) + st.mcu;
    for (long long m = mb + lane8; m < mb + st.n_mcu; m += JS_FIX_LANES) {
        const int my = (int)(m / f.mcux), mx = (int)(m - (long long)my * f.mcux);
        if (c0) {
            for (int bi = 0; bi < nb0; ++bi) {
                const int dy = h0 == 2 ? (bi >> 1) : bi, dx = h0 == 2 ? (bi & 1) : 0;
                short* bp = coef + ((size_t)f.c[0].coef_base + (size_t)(my * v0 + dy) * f.c[0].wb + (mx * h0 + dx)) * 64;
                bp[0] = (short)(bp[0] + c0);
            }
        }
        if (f.nc > 1) {
            if (c1) { short* bp = coef + ((size_t)f.c[1].coef_base + (size_t)my * f.c[1].wb + mx) * 64; bp[0] = (short)(bp[0] + c1); }
            if (c2) { short* bp = coef + ((size_t)f.c[2].coef_base + (size_t)my * f.c[2].wb + mx) * 64; bp[0] = (short)(bp[0] + c2); }
        }
    }
}

// ---- K-J3: dequantise + jidctint.c jpeg_idct_islow ------------------------------------------------------------------------------
#define JI_CONST_BITS 13
#define JI_PASS1_BITS 2
__device__ __forceinline__ int ji_descale(int x, int n) { return (x + (1 << (n - 1))) >> n; }
__device__ __forceinline__ void ji_idct_1d(const int in[8], int out[8], int shift) {
    int z2 = in[2], z3 = in[6];
    int z1 = (z2 + z3) * 4433;
    int tmp2 = z1 + z3 * (-15137);
    int tmp3 = z1 + z2 * 6270;
    z2 = in[0]; z3 = in[4];
    int tmp0 = (z2 + z3) * (1 << JI_CONST_BITS);
    int tmp1 = (z2 - z3) * (1 << JI_CONST_BITS);
    const int tmp10 = tmp0 + tmp3, tmp13 = tmp0 - tmp3, tmp11 = tmp1 + tmp2, tmp12 = tmp1 - tmp2;
    tmp0 = in[7]; tmp1 = in[5]; tmp2 = in[3]; tmp3 = in[1];
    z1 = tmp0 + tmp3; z2 = tmp1 + tmp2; z3 = tmp0 + tmp2;
    int z4 = tmp1 + tmp3;
    const int z5 = (z3 + z4) * 9633;
    tmp0 *= 2446; tmp1 *= 16819; tmp2 *= 25172; tmp3 *= 12299;
    z1 *= -7373; z2 *= -20995; z3 *= -16069; z4 *= -3196;
    z3 += z5; z4 += z5;
    tmp0 += z1 + z3; tmp1 += z2 + z4; tmp2 += z2 + z3; tmp3 += z1 + z4;
    out[0] = ji_descale(tmp10 + tmp3, shift); out[7] = ji_descale(tmp10 - tmp3, shift);
    out[1] = ji_descale(tmp11 + tmp2, shift); out[6] = ji_descale(tmp11 - tmp2, shift);
    out[2] = ji_descale(tmp12 + tmp1, shift); out[5] = ji_descale(tmp12 - tmp1, shift);
    out[3] = ji_descale(tmp13 + tmp0, shift); out[4] = ji_descale(tmp13 - tmp0, shift);
}
__device__ __forceinline__ unsigned ji_range_limit(int x) {   // range_limit[x & RANGE_MASK] of libjpeg's post-IDCT table (jdmaster.c)
    const int i = x & 1023;
    return i < 128 ? 128u + i : (i < 512 ? 255u : (i < 896 ? 0u : (unsigned)(i - 896)));
}
// 256 threads = 32 blocks of 8x8; thread (g, t): row t of the coefficient load, column t of pass 1, row t of pass 2
__global__ void __launch_bounds__(256) jpeg_idct_kernel(const JpegDev* __restrict__ files, const JpegTables* __restrict__ tables,
                                                        const unsigned* __restrict__ grp_prefix, int n_files,
                                                        const short* __restrict__ coef, unsigned char* __restrict__ planes) {
    // `files` / `tables` point at the first file of the unit; the grid is flat over the files' groups of 32 blocks (grp_prefix), so a
    // batch of mixed page sizes launches no idle thread blocks; one binary search per thread block
    __shared__ short s_coef[32][72];
    __shared__ int s_ws[32][72];
    __shared__ int s_fi;
    if (grp_prefix) {
        if (threadIdx.x == 0) {
            int lo = 0, hi = n_files;
            while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (grp_prefix[mid] <= blockIdx.x) lo = mid; else hi = mid; }
            s_fi = lo;
        }
        __syncthreads();
    }
    const int g = threadIdx.x >> 3, t = threadIdx.x & 7;
    const int fi = grp_prefix ? s_fi : (int)blockIdx.y;               // uniform batches: grid (groups, files), no search
    const JpegDev& f = files[fi];
    const unsigned lb = (grp_prefix ? blockIdx.x - grp_prefix[fi] : blockIdx.x) * 32u + g;        // block index inside the file
    const bool live = lb < f.n_blocks;
    const unsigned B = f.block_base + lb;
    int ci = 0;
    unsigned local = 0;
    bool ac = false;       // this thread's coefficient row holds something besides the block's DC term
    int dc = 0;
    if (live) {
        while (ci + 1 < f.nc && B >= f.c[ci + 1].coef_base) ++ci;
        local = B - f.c[ci].coef_base;
        const uint4 cv = __ldg(reinterpret_cast<const uint4*>(coef + (size_t)B * 64 + t * 8));
        *reinterpret_cast<uint4*>(&s_coef[g][t * 8]) = cv;
        ac = ((t == 0 ? (cv.x & 0xffff0000u) : cv.x) | cv.y | cv.z | cv.w) != 0u;
        dc = (int)(short)(cv.x & 0xffffu);
    }
    // DC-only blocks (the blank background of a page: most blocks of a text page): both passes collapse to one value — exactly what
    // jidctint.c's zero-AC shortcuts compute, and what the full passes below would (DESCALE(dc << 13, 11) == dc << 2 without a rounding
    // carry) — so the eight threads of such a block skip the passes and the shared-memory transposes
    const unsigned grp = (__ballot_sync(0xffffffffu, ac) >> (threadIdx.x & 24)) & 0xffu;
    const bool full = live && grp != 0u;
    dc = __shfl_sync(0xffffffffu, dc, threadIdx.x & 24);
    __syncwarp();
    if (full) {
        const unsigned short* q = tables[fi].qt[f.c[ci].tq];
        int in[8], o[8];
#pragma unroll
        for (int r = 0; r < 8; ++r) in[r] = (int)s_coef[g][r * 8 + t] * (int)__ldg(q + r * 8 + t);
        ji_idct_1d(in, o, JI_CONST_BITS - JI_PASS1_BITS);
#pragma unroll
        for (int r = 0; r < 8; ++r) s_ws[g][r * 8 + t] = o[r];
    }
    __syncwarp();
    if (live) {
        const JpegComp& c = f.c[ci];
        unsigned lo4, hi4;
        if (full) {
            int in[8], o[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) in[k] = s_ws[g][t * 8 + k];
            ji_idct_1d(in, o, JI_CONST_BITS + JI_PASS1_BITS + 3);
            lo4 = ji_range_limit(o[0]) | (ji_range_limit(o[1]) << 8) | (ji_range_limit(o[2]) << 16) | (ji_range_limit(o[3]) << 24);
            hi4 = ji_range_limit(o[4]) | (ji_range_limit(o[5]) << 8) | (ji_range_limit(o[6]) << 16) | (ji_range_limit(o[7]) << 24);
        } else {
            const int w0 = (dc * (int)__ldg(tables[fi].qt[c.tq])) << JI_PASS1_BITS;                        // pass 1, column 0, every row
            const unsigned px = ji_range_limit(ji_descale(w0 << JI_CONST_BITS, JI_CONST_BITS + JI_PASS1_BITS + 3));   // pass 2, every column
            lo4 = hi4 = px * 0x01010101u;
        }
        const unsigned by = local / (unsigned)c.wb, bx = local - by * (unsigned)c.wb;
        unsigned char* dst = planes + c.plane_off + ((size_t)by * 8 + t) * ((size_t)c.wb * 8) + (size_t)bx * 8;
        *reinterpret_cast<uint2*>(dst) = make_uint2(lo4, hi4);
    }
}

// ---- K-J4: jdsample.c fancy up-sampling + jdcolor.c YCbCr -> RGB -------------------------------------------------------------------
// One block = 1024 pixels of one output row of one file (grid: x chunks, rows, files; blocks outside a file's extent exit), one
// thread = 8 consecutive pixels: Y as one aligned 8-byte load, chroma as aligned words plus the two neighbour columns the triangle
// filter needs (libjpeg's edge rules are exactly "clamp the neighbour index": (3c + c + 8) >> 4 == (4c + 8) >> 4).  The 24 output
// bytes per thread are staged in shared memory and leave as coalesced 32-bit stores whatever the alignment of the row.
#define JC_PX 8
#define JC_THREADS 256   // upper bound; the launch uses the smallest multiple of 32 threads that covers the widest row (<= 2048 px per block)
__device__ __forceinline__ void jc_chroma8(const unsigned char* __restrict__ pl, int stride, int w, int h, int hs, int vs, int x0, int y, int out[8]) {
    if (hs == 1) {
        const int r = vs == 2 ? (y >> 1) : y;
        const unsigned char* a = pl + (size_t)r * stride + x0;
        const uint2 va = __ldg(reinterpret_cast<const uint2*>(a));
        if (vs == 1) {
#pragma unroll
            for (int k = 0; k < 8; ++k) out[k] = (int)(((k < 4 ? va.x : va.y) >> (8 * (k & 3))) & 0xFFu);
            return;
        }
        const int lower = y & 1;
        const int r1 = lower ? min(r + 1, h - 1) : max(r - 1, 0);
        const uint2 vb = __ldg(reinterpret_cast<const uint2*>(pl + (size_t)r1 * stride + x0));
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int n0 = (int)(((k < 4 ? va.x : va.y) >> (8 * (k & 3))) & 0xFFu), n1 = (int)(((k < 4 ? vb.x : vb.y) >> (8 * (k & 3))) & 0xFFu);
            out[k] = (n0 * 3 + n1 + (lower ? 2 : 1)) >> 2;
        }
        return;
    }
    // hs == 2: chroma columns i0 .. i0+3 cover the 8 pixels; neighbours i0-1 and i0+4 clamped into [0, w-1]
    const int i0 = x0 >> 1;
    const int r = vs == 2 ? (y >> 1) : y;
    const unsigned char* a = pl + (size_t)r * stride;
    const unsigned wa = __ldg(reinterpret_cast<const unsigned*>(a + i0));
    const int il = max(i0 - 1, 0), ir = min(i0 + 4, w - 1);
    int c[6];
    if (vs == 1) {
        c[0] = __ldg(a + il); c[5] = __ldg(a + ir);
#pragma unroll
        for (int k = 0; k < 4; ++k) c[1 + k] = (int)((wa >> (8 * k)) & 0xFFu);
        if (w <= 2) {   // h2v1_upsample: replication
#pragma unroll
            for (int k = 0; k < 8; ++k) out[k] = c[1 + (k >> 1)];
            return;
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int i = i0 + k;
            const int prev = i == 0 ? c[1 + k] : c[k], next = i >= w - 1 ? c[1 + k] : c[2 + k];
            out[2 * k] = (c[1 + k] * 3 + prev + 1) >> 2;
            out[2 * k + 1] = (c[1 + k] * 3 + next + 2) >> 2;
        }
        return;
    }
    const int lower = y & 1;
    const int r1 = lower ? min(r + 1, h - 1) : max(r - 1, 0);
    const unsigned char* b = pl + (size_t)r1 * stride;
    const unsigned wb = __ldg(reinterpret_cast<const unsigned*>(b + i0));
    if (w <= 2) {       // h2v2_upsample: replication
#pragma unroll
        for (int k = 0; k < 8; ++k) out[k] = (int)((wa >> (8 * (k >> 1))) & 0xFFu);
        return;
    }
    c[0] = __ldg(a + il) * 3 + __ldg(b + il);
    c[5] = __ldg(a + ir) * 3 + __ldg(b + ir);
#pragma unroll
    for (int k = 0; k < 4; ++k) c[1 + k] = (int)((wa >> (8 * k)) & 0xFFu) * 3 + (int)((wb >> (8 * k)) & 0xFFu);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int i = i0 + k;
        const int prev = i == 0 ? c[1 + k] : c[k], next = i >= w - 1 ? c[1 + k] : c[2 + k];
        out[2 * k] = (c[1 + k] * 3 + prev + 8) >> 4;
        out[2 * k + 1] = (c[1 + k] * 3 + next + 7) >> 4;
    }
}
// FIXED = 1: every file of the launch is 3-component 4:2:0 (the sampling factors fold into the code; checked by the host), 0: per file
template <int FIXED>
__global__ void __launch_bounds__(JC_THREADS) jpeg_color_kernel(const JpegDev* __restrict__ files, const unsigned char* __restrict__ planes,
                                                                uint8_t* const* __restrict__ outs, const unsigned* __restrict__ unit_prefix, int n_files) {
    __shared__ unsigned s_rgb[JC_THREADS * 6 + 1];
    __shared__ int s_fi;
    int fi, y, xb;
    const int chunk_px = (int)blockDim.x * JC_PX;
    if (unit_prefix) {   // mixed page sizes: flat grid over (file, row, chunk of blockDim.x * 8 pixels), one binary search per thread block
        if (threadIdx.x == 0) {
            int lo = 0, hi = n_files;
            while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (unit_prefix[mid] <= blockIdx.x) lo = mid; else hi = mid; }
            s_fi = lo;
        }
        __syncthreads();
        fi = s_fi;
        const int xchunks = (files[fi].X + chunk_px - 1) / chunk_px;
        const int u = (int)(blockIdx.x - unit_prefix[fi]);
        y = u / xchunks; xb = (u - y * xchunks) * chunk_px;
    } else { fi = blockIdx.z; y = blockIdx.y; xb = blockIdx.x * chunk_px; }   // uniform batch: grid (x chunks, rows, files)
    const JpegDev& f = files[fi];
    if (y >= f.Y || xb >= f.X) return;
    const int x0 = xb + threadIdx.x * JC_PX;
    // the sample planes are padded to whole blocks (stride wb*8 >= X rounded up to 8), so a thread with x0 < X may read its 8 samples
    if (x0 < f.X) {
        const uint2 yv = __ldg(reinterpret_cast<const uint2*>(planes + f.c[0].plane_off + (size_t)y * ((size_t)f.c[0].wb * 8) + x0));
        unsigned char o[24];
        if (!FIXED && f.nc == 1) {
#pragma unroll
            for (int k = 0; k < 8; ++k) { const unsigned v = ((k < 4 ? yv.x : yv.y) >> (8 * (k & 3))) & 0xFFu; o[3 * k] = o[3 * k + 1] = o[3 * k + 2] = (unsigned char)v; }
        } else {
            const int hs = FIXED ? 2 : f.max_h / f.c[1].h, vs = FIXED ? 2 : f.max_v / f.c[1].v;
            int cb[8], cr[8];
            jc_chroma8(planes + f.c[1].plane_off, f.c[1].wb * 8, f.c[1].ds_w, f.c[1].ds_h, hs, vs, x0, y, cb);
            jc_chroma8(planes + f.c[2].plane_off, f.c[2].wb * 8, f.c[2].ds_w, f.c[2].ds_h, hs, vs, x0, y, cr);
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const int yy = (int)(((k < 4 ? yv.x : yv.y) >> (8 * (k & 3))) & 0xFFu);
                const int u = cb[k] - 128, v = cr[k] - 128;
                const int r = yy + ((91881 * v + 32768) >> 16);
                const int g = yy + ((-22554 * u + 32768 - 46802 * v) >> 16);
                const int b = yy + ((116130 * u + 32768) >> 16);
                o[3 * k] = (unsigned char)min(max(r, 0), 255); o[3 * k + 1] = (unsigned char)min(max(g, 0), 255); o[3 * k + 2] = (unsigned char)min(max(b, 0), 255);
            }
        }
#pragma unroll
        for (int k = 0; k < 6; ++k) s_rgb[threadIdx.x * 6 + k] = o[4 * k] | (o[4 * k + 1] << 8) | (o[4 * k + 2] << 16) | ((unsigned)o[4 * k + 3] << 24);
    }
    __syncthreads();
    const int npx = min((int)blockDim.x * JC_PX, f.X - xb);
    const int nbytes = 3 * npx;
    unsigned char* gp = outs[fi] + ((size_t)y * f.X + xb) * 3;
    const int head = min((int)((4u - ((unsigned)(uintptr_t)gp & 3u)) & 3u), nbytes);
    const unsigned char* sb = reinterpret_cast<const unsigned char*>(s_rgb);
    if ((int)threadIdx.x < head) gp[threadIdx.x] = sb[threadIdx.x];
    const int nw = (nbytes - head) >> 2;
    unsigned* gw = reinterpret_cast<unsigned*>(gp + head);
    const unsigned sh = 8u * (unsigned)(head & 3);
    for (int wi = threadIdx.x; wi < nw; wi += blockDim.x) {
        const int so = (head + 4 * wi) >> 2;
        gw[wi] = __funnelshift_r(s_rgb[so], s_rgb[so + 1], sh);
    }
    const int tail0 = head + 4 * nw;
    if ((int)threadIdx.x < nbytes - tail0) gp[tail0 + threadIdx.x] = sb[tail0 + threadIdx.x];
}

// K-J4 for 3-component 4:2:0 files of one size (the common case: every page of a scanner / camera batch): one thread = 8 pixels of
// TWO output rows.  Rows 2r+1 and 2r+2 use the same two chroma rows (r and r+1: 3a + b for the upper of the two, a + 3b for the lower), so
// the chroma words are loaded and unpacked once for 16 pixels; everything is the same arithmetic as jc_chroma8 + the generic kernel
// (jdsample.c h2v2_fancy_upsample, jdcolor.c), written for one sampling so the compiler sees constants (ncu: the generic kernel issues
// 83 instructions per pixel at 77 % issue utilisation — it is instruction-bound, not memory-bound).
// grid (x chunks, Y / 2 + 1 row pairs, files); row pair p = output rows 2p - 1 and 2p.  The host picks it only for chroma planes wider
// than two samples (jdsample.c falls back to replication below that; the generic kernel has that branch).
__global__ void __launch_bounds__(JC_THREADS) jpeg_color420_kernel(const JpegDev* __restrict__ files, const unsigned char* __restrict__ planes,
                                                                   uint8_t* const* __restrict__ outs, const unsigned* __restrict__ unit_prefix, int n_files) {
    __shared__ unsigned s_rgb[2][JC_THREADS * 6 + 1];
    __shared__ int s_fi;
    int fi, p, xb;
    const int chunk_px = (int)blockDim.x * JC_PX;
    if (unit_prefix) {   // pages of different sizes: flat grid over (file, row pair, chunk), one binary search per thread block
        if (threadIdx.x == 0) {
            int lo = 0, hi = n_files;
            while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (unit_prefix[mid] <= blockIdx.x) lo = mid; else hi = mid; }
            s_fi = lo;
        }
        __syncthreads();
        fi = s_fi;
        const int xchunks = (files[fi].X + chunk_px - 1) / chunk_px;
        const int u = (int)(blockIdx.x - unit_prefix[fi]);
        p = u / xchunks; xb = (u - p * xchunks) * chunk_px;
    } else { fi = blockIdx.z; p = blockIdx.y; xb = blockIdx.x * chunk_px; }
    const JpegDev& f = files[fi];
    const int X = f.X, Y = f.Y;
    const int ya = 2 * p - 1, yb = 2 * p;
    if (xb >= X || ya >= Y) return;
    const int x0 = xb + threadIdx.x * JC_PX;
    const bool has_a = ya >= 0, has_b = yb < Y;
    if (x0 < X) {
        const int cw = f.c[1].ds_w, ch = f.c[1].ds_h, cstride = f.c[1].wb * 8;
        const int rA = min(max(p - 1, 0), ch - 1), rB = min(p, ch - 1);
        const int i0 = x0 >> 1, il = max(i0 - 1, 0), ir = min(i0 + 4, cw - 1);
        int sa[2][6], sb[2][6];   // [plane][column i0-1 .. i0+4]: 3a + b (row ya) and a + 3b (row yb)
#pragma unroll
        for (int pl = 0; pl < 2; ++pl) {
            const unsigned char* base = planes + f.c[1 + pl].plane_off;
            const unsigned char* a = base + (size_t)rA * cstride;
            const unsigned char* b = base + (size_t)rB * cstride;
            const unsigned wa = __ldg(reinterpret_cast<const unsigned*>(a + i0)), wb = __ldg(reinterpret_cast<const unsigned*>(b + i0));
            int ca[6], cb[6];
            ca[0] = __ldg(a + il); ca[5] = __ldg(a + ir); cb[0] = __ldg(b + il); cb[5] = __ldg(b + ir);
#pragma unroll
            for (int k = 0; k < 4; ++k) { ca[1 + k] = (int)__byte_perm(wa, 0, 0x4440 + k); cb[1 + k] = (int)__byte_perm(wb, 0, 0x4440 + k); }
#pragma unroll
            for (int k = 0; k < 6; ++k) { sa[pl][k] = ca[k] * 3 + cb[k]; sb[pl][k] = cb[k] * 3 + ca[k]; }
        }
        const size_t ystride = (size_t)f.c[0].wb * 8;
        const unsigned char* yp = planes + f.c[0].plane_off + x0;
#pragma unroll
        for (int row = 0; row < 2; ++row) {
            if (row == 0 ? !has_a : !has_b) continue;
            const uint2 yv = __ldg(reinterpret_cast<const uint2*>(yp + (size_t)(row == 0 ? ya : yb) * ystride));
            unsigned char o[24];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int i = i0 + k;
                int uv[2][2];
#pragma unroll
                for (int pl = 0; pl < 2; ++pl) {
                    const int* s6 = row == 0 ? sa[pl] : sb[pl];
                    const int c = s6[1 + k];
                    const int prev = i == 0 ? c : s6[k], next = i >= cw - 1 ? c : s6[2 + k];
                    uv[pl][0] = ((c * 3 + prev + 8) >> 4) - 128;
                    uv[pl][1] = ((c * 3 + next + 7) >> 4) - 128;
                }
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int kk = 2 * k + h;
                    const int yy = (int)__byte_perm(kk < 4 ? yv.x : yv.y, 0, 0x4440 + (kk & 3));
                    const int u = uv[0][h], v = uv[1][h];
                    const int r = yy + ((91881 * v + 32768) >> 16);
                    const int g = yy + ((-22554 * u + 32768 - 46802 * v) >> 16);
                    const int bl = yy + ((116130 * u + 32768) >> 16);
                    o[3 * kk] = (unsigned char)min(max(r, 0), 255); o[3 * kk + 1] = (unsigned char)min(max(g, 0), 255); o[3 * kk + 2] = (unsigned char)min(max(bl, 0), 255);
                }
            }
#pragma unroll
            for (int k = 0; k < 6; ++k) s_rgb[row][threadIdx.x * 6 + k] = o[4 * k] | (o[4 * k + 1] << 8) | (o[4 * k + 2] << 16) | ((unsigned)o[4 * k + 3] << 24);
        }
    }
    __syncthreads();
    const int npx = min((int)blockDim.x * JC_PX, X - xb);
    const int nbytes = 3 * npx;
#pragma unroll
    for (int row = 0; row < 2; ++row) {
        if (row == 0 ? !has_a : !has_b) continue;
        unsigned char* gp = outs[fi] + ((size_t)(row == 0 ? ya : yb) * X + xb) * 3;
        const int head = min((int)((4u - ((unsigned)(uintptr_t)gp & 3u)) & 3u), nbytes);
        const unsigned char* sbp = reinterpret_cast<const unsigned char*>(s_rgb[row]);
        if ((int)threadIdx.x < head) gp[threadIdx.x] = sbp[threadIdx.x];
        const int nw = (nbytes - head) >> 2;
        unsigned* gw = reinterpret_cast<unsigned*>(gp + head);
        const unsigned sh = 8u * (unsigned)(head & 3);
        for (int wi = threadIdx.x; wi < nw; wi += blockDim.x) {
            const int so = (head + 4 * wi) >> 2;
            gw[wi] = __funnelshift_r(s_rgb[row][so], s_rgb[row][so + 1], sh);
        }
        const int tail0 = head + 4 * nw;
        if ((int)threadIdx.x < nbytes - tail0) gp[tail0 + threadIdx.x] = sbp[tail0 + threadIdx.x];
    }
}

// ---- host: two phases ---------------------------------------------------------------------------------------------------------------
// Phase 1 (entropy): everything that is a serial chain per restart interval — K-J1 + K-J2 for ALL files of a call in one launch each,
// on the stream the caller names (run_pages: the copy stream, right behind the upload of the files).  The Huffman kernel's
// duration is the longest interval's chain whatever the number of files, so it is paid once per call, not once per unit.
// Phase 2 (pixels): K-J3 + K-J4 for a range of files, on the stream of the unit that consumes the pages.
// d_bytes[i]: device pointer to file i (the whole file; the entropy-coded data starts at infos[i].ecs_off).
retto_b200_status rt_jpeg_entropy_enqueue(retto_b200_ctx* ctx, cudaStream_t st, const JpegInfo* infos, const uint8_t* const* d_bytes, int n) {
    retto_b200_ctx::JpegBatch& JB = ctx->jpeg;
    JB.n = 0;
    if (n <= 0) return RETTO_B200_OK;
    const size_t desc_bytes = (sizeof(JpegDev) * (size_t)n + 15) & ~size_t(15);
    long long total_seg = 0;
    size_t n_tblocks = 0;
    // Per file: long restart intervals (above all a file without restart markers = ONE interval) are cut into sub-sequences and decoded
    // through the self-synchronising parse (K-J2s); files with shorter intervals are parallel enough for K-J2.
    // Measured (256 pages 1280x1280, q90, 4:2:0): no restart markers: K-J2 42.6 ms -> K-J2s 6.0 ms (parse 1.4 + look-ahead 2.2 + resolve 0.3 +
    // decode 1.7 + DC fix-up 0.3); one interval per MCU row (2.2 KB on average): tiered K-J2 2.84 ms, K-J2s 3.65 ms.
    static const bool no_sub = getenv("RETTO_B200_JPEG_NOSUB") != nullptr;       // A/B: never
    static const bool force_sub = getenv("RETTO_B200_JPEG_SUB") != nullptr;      // A/B / tests: every file
    std::vector<unsigned char> file_sub(n, 0);
    long long total_seg_all = 0;
    int n_sub_files = 0;
    for (int i = 0; i < n; ++i) {
        total_seg_all += infos[i].n_seg;
        file_sub[i] = !no_sub && infos[i].n_seg > 0 && (force_sub || infos[i].ecs_len / (unsigned long long)infos[i].n_seg > JS_MIN_MEAN_INTERVAL);
        if (file_sub[i]) ++n_sub_files; else total_seg += infos[i].n_seg;
    }
    // jpeg_huff_kernel layout: tiered (16 warps share 112 intervals, the 8 longest alone in their warp) while every block of the batch
    // is resident at once (4 blocks of 512 threads per SM); dense (128 intervals per 128 threads) for batches with more intervals
    // than that — short restart intervals, where throughput, not the longest chain, is the bound.  RETTO_B200_JPEG_DENSE=1: A/B.
    static const bool force_dense = getenv("RETTO_B200_JPEG_DENSE") != nullptr;
    const bool tiered = !force_dense && total_seg <= 148LL * 4 * JH_T_IPB;
    const int jh_ipb = tiered ? JH_T_IPB : JH_THREADS;   // restart intervals per block
    for (int i = 0; i < n; ++i) if (!file_sub[i]) n_tblocks += ((size_t)infos[i].n_seg + jh_ipb - 1) / jh_ipb;
    if (total_seg_all > 0x3fffffffLL) { ctx->set_error("jpeg decode: too many restart intervals"); return RETTO_B200_ERR_CAPACITY; }
    const size_t tb_bytes = (sizeof(int) * 2 * n_tblocks + 15) & ~size_t(15);
    const size_t head_bytes = desc_bytes + tb_bytes;
    const size_t tab_bytes = sizeof(JpegTables) * (size_t)n;
    // the descriptor blob is built in pinned memory owned by the batch (not a recycled staging slot: the copy runs on `st`)
    RT_CUDA_OK(ctx, JB.h_desc.ensure(head_bytes + tab_bytes));
    char* hp = JB.h_desc.as<char>();
    JpegDev* hd = reinterpret_cast<JpegDev*>(hp);
    int* h_tb_file = reinterpret_cast<int*>(hp + desc_bytes);
    int* h_tb_first = h_tb_file + n_tblocks;
    JpegTables* h_tab = reinterpret_cast<JpegTables*>(hp + head_bytes);
    unsigned long long blocks = 0, plane_bytes = 0, clean_bytes = 0;
    int seg_base = 0;
    size_t tb = 0;
    JB.X.resize(n); JB.Y.resize(n); JB.n_blocks.resize(n); JB.is420.resize(n);
    for (int i = 0; i < n; ++i) {
        const JpegInfo& J = infos[i];
        JpegDev& D = hd[i];
        memset(&D, 0, sizeof(D));
        if (J.status != RETTO_B200_OK || J.ecs_len > 0xfffffff0u) { ctx->set_error("jpeg decode: file " + std::to_string(i) + " was not parsed"); return RETTO_B200_ERR_INVALID_ARG; }
        D.ecs = d_bytes[i] + J.ecs_off; D.ecs_len = (unsigned)J.ecs_len;
        D.X = J.X; D.Y = J.Y; D.nc = J.nc; D.max_h = J.max_h; D.max_v = J.max_v; D.mcux = J.mcux; D.mcuy = J.mcuy; D.ri = J.ri; D.n_seg = J.n_seg;
        D.seg_base = seg_base; D.thread_base = seg_base; D.status_slot = i;
        D.block_base = (unsigned)blocks;
        D.clean_off = clean_bytes;
        clean_bytes += ((unsigned long long)J.ecs_len + 32 + 15) & ~15ULL;
        for (int c = 0; c < J.nc; ++c) {
            JpegComp& C = D.c[c];
            C.h = J.h[c]; C.v = J.v[c]; C.tq = J.tq[c]; C.td = J.td[c]; C.ta = J.ta[c];
            C.wb = J.mcux * J.h[c]; C.hb = J.mcuy * J.v[c];
            C.ds_w = (J.X * J.h[c] + J.max_h - 1) / J.max_h; C.ds_h = (J.Y * J.v[c] + J.max_v - 1) / J.max_v;
            C.coef_base = (unsigned)blocks;
            C.plane_off = plane_bytes;
            blocks += (unsigned long long)C.wb * C.hb;
            plane_bytes += ((unsigned long long)C.wb * C.hb * 64 + 15) & ~15ULL;
        }
        if (blocks > 0x7fffffffULL) { ctx->set_error("jpeg decode: batch too large (coefficient blocks)"); return RETTO_B200_ERR_CAPACITY; }
        D.n_blocks = (unsigned)blocks - D.block_base;
        JB.X[i] = J.X; JB.Y[i] = J.Y; JB.n_blocks[i] = D.n_blocks;
        JB.is420[i] = (J.nc == 3 && J.h[0] == 2 && J.v[0] == 2 && J.h[1] == 1 && J.v[1] == 1 && J.h[2] == 1 && J.v[2] == 1) ? 1 : 0;
        if (!file_sub[i]) for (int j0 = 0; j0 < J.n_seg; j0 += jh_ipb) { h_tb_file[tb] = i; h_tb_first[tb] = j0; ++tb; }
        seg_base += J.n_seg;
        for (int k = 0; k < 4; ++k) {
            memcpy(h_tab[i].qt[k], J.qt[k], sizeof(J.qt[k]));
            memcpy(h_tab[i].dc[k].bits, J.dc[k].bits, 17); memcpy(h_tab[i].dc[k].vals, J.dc[k].vals, 256);
            memcpy(h_tab[i].ac[k].bits, J.ac[k].bits, 17); memcpy(h_tab[i].ac[k].vals, J.ac[k].vals, 256);
        }
    }
    // grow-only buffers (growth synchronises ctx->stream; `st` holds no work on them: run_pages drains the copy stream before it returns)
    RT_CUDA_OK(ctx, ctx->d_jpeg_desc.ensure(head_bytes + tab_bytes, ctx->stream));
    RT_CUDA_OK(ctx, ctx->d_jpeg_seg.ensure(sizeof(unsigned) * ((size_t)std::max(seg_base, 1) + (size_t)n) + sizeof(int) * (size_t)n, ctx->stream));
    RT_CUDA_OK(ctx, ctx->d_jpeg_coef.ensure((size_t)blocks * 128, ctx->stream));
    RT_CUDA_OK(ctx, ctx->d_jpeg_planes.ensure((size_t)plane_bytes, ctx->stream));
    RT_CUDA_OK(ctx, ctx->d_jpeg_clean.ensure((size_t)clean_bytes + 64, ctx->stream));
    RT_CUDA_OK(ctx, cudaMemcpyAsync(ctx->d_jpeg_desc.p, hp, head_bytes + tab_bytes, cudaMemcpyHostToDevice, st));
    const char* dp = ctx->d_jpeg_desc.as<char>();
    const JpegDev* d_files = reinterpret_cast<const JpegDev*>(dp);
    const int* d_tb_file = reinterpret_cast<const int*>(dp + desc_bytes);
    const int* d_tb_first = d_tb_file + n_tblocks;
    const JpegTables* d_tab = reinterpret_cast<const JpegTables*>(dp + head_bytes);
    unsigned* d_seg = ctx->d_jpeg_seg.as<unsigned>();
    unsigned* d_clean_len = d_seg + std::max(seg_base, 1);
    int* d_status = reinterpret_cast<int*>(d_clean_len + n);
    ctx->jpeg_status_dev = d_status;
    JB.n = n; JB.desc_bytes = desc_bytes; JB.head_bytes = head_bytes;
    // last readable word of the clean arena (clean_bytes + 64 bytes were ensured): the bit readers clamp their look-ahead pointer to it
    const unsigned* d_wend = reinterpret_cast<const unsigned*>(ctx->d_jpeg_clean.as<unsigned char>()) + ((size_t)clean_bytes + 60) / 4;
    RT_CUDA_OK(ctx, cudaMemsetAsync(d_status, 0, sizeof(int) * (size_t)n, st));
    // the coefficient planes are zeroed (the Huffman kernels store non-zero terms only) on the context's own stream when `st` is another
    // one — run_pages: the copy stream, still busy with the files' uploads — so 1.3 GB of writes overlap the PCIe copies
    if (st != ctx->stream) {
        if (!ctx->ev_jpeg_zero) RT_CUDA_OK(ctx, cudaEventCreateWithFlags(&ctx->ev_jpeg_zero, cudaEventDisableTiming));
        RT_CUDA_OK(ctx, cudaMemsetAsync(ctx->d_jpeg_coef.p, 0, (size_t)blocks * 128, ctx->stream));
        RT_CUDA_OK(ctx, cudaEventRecord(ctx->ev_jpeg_zero, ctx->stream));
        RT_CUDA_OK(ctx, cudaStreamWaitEvent(st, ctx->ev_jpeg_zero, 0));
    } else RT_CUDA_OK(ctx, cudaMemsetAsync(ctx->d_jpeg_coef.p, 0, (size_t)blocks * 128, st));
    ctx->timer_stream = st;
    RT_LAUNCH_BEGIN(ctx, "jpeg_scan_kernel");
    jpeg_scan_kernel<<<n, 256, 0, st>>>(d_files, d_seg, d_status, ctx->d_jpeg_clean.as<unsigned char>(), d_clean_len);
    RT_LAUNCH_CHECK(ctx);
    const bool use_sub = n_sub_files > 0;
    total_seg = total_seg_all;
    if (use_sub) {
        // host tables: slot ranges per file (upper bounds: the interval lengths are only known on the device), file of every interval,
        // thread blocks of 128 slots of one file
        std::vector<int> sub_base(n + 1, 0);
        for (int i = 0; i < n; ++i) {
            const unsigned long long cap = file_sub[i] ? infos[i].ecs_len / JS_SUB_BYTES + (unsigned long long)infos[i].n_seg + 1 : 0ull;
            if ((unsigned long long)sub_base[i] + cap > 0x3fffffffULL) { ctx->set_error("jpeg decode: batch too large (sub-sequences)"); return RETTO_B200_ERR_CAPACITY; }
            sub_base[i + 1] = sub_base[i] + (int)cap;
        }
        const size_t S = (size_t)sub_base[n];
        size_t nsb = 0;
        for (int i = 0; i < n; ++i) nsb += ((size_t)(sub_base[i + 1] - sub_base[i]) + 127) / 128;
        const size_t tab_ints = (size_t)(n + 1) + (size_t)total_seg + nsb + (nsb + 1);
        const size_t tab_b = (tab_ints * 4 + 15) & ~size_t(15);
        RT_CUDA_OK(ctx, JB.h_sub.ensure(tab_b));
        int* ht = JB.h_sub.as<int>();
        int* h_sub_base = ht;
        int* h_seg_file = h_sub_base + (n + 1);
        int* h_sb_file = h_seg_file + total_seg;
        int* h_sb_first = h_sb_file + nsb;
        memcpy(h_sub_base, sub_base.data(), sizeof(int) * (size_t)(n + 1));
        {
            size_t g = 0, bq = 0;
            for (int i = 0; i < n; ++i) {
                for (int j = 0; j < infos[i].n_seg; ++j) h_seg_file[g++] = i;
                for (int s0 = sub_base[i]; s0 < sub_base[i + 1]; s0 += 128) { h_sb_file[bq] = i; h_sb_first[bq] = s0; ++bq; }
            }
            h_sb_first[nsb] = (int)S;
        }
        const size_t off_isub = tab_b;
        const size_t off_slots = (off_isub + (size_t)total_seg * 4 + 15) & ~size_t(15);
        const size_t off_states = off_slots + S * 16, off_syncs = off_states + S * 16, off_starts = off_syncs + S * 16, off_preds = off_starts + S * 16;
        const size_t off_cks = off_preds + S * 16;
        const size_t off_counts = off_cks + S * JS_CK * sizeof(JsCk), off_bitmaps = off_counts + S * JS_WORDS * 2;
        RT_CUDA_OK(ctx, ctx->d_jpeg_sub.ensure(off_bitmaps + S * JS_WORDS * 4, ctx->stream));
        char* db = ctx->d_jpeg_sub.as<char>();
        RT_CUDA_OK(ctx, cudaMemcpyAsync(db, ht, tab_b, cudaMemcpyHostToDevice, st));
        const int* d_sub_base = reinterpret_cast<const int*>(db);
        const int* d_seg_file = d_sub_base + (n + 1);
        const int* d_sb_file = d_seg_file + total_seg;
        const int* d_sb_first = d_sb_file + nsb;
        int* d_isub = reinterpret_cast<int*>(db + off_isub);
        JsSlot* d_slots = reinterpret_cast<JsSlot*>(db + off_slots);
        JsState* d_states = reinterpret_cast<JsState*>(db + off_states);
        JsSync* d_syncs = reinterpret_cast<JsSync*>(db + off_syncs);
        JsStart* d_starts = reinterpret_cast<JsStart*>(db + off_starts);
        int4* d_preds = reinterpret_cast<int4*>(db + off_preds);
        JsCk* d_cks = reinterpret_cast<JsCk*>(db + off_cks);
        unsigned short* d_counts = reinterpret_cast<unsigned short*>(db + off_counts);
        unsigned* d_bitmaps = reinterpret_cast<unsigned*>(db + off_bitmaps);
        const int smem = (int)(sizeof(HuffDev) * 6);
        if (!ctx->jpeg_sub_attr_set) {
            RT_CUDA_OK(ctx, cudaFuncSetAttribute(jpeg_sync_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
            RT_CUDA_OK(ctx, cudaFuncSetAttribute(jpeg_sync_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
            RT_CUDA_OK(ctx, cudaFuncSetAttribute(jpeg_huff_sub_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
            ctx->jpeg_sub_attr_set = true;
        }
        const unsigned char* d_clean = ctx->d_jpeg_clean.as<unsigned char>();
        ctx->timer_stream = st;
        RT_LAUNCH_BEGIN(ctx, "jpeg_sub_kernel");
        jpeg_sub_kernel<<<n, 256, 0, st>>>(d_files, d_seg, d_clean_len, d_sub_base, d_isub, d_slots);
        RT_LAUNCH_CHECK(ctx);
        ctx->timer_stream = st;
        RT_LAUNCH_BEGIN(ctx, "jpeg_sync_kernel<1>");
        jpeg_sync_kernel<1><<<(unsigned)nsb, 128, smem, st>>>(d_files, d_sb_file, d_sb_first, d_tab, d_seg, d_clean, d_clean_len, d_slots, d_bitmaps, d_counts, d_states, d_syncs, d_cks, d_wend);
        RT_LAUNCH_CHECK(ctx);
        ctx->timer_stream = st;
        RT_LAUNCH_BEGIN(ctx, "jpeg_sync_kernel<2>");
        jpeg_sync_kernel<2><<<(unsigned)nsb, 128, smem, st>>>(d_files, d_sb_file, d_sb_first, d_tab, d_seg, d_clean, d_clean_len, d_slots, d_bitmaps, d_counts, d_states, d_syncs, d_cks, d_wend);
        RT_LAUNCH_CHECK(ctx);
        ctx->timer_stream = st;
        RT_LAUNCH_BEGIN(ctx, "jpeg_resolve_kernel");
        jpeg_resolve_kernel<<<(unsigned)((total_seg + 127) / 128), 128, 0, st>>>(d_files, d_seg_file, (int)total_seg, d_sub_base, d_isub, d_slots, d_bitmaps, d_counts, d_syncs, d_cks, d_starts);
        RT_LAUNCH_CHECK(ctx);
        ctx->timer_stream = st;
        RT_LAUNCH_BEGIN(ctx, "jpeg_huff_sub_kernel");
        jpeg_huff_sub_kernel<<<(unsigned)nsb, 128, smem, st>>>(d_files, d_sb_file, d_sb_first, d_tab, d_seg, d_clean, d_clean_len, d_slots, d_starts, d_preds, ctx->d_jpeg_coef.as<short>(), d_wend);
        RT_LAUNCH_CHECK(ctx);
        ctx->timer_stream = st;
        RT_LAUNCH_BEGIN(ctx, "jpeg_dcfix_kernel");
        jpeg_dcfix_kernel<<<(unsigned)nsb * JS_FIX_LANES, 128, 0, st>>>(d_files, d_sb_file, d_sb_first, d_slots, d_starts, d_preds, ctx->d_jpeg_coef.as<short>());
        RT_LAUNCH_CHECK(ctx);
    }
    if (n_tblocks == 0) { ctx->timer_stream = nullptr; return RETTO_B200_OK; }
    ctx->timer_stream = st;
    RT_LAUNCH_BEGIN(ctx, "jpeg_huff_kernel");
    if (!ctx->jpeg_huff_attr_set) {
        RT_CUDA_OK(ctx, cudaFuncSetAttribute(jpeg_huff_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(HuffDev) * 6)));
        ctx->jpeg_huff_attr_set = true;
    }
    jpeg_huff_kernel<<<(unsigned)n_tblocks, tiered ? JH_T_THREADS : JH_THREADS, sizeof(HuffDev) * 6, st>>>(d_files, d_tb_file, d_tb_first, d_tab, d_seg,
                                                                 ctx->d_jpeg_clean.as<unsigned char>(), d_clean_len, ctx->d_jpeg_coef.as<short>(), tiered ? 1 : 0, d_wend);
    RT_LAUNCH_CHECK(ctx);
    ctx->timer_stream = nullptr;
    return RETTO_B200_OK;
}

// Phase 2 for files [first, first + n) of the batch whose entropy phase `owner` ran; kernels go on lane->stream (lane == owner unless
// run_pages pipelines units over two lanes).  d_out[k]: HWC u8 RGB of file first + k.
retto_b200_status rt_jpeg_pixels_enqueue(retto_b200_ctx* owner, retto_b200_ctx* lane, int first, int n, uint8_t* const* d_out) {
    const retto_b200_ctx::JpegBatch& JB = owner->jpeg;
    if (n <= 0) return RETTO_B200_OK;
    if (first < 0 || first + n > JB.n) { lane->set_error("jpeg decode: unit outside the decoded batch"); return RETTO_B200_ERR_INVALID_ARG; }
    // block width of the colour kernel: an exact fit when every page of the unit has the same width, 64 threads (512 px) otherwise
    bool uniform = true;
    for (int i = first + 1; i < first + n; ++i) uniform &= JB.X[i] == JB.X[first];
    const int jc_threads = uniform ? std::min(JC_THREADS, ((JB.X[first] + JC_PX - 1) / JC_PX + 31) / 32 * 32) : 64;
    // one upload: output pointers, then the two flat-grid prefix tables
    const size_t o_bytes = (sizeof(uint8_t*) * (size_t)n + 15) & ~size_t(15), p_bytes = (sizeof(unsigned) * ((size_t)n + 1) + 15) & ~size_t(15);
    int slot = -1;
    void* sp = nullptr;
    RT_TRY(rt_stage_begin(lane, o_bytes + 2 * p_bytes, &slot, &sp));
    memcpy(sp, d_out, sizeof(uint8_t*) * (size_t)n);
    unsigned* h_ip = reinterpret_cast<unsigned*>(reinterpret_cast<char*>(sp) + o_bytes);
    unsigned* h_cp = reinterpret_cast<unsigned*>(reinterpret_cast<char*>(sp) + o_bytes + p_bytes);
    bool all420 = true, wide420 = true;   // 4:2:0 throughout / chroma planes wider than two samples (jpeg_color420_kernel's domain)
    for (int i = first; i < first + n; ++i) { all420 &= JB.is420[i] != 0; wide420 &= JB.X[i] > 4; }
    static const bool no_c420 = getenv("RETTO_B200_JPEG_NO_C420") != nullptr;   // A/B: the one-row kernels
    const bool c420 = all420 && wide420 && !no_c420;
    unsigned long long ig = 0, cu = 0;
    for (int k = 0; k < n; ++k) {
        const int i = first + k;
        h_ip[k] = (unsigned)ig; h_cp[k] = (unsigned)cu;
        ig += (JB.n_blocks[i] + 31) / 32;
        cu += (unsigned long long)(c420 ? JB.Y[i] / 2 + 1 : JB.Y[i]) * ((JB.X[i] + jc_threads * JC_PX - 1) / (jc_threads * JC_PX));   // flat-grid units: row pairs / rows
    }
    h_ip[n] = (unsigned)ig; h_cp[n] = (unsigned)cu;
    if (ig > 0x7fffffffULL || cu > 0x7fffffffULL) { lane->stage_slots[slot].busy = false; lane->set_error("jpeg decode: unit too large"); return RETTO_B200_ERR_CAPACITY; }
    RT_TRY(rt_stage_commit(lane, lane->d_jpeg_out, slot, o_bytes + 2 * p_bytes));
    const char* op = lane->d_jpeg_out.as<char>();
    uint8_t* const* d_outs = reinterpret_cast<uint8_t* const*>(op);
    const unsigned* d_ip = reinterpret_cast<const unsigned*>(op + o_bytes);
    const unsigned* d_cp = reinterpret_cast<const unsigned*>(op + o_bytes + p_bytes);
    const char* dp = owner->d_jpeg_desc.as<char>();
    const JpegDev* d_files = reinterpret_cast<const JpegDev*>(dp) + first;
    const JpegTables* d_tab = reinterpret_cast<const JpegTables*>(dp + JB.head_bytes) + first;
    cudaStream_t st = lane->stream;
    bool same_blocks = true;
    for (int i = first + 1; i < first + n; ++i) same_blocks &= JB.n_blocks[i] == JB.n_blocks[first] && JB.Y[i] == JB.Y[first];
    RT_LAUNCH_BEGIN(lane, "jpeg_idct_kernel");
    if (uniform && same_blocks) jpeg_idct_kernel<<<dim3((JB.n_blocks[first] + 31) / 32, (unsigned)n), 256, 0, st>>>(d_files, d_tab, nullptr, n, owner->d_jpeg_coef.as<short>(), owner->d_jpeg_planes.as<unsigned char>());
    else jpeg_idct_kernel<<<(unsigned)ig, 256, 0, st>>>(d_files, d_tab, d_ip, n, owner->d_jpeg_coef.as<short>(), owner->d_jpeg_planes.as<unsigned char>());
    RT_LAUNCH_CHECK(lane);
    RT_LAUNCH_BEGIN(lane, "jpeg_color_kernel");
    if (uniform && same_blocks && c420)
        jpeg_color420_kernel<<<dim3((unsigned)((JB.X[first] + jc_threads * JC_PX - 1) / (jc_threads * JC_PX)), (unsigned)(JB.Y[first] / 2 + 1), (unsigned)n), jc_threads, 0, st>>>(
            d_files, owner->d_jpeg_planes.as<unsigned char>(), d_outs, nullptr, n);
    else if (c420)
        jpeg_color420_kernel<<<(unsigned)cu, jc_threads, 0, st>>>(d_files, owner->d_jpeg_planes.as<unsigned char>(), d_outs, d_cp, n);
    else if (uniform && same_blocks && all420)
        jpeg_color_kernel<1><<<dim3((unsigned)((JB.X[first] + jc_threads * JC_PX - 1) / (jc_threads * JC_PX)), (unsigned)JB.Y[first], (unsigned)n), jc_threads, 0, st>>>(
            d_files, owner->d_jpeg_planes.as<unsigned char>(), d_outs, nullptr, n);
    else if (uniform && same_blocks)
        jpeg_color_kernel<0><<<dim3((unsigned)((JB.X[first] + jc_threads * JC_PX - 1) / (jc_threads * JC_PX)), (unsigned)JB.Y[first], (unsigned)n), jc_threads, 0, st>>>(
            d_files, owner->d_jpeg_planes.as<unsigned char>(), d_outs, nullptr, n);
    else jpeg_color_kernel<0><<<(unsigned)cu, jc_threads, 0, st>>>(d_files, owner->d_jpeg_planes.as<unsigned char>(), d_outs, d_cp, n);
    RT_LAUNCH_CHECK(lane);
    return RETTO_B200_OK;
}

// ImageHelper::new_from_raw_img_flow for a batch of files: upload, decode, per-file status
extern "C" retto_b200_status retto_b200_decode_images(retto_b200_ctx* ctx, const retto_b200_encoded* h_imgs, int32_t n, uint8_t* const* d_out,
                                                      int32_t* h_status) {
    RtDeviceGuard _dg(ctx);
    if (!ctx || n < 0 || (n > 0 && (!h_imgs || !d_out || !h_status))) return RETTO_B200_ERR_INVALID_ARG;
    if (n == 0) return RETTO_B200_OK;
    std::vector<JpegInfo> infos(n);
    std::vector<int> ok;
    size_t blob = 0;
    for (int i = 0; i < n; ++i) {
        h_status[i] = rt_jpeg_parse(h_imgs[i].bytes, (size_t)h_imgs[i].n_bytes, &infos[i]);
        if (h_status[i] == RETTO_B200_OK) { ok.push_back(i); blob += ((size_t)h_imgs[i].n_bytes + 15 + 16) & ~size_t(15); }
    }
    retto_b200_status ret = RETTO_B200_OK;
    for (int i = 0; i < n; ++i) if (h_status[i] != RETTO_B200_OK) { ret = (retto_b200_status)h_status[i]; ctx->set_error("decode_images: file " + std::to_string(i) + " is not a supported baseline JPEG"); }
    if (ok.empty()) return ret;
    cudaStream_t st = ctx->stream;
    RT_CUDA_OK(ctx, ctx->d_jpeg_blob.ensure(blob + 16, st));
    std::vector<JpegInfo> sel(ok.size());
    std::vector<const uint8_t*> db(ok.size());
    std::vector<uint8_t*> dout(ok.size());
    size_t off = 0;
    for (size_t k = 0; k < ok.size(); ++k) {
        const int i = ok[k];
        uint8_t* d = ctx->d_jpeg_blob.as<uint8_t>() + off;
        RT_CUDA_OK(ctx, cudaMemcpyAsync(d, h_imgs[i].bytes, (size_t)h_imgs[i].n_bytes, cudaMemcpyHostToDevice, st));
        off += ((size_t)h_imgs[i].n_bytes + 15 + 16) & ~size_t(15);
        sel[k] = infos[i]; db[k] = d; dout[k] = d_out[i];
        if (!d_out[i]) { ctx->set_error("decode_images: null output"); return RETTO_B200_ERR_INVALID_ARG; }
    }
    RT_TRY(rt_jpeg_entropy_enqueue(ctx, st, sel.data(), db.data(), (int)ok.size()));
    RT_TRY(rt_jpeg_pixels_enqueue(ctx, ctx, 0, (int)ok.size(), dout.data()));
    std::vector<int> dev_status(ok.size());
    RT_CUDA_OK(ctx, cudaMemcpyAsync(dev_status.data(), ctx->jpeg_status_dev, sizeof(int) * ok.size(), cudaMemcpyDeviceToHost, st));
    RT_CUDA_OK(ctx, cudaStreamSynchronize(st));
    for (size_t k = 0; k < ok.size(); ++k)
        if (dev_status[k] != 0) { h_status[ok[k]] = dev_status[k]; ret = (retto_b200_status)dev_status[k]; ctx->set_error("decode_images: restart markers of file " + std::to_string(ok[k]) + " do not match its DRI header"); }
    return ret;
}
