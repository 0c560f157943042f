// Device side of h2j_b200: the YUV -> JPEG stage of the reference (src/Encoder.cpp:89-297, i.e. libavcodec's
// mjpeg encoder as the reference configures it) as sm_100a kernels.  One launch handles a batch of same-sized
// frames; frames never interact, so the batch index is simply a grid dimension.
//
//   K1  mbvar_kernel        luma 16x16 variance sums      -> rate-control input        (HBM bound, read 1 B/px)
//   K1b frame_setup_kernel  qscale, DQT, quantiser constants per frame
//   K2  fdct_quant_kernel   (range convert +) edge replicate + FDCT + quantise + zigzag + AC symbol histogram
//   K3  huffman_kernel      DC histogram, 4 optimal (package-merge) tables, code tables, JPEG header
//   K4  entropy_kernel      per-block bit lengths, decoupled look-back scan over tiles, bit packing
//   K5  stuff_kernel        0xFF -> 0xFF00 scan/compact behind the header, EOI, final size
//   K6  pack_kernel         optional: JPEGs of a batch packed back to back for one D2H copy
//   convert_pad_kernel      kernel 1 on its own: range convert + MCU padding to planes (h2j_convert_pad)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "h2j_math.cuh"

namespace h2j {

// ------------------------------------------------------------------------------------------------
// layout shared by host and device
// ------------------------------------------------------------------------------------------------
struct FrameLayout {
    int w, h;          // luma size
    int cw, ch;        // chroma size the encoder reads: w>>1, h>>1 (mpegvideo_enc.c load_input_picture)
    int y_pitch, c_pitch;
    long long u_off, v_off, frame_stride;  // bytes from the frame base / between frames
    int mcu_w, mcu_h, n_mcu, n_blocks;
    int aligned8;      // every row of every plane starts on an 8-byte boundary -> 64-bit loads
    int aligned16;     // ... 16-byte boundary -> 128-bit loads (mbvar)
    int range_mode;
    int fixed_qscale;
};

constexpr int kFdctMcusPerTile = 16;                       // 96 blocks, 3 warps
constexpr int kFdctThreads = kFdctMcusPerTile * 6;
constexpr int kFdctStageWords = 36;                         // 144-byte stride per staged block (bank-conflict free 128-bit I/O)
constexpr int kEntropyThreads = 128;                        // blocks per entropy tile
constexpr int kMaxBitsPerBlock = 27 * 64;                   // DC (16+11) + 63 * (16+11) — ZRLs only ever replace coefficients
constexpr int kEntropyBufWords = kEntropyThreads * kMaxBitsPerBlock / 32 + 2;
constexpr int kHuffThreads = 128;
constexpr int kStuffThreads = 512;
constexpr int kQscaleLutSize = 65536;

// Per-frame table block written by K1b/K3, read by K2/K4/K5.
struct FrameTab {
    uint32_t qpack[64];      // raster order: q | (bias*q) << 16
    uint8_t dqt_zz[64];      // DQT payload (zigzag order)
    uint8_t intra[64];       // raster order (inspection)
    uint32_t hcode[4][256];  // (code << 5) | size; classes: 0 DC luma, 1 DC chroma, 2 AC luma, 3 AC chroma
    uint8_t bits[4][17];
    uint8_t vals[4][256];
    int nvals[4];
    int qscale;
    int header_bytes;
    int status;              // h2j_status of this frame
    int pad_;
    long long mb_var_sum;
    long long scan_bits;
    long long stuffed_ff;
    long long jpeg_bytes;
};

// Per-frame state zeroed by one memset at the start of every batch.
struct FrameState {
    unsigned long long var_sum;
    unsigned long long scan_bits;   // written by the last entropy tile
    unsigned int hist[4][256];      // [0],[1] filled by K3 (DC), [2],[3] by K2 (AC)
};

__constant__ uint8_t c_zigzag[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48,
                                     41, 34, 27, 20, 13, 6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,
                                     30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};
// ff_mpeg1_default_intra_matrix, raster order
__constant__ uint8_t c_mpeg1_intra[64] = {8,  16, 19, 22, 26, 27, 29, 34, 16, 16, 22, 24, 27, 29, 34, 37, 19, 22, 26, 27, 29, 34,
                                          34, 38, 22, 22, 26, 27, 29, 34, 37, 40, 22, 26, 27, 29, 32, 35, 40, 48, 26, 27, 29, 32,
                                          35, 40, 48, 58, 26, 27, 29, 34, 38, 46, 56, 69, 27, 29, 35, 38, 46, 56, 69, 83};
// swscale limited->full LUTs, [0] luma, [1] chroma; filled by the host at create time from the closed form
__constant__ uint8_t c_range_lut[2][256];

// compile-time zigzag for the register-resident block
__host__ __device__ constexpr int zz_of(int k)
{
    constexpr int t[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48,
                           41, 34, 27, 20, 13, 6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,
                           30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};
    return t[k];
}

__device__ __forceinline__ uint2 ldg64(const uint8_t *p) { return __ldg(reinterpret_cast<const uint2 *>(p)); }
__device__ __forceinline__ uint4 ldg128(const uint8_t *p) { return __ldg(reinterpret_cast<const uint4 *>(p)); }

// ------------------------------------------------------------------------------------------------
// K1: mb_var_thread (mpegvideo_enc.c) — sum over macroblocks of ((norm1 - sum^2/256 + 628) >> 8)
// grid (mcu_h, n_frames), one thread per macroblock column, 16 independent 128-bit loads in flight.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) mbvar_kernel(const uint8_t *__restrict__ frames, FrameLayout L,
                                                    FrameState *__restrict__ state)
{
    const int f = blockIdx.y, my = blockIdx.x;
    const uint8_t *Y = frames + (long long)f * L.frame_stride;
    int local = 0;
    for (int mx = threadIdx.x; mx < L.mcu_w; mx += blockDim.x) {
        unsigned sum = 0, norm = 0;
        const int x0 = mx * 16;
        if (L.aligned16 && x0 + 16 <= L.w && L.range_mode == 0) {
            uint4 v[16];
#pragma unroll
            for (int r = 0; r < 16; r++) {
                const int y = min(my * 16 + r, L.h - 1);
                v[r] = ldg128(Y + (long long)y * L.y_pitch + x0);
            }
#pragma unroll
            for (int r = 0; r < 16; r++) {
                sum = __dp4a(v[r].x, 0x01010101u, sum); norm = __dp4a(v[r].x, v[r].x, norm);
                sum = __dp4a(v[r].y, 0x01010101u, sum); norm = __dp4a(v[r].y, v[r].y, norm);
                sum = __dp4a(v[r].z, 0x01010101u, sum); norm = __dp4a(v[r].z, v[r].z, norm);
                sum = __dp4a(v[r].w, 0x01010101u, sum); norm = __dp4a(v[r].w, v[r].w, norm);
            }
        } else {
            for (int r = 0; r < 16; r++) {
                const int y = min(my * 16 + r, L.h - 1);
                const uint8_t *row = Y + (long long)y * L.y_pitch;
                for (int c = 0; c < 16; c++) {
                    unsigned p = row[min(x0 + c, L.w - 1)];
                    if (L.range_mode) p = c_range_lut[0][p];
                    sum += p;
                    norm += p * p;
                }
            }
        }
        local += (int)(norm - ((sum * sum) >> 8) + 500u + 128u) >> 8;
    }
    // block reduce
    __shared__ int warp_sums[4];
#pragma unroll
    for (int o = 16; o; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
    if ((threadIdx.x & 31) == 0) warp_sums[threadIdx.x >> 5] = local;
    __syncthreads();
    if (threadIdx.x == 0) {
        long long t = 0;
        for (int i = 0; i < (int)(blockDim.x >> 5); i++) t += warp_sums[i];
        atomicAdd(&state[f].var_sum, (unsigned long long)t);
    }
}

// ------------------------------------------------------------------------------------------------
// K1b: first-frame rate control (ratecontrol.c) + MJPEG matrix set-up.  <<<n_frames, 64>>>
// The double-precision tail of the rate control (pow, float rounding, clipping) is folded by the host into
// qscale_lut[n], n = (int)(826.0 * sqrt(mb_var_sum) / 236.0)  — the device only evaluates the IEEE-exact
// part (sqrt, one multiply, one divide), so the choice is identical to the host libm's.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(64) frame_setup_kernel(FrameLayout L, const FrameState *__restrict__ state,
                                                         const uint8_t *__restrict__ qscale_lut, FrameTab *__restrict__ tabs)
{
    const int f = blockIdx.x, i = threadIdx.x;
    __shared__ int s_q;
    if (i == 0) {
        const long long var = (long long)state[f].var_sum;
        int q;
        if (L.fixed_qscale > 0) q = L.fixed_qscale;
        else {
            const double bits = __ddiv_rn(__dmul_rn(826.0, sqrt((double)var)), 236.0);  // predict_size()
            int n = (int)bits;
            if (n > kQscaleLutSize - 1) n = kQscaleLutSize - 1;
            if (n < 0) n = 0;
            q = qscale_lut[n];
        }
        s_q = q;
        tabs[f].qscale = q;
        tabs[f].mb_var_sum = var;
        tabs[f].status = 0;
    }
    __syncthreads();
    uint8_t m;
    uint32_t pk;
    quant_entry(s_q, c_mpeg1_intra[i], i, &m, &pk);
    tabs[f].qpack[i] = pk;
    tabs[f].intra[i] = m;
    // DQT is written in zigzag order: position k holds raster index zigzag[k]
    uint8_t mk;
    uint32_t pk2;
    quant_entry(s_q, c_mpeg1_intra[c_zigzag[i]], c_zigzag[i], &mk, &pk2);
    tabs[f].dqt_zz[i] = mk;
}

// ------------------------------------------------------------------------------------------------
// K2: pixels -> quantised zigzag coefficients.  One thread owns one 8x8 block, all 64 values in registers.
// A tile is 16 consecutive MCUs (linear MCU order, may wrap to the next MCU row):
//   warp 0: Y0/Y1 of the 16 MCUs  (32 horizontally adjacent blocks -> 256 contiguous bytes per pixel row)
//   warp 1: Y2/Y3
//   warp 2: Cb of the 16 MCUs (lanes 0-15), Cr (lanes 16-31)
// Results are staged in shared memory in MCU order and leave as one contiguous 12 KiB run of 128-bit stores.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void load_block_pixels(const uint8_t *__restrict__ P, int pitch, int pw, int ph, int bx, int by,
                                                  bool fast, const uint8_t *lut, int (&v)[64])
{
    if (fast && bx + 8 <= pw) {
        uint2 rows[8];
#pragma unroll
        for (int r = 0; r < 8; r++) {
            const int y = min(by + r, ph - 1);
            rows[r] = ldg64(P + (long long)y * pitch + bx);
        }
#pragma unroll
        for (int r = 0; r < 8; r++) {
            v[r * 8 + 0] = rows[r].x & 0xff;
            v[r * 8 + 1] = (rows[r].x >> 8) & 0xff;
            v[r * 8 + 2] = (rows[r].x >> 16) & 0xff;
            v[r * 8 + 3] = rows[r].x >> 24;
            v[r * 8 + 4] = rows[r].y & 0xff;
            v[r * 8 + 5] = (rows[r].y >> 8) & 0xff;
            v[r * 8 + 6] = (rows[r].y >> 16) & 0xff;
            v[r * 8 + 7] = rows[r].y >> 24;
        }
    } else {
#pragma unroll
        for (int r = 0; r < 8; r++) {
            const int y = min(by + r, ph - 1);
            const uint8_t *row = P + (long long)y * pitch;
#pragma unroll
            for (int c = 0; c < 8; c++) v[r * 8 + c] = row[min(bx + c, pw - 1)];
        }
    }
    if (lut) {
#pragma unroll
        for (int i = 0; i < 64; i++) v[i] = lut[v[i]];
    }
}

__global__ void __launch_bounds__(kFdctThreads) fdct_quant_kernel(const uint8_t *__restrict__ frames, FrameLayout L,
                                                                  const FrameTab *__restrict__ tabs,
                                                                  FrameState *__restrict__ state,
                                                                  int16_t *__restrict__ coefs,            // [frame][n_blocks][64] zigzag
                                                                  unsigned long long *__restrict__ masks, // [frame][n_blocks]
                                                                  int16_t *__restrict__ dcs,              // [frame][n_blocks]
                                                                  long long blocks_cap, int tiles_per_cta)
{
    __shared__ __align__(16) uint32_t s_stage[kFdctThreads * kFdctStageWords];
    __shared__ uint32_t s_qpack[64];
    __shared__ unsigned int s_hist[2][256];
    __shared__ unsigned long long s_mask[kFdctThreads];

    const int f = blockIdx.y;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid < 64) s_qpack[tid] = tabs[f].qpack[tid];
    for (int i = tid; i < 512; i += kFdctThreads) (&s_hist[0][0])[i] = 0;
    __syncthreads();

    const uint8_t *base = frames + (long long)f * L.frame_stride;
    const int mcu_l = warp < 2 ? (lane >> 1) : (lane & 15);
    const int n = warp < 2 ? (warp * 2 + (lane & 1)) : (4 + (lane >> 4));
    const int slot = mcu_l * 6 + n;
    const int cls = n < 4 ? 0 : 1;
    const uint8_t *lut = L.range_mode ? c_range_lut[cls] : nullptr;
    const uint8_t *P = n < 4 ? base : (n == 4 ? base + L.u_off : base + L.v_off);
    const int pitch = n < 4 ? L.y_pitch : L.c_pitch;
    const int pw = n < 4 ? L.w : L.cw, ph = n < 4 ? L.h : L.ch;
    const int n_tiles = (L.n_mcu + kFdctMcusPerTile - 1) / kFdctMcusPerTile;

    for (int t = 0; t < tiles_per_cta; t++) {
        const int tile = blockIdx.x * tiles_per_cta + t;
        if (tile >= n_tiles) break;
        const int m = tile * kFdctMcusPerTile + mcu_l;
        const bool valid = m < L.n_mcu;
        unsigned long long mask = 0;
        if (valid) {
            const int my = m / L.mcu_w, mx = m - my * L.mcu_w;
            const int bx = n < 4 ? mx * 16 + (n & 1) * 8 : mx * 8;
            const int by = n < 4 ? my * 16 + (n >> 1) * 8 : my * 8;
            int v[64];
            load_block_pixels(P, pitch, pw, ph, bx, by, L.aligned8 != 0, lut, v);
            fdct_8x8(v);
            v[0] = quant_dc(v[0]);
#pragma unroll
            for (int i = 1; i < 64; i++) v[i] = quant_ac(v[i], s_qpack[i]);
            // zigzag + pack two levels per word + non-zero mask
            uint32_t *dst = s_stage + slot * kFdctStageWords;
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
                uint4 o;
                uint32_t wv[4];
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const int lo = v[zz_of(2 * (j + k))], hi = v[zz_of(2 * (j + k) + 1)];
                    wv[k] = (uint32_t)(lo & 0xffff) | ((uint32_t)hi << 16);
                    if (lo != 0) mask |= 1ull << (2 * (j + k));
                    if (hi != 0) mask |= 1ull << (2 * (j + k) + 1);
                }
                o.x = wv[0]; o.y = wv[1]; o.z = wv[2]; o.w = wv[3];
                *reinterpret_cast<uint4 *>(dst + j) = o;
            }
        }
        s_mask[slot] = mask;
        __syncthreads();

        // AC symbol statistics of the thread's own block (ff_mjpeg_encode_coef / record_block, AC part)
        if (valid) {
            const int16_t *lv = reinterpret_cast<const int16_t *>(s_stage + slot * kFdctStageWords);
            unsigned long long mm = mask & ~1ull;
            int prev = 0;
            unsigned int *hist = s_hist[cls];
            while (mm) {
                const int k = __ffsll((long long)mm) - 1;
                mm &= mm - 1;
                const int run = k - prev - 1;
                prev = k;
                const int a = abs((int)lv[k]);
                const int nb = 32 - __clz(a);
                if (run >= 16) atomicAdd(&hist[0xf0], (unsigned)(run >> 4));
                atomicAdd(&hist[((run & 15) << 4) | nb], 1u);
            }
            if (prev < 63) atomicAdd(&hist[0], 1u);
        }

        // contiguous copy-out: coefficient rows, masks, DC side array
        const long long blk0 = (long long)f * blocks_cap + (long long)tile * kFdctThreads;
        const int blocks_here = min(kFdctThreads, L.n_blocks - tile * kFdctThreads);
        uint4 *gdst = reinterpret_cast<uint4 *>(coefs + blk0 * 64);
        for (int c = tid; c < blocks_here * 8; c += kFdctThreads) {
            const int b = c >> 3, part = c & 7;
            gdst[c] = *reinterpret_cast<const uint4 *>(s_stage + b * kFdctStageWords + part * 4);
        }
        if (tid < blocks_here) {
            masks[blk0 + tid] = s_mask[tid];
            dcs[blk0 + tid] = (int16_t)(s_stage[tid * kFdctStageWords] & 0xffff);
        }
        __syncthreads();
    }

    for (int i = tid; i < 512; i += kFdctThreads) {
        const unsigned c = (&s_hist[0][0])[i];
        if (c) atomicAdd(&state[f].hist[2 + (i >> 8)][i & 255], c);
    }
}

// ------------------------------------------------------------------------------------------------
// K3: optimal Huffman tables (mjpegenc_huffman.c) + code tables + header (mjpegenc_common.c).
// <<<n_frames, 128>>>: the DC histogram is built by all threads from the DC side array; then lane 0 of warp t
// builds table t.  AV_QSORT is reproduced step for step because it is not stable and the order of equal
// counts/lengths decides the DHT bytes.  The package-merge keeps, per level, only the item probabilities and
// the running count of leaves; that is enough to recover the code lengths and gives the same answer as the
// reference's list-copying formulation (leaves are consumed in sorted order, packages in pairs).
// ------------------------------------------------------------------------------------------------
struct HuffPair { int a, b; };  // {value, prob} or {code, length}

__device__ void av_qsort_pairs(HuffPair *p, int num)
{
#define H2J_CMP(x, y) ((x)->b - (y)->b)
#define H2J_SWAP(x, y) do { HuffPair t_ = (x); (x) = (y); (y) = t_; } while (0)
    HuffPair *stack[64][2];
    int sp = 1;
    stack[0][0] = p;
    stack[0][1] = p + num - 1;
    while (sp) {
        HuffPair *start = stack[--sp][0];
        HuffPair *end = stack[sp][1];
        while (start < end) {
            if (start < end - 1) {
                int checksort = 0;
                HuffPair *right = end - 2;
                HuffPair *left = start + 1;
                HuffPair *mid = start + ((end - start) >> 1);
                if (H2J_CMP(start, end) > 0) {
                    if (H2J_CMP(end, mid) > 0) H2J_SWAP(*start, *mid);
                    else H2J_SWAP(*start, *end);
                } else {
                    if (H2J_CMP(start, mid) > 0) H2J_SWAP(*start, *mid);
                    else checksort = 1;
                }
                if (H2J_CMP(mid, end) > 0) {
                    H2J_SWAP(*mid, *end);
                    checksort = 0;
                }
                if (start == end - 2) break;
                H2J_SWAP(end[-1], *mid);
                while (left <= right) {
                    while (left <= right && H2J_CMP(left, end - 1) < 0) left++;
                    while (left <= right && H2J_CMP(right, end - 1) > 0) right--;
                    if (left <= right) {
                        H2J_SWAP(*left, *right);
                        left++;
                        right--;
                    }
                }
                H2J_SWAP(end[-1], *left);
                if (checksort && (mid == left - 1 || mid == left)) {
                    mid = start;
                    while (mid < end && H2J_CMP(mid, mid + 1) <= 0) mid++;
                    if (mid == end) break;
                }
                if (end - left < left - start) {
                    stack[sp][0] = start;
                    stack[sp++][1] = right;
                    start = left + 1;
                } else {
                    stack[sp][0] = left + 1;
                    stack[sp++][1] = end;
                    end = right;
                }
            } else {
                if (H2J_CMP(start, end) > 0) H2J_SWAP(*start, *end);
                break;
            }
        }
    }
#undef H2J_CMP
#undef H2J_SWAP
}

constexpr int kPmMaxItems = 516;  // a level holds at most 2 * 257 items
struct HuffScratch {
    HuffPair sorted[258];
    HuffPair distinct[258];
    int prob[2][kPmMaxItems];
    unsigned short leaves[17][kPmMaxItems];  // leaves[t][p] = number of leaf items among the first p items of level t
    int nitems[17];
    unsigned char nbits_by_value[260];
};

__device__ void build_one_table(const unsigned int *__restrict__ hist, HuffScratch *S, uint8_t *bits, uint8_t *vals, int *nvals_out,
                                uint32_t *hcode)
{
    // ff_mjpeg_encode_huffman_close
    int nval = 0;
    for (int i = 0; i < 256; i++)
        if (hist[i]) {
            S->sorted[nval].a = i;
            S->sorted[nval].b = (int)hist[i];
            nval++;
        }
    S->sorted[nval].a = 256;
    S->sorted[nval].b = 0;
    const int size = nval + 1;
    av_qsort_pairs(S->sorted, size);

    // ff_mjpegenc_huffman_compute_bits, max_length 16: levels 0..15 take leaves, level 16 only packages
    int cur = 0;
    for (int k = 0; k < size; k++) { S->prob[cur][k] = S->sorted[k].b; S->leaves[0][k] = (unsigned short)k; }
    S->leaves[0][size] = (unsigned short)size;
    S->nitems[0] = size;
    for (int t = 1; t <= 16; t++) {
        const int *pp = S->prob[cur];
        int *np = S->prob[cur ^ 1];
        const int from_n = S->nitems[t - 1];
        int i = (t < 16) ? 0 : size, j = 0, k = 0, nl = 0;
        S->leaves[t][0] = 0;
        while (i < size || j + 1 < from_n) {
            if (i < size && (j + 1 >= from_n || S->sorted[i].b < pp[j] + pp[j + 1])) {
                np[k] = S->sorted[i].b;
                i++;
                nl++;
            } else {
                np[k] = pp[j] + pp[j + 1];
                j += 2;
            }
            k++;
            S->leaves[t][k] = (unsigned short)nl;
        }
        S->nitems[t] = k;
        cur ^= 1;
    }
    for (int i = 0; i < 257; i++) S->nbits_by_value[i] = 0;
    {
        int p = (size - 1 < S->nitems[16]) ? size - 1 : S->nitems[16];
        for (int t = 16; t >= 0 && p > 0; t--) {
            const int nl = S->leaves[t][p];
            for (int r = 0; r < nl; r++) S->nbits_by_value[S->sorted[r].a]++;
            p = 2 * (p - nl);
        }
    }
    int j = 0;
    for (int i = 0; i < 256; i++)
        if (S->nbits_by_value[i]) {
            S->distinct[j].a = i;
            S->distinct[j].b = S->nbits_by_value[i];
            j++;
        }
    av_qsort_pairs(S->distinct, nval);  // by length
    for (int i = 0; i <= 16; i++) bits[i] = 0;
    for (int i = 0; i < nval; i++) {
        vals[i] = (uint8_t)S->distinct[i].a;
        bits[S->distinct[i].b]++;
    }
    for (int i = nval; i < 256; i++) vals[i] = 0;
    *nvals_out = nval;
    // ff_mjpeg_build_huffman_codes
    for (int i = 0; i < 256; i++) hcode[i] = 0;
    int k = 0, code = 0;
    for (int i = 1; i <= 16; i++) {
        const int nb = bits[i];
        for (int q = 0; q < nb; q++) {
            const int sym = vals[k++];
            hcode[sym] = ((uint32_t)code << 5) | (uint32_t)i;
            code++;
        }
        code <<= 1;
    }
}

// previous block of the same component in coding order, -1 if none (predictor 128)
__device__ __forceinline__ int dc_pred_index(int b)
{
    const int n = b % 6, m = b / 6;
    if (n >= 1 && n <= 3) return b - 1;
    if (m == 0) return -1;
    return n == 0 ? b - 3 : b - 6;
}

__global__ void __launch_bounds__(kHuffThreads) huffman_kernel(FrameLayout L, FrameTab *__restrict__ tabs, FrameState *__restrict__ state,
                                                               const int16_t *__restrict__ dcs, long long blocks_cap,
                                                               uint8_t *__restrict__ out, long long out_cap,
                                                               const char *__restrict__ comment, int comment_len)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    HuffScratch *scratch = reinterpret_cast<HuffScratch *>(smem_raw);
    __shared__ unsigned int s_dc[2][16];
    const int f = blockIdx.x, tid = threadIdx.x;
    if (tid < 32) (&s_dc[0][0])[tid] = 0;
    __syncthreads();
    const int16_t *dc = dcs + (long long)f * blocks_cap;
    for (int b = tid; b < L.n_blocks; b += kHuffThreads) {
        const int pi = dc_pred_index(b);
        const int diff = (int)dc[b] - (pi < 0 ? 128 : (int)dc[pi]);
        const int nb = diff ? 32 - __clz(abs(diff)) : 0;
        atomicAdd(&s_dc[(b % 6) < 4 ? 0 : 1][nb], 1u);
    }
    __syncthreads();
    if (tid < 32) state[f].hist[tid >> 4][tid & 15] = (&s_dc[0][0])[tid];
    __syncthreads();
    FrameTab *T = tabs + f;
    if ((tid & 31) == 0) {
        const int t = tid >> 5;
        build_one_table(state[f].hist[t], scratch + t, T->bits[t], T->vals[t], &T->nvals[t], T->hcode[t]);
    }
    __syncthreads();
    // ---- header: SOI, COM, DQT, DHT, SOF0, SOS (ff_mjpeg_encode_picture_header) -------------------
    if (tid == 0) {
        uint8_t *o = out + (long long)f * out_cap;
        int p = 0;
        auto put8 = [&](int v) { if (p < out_cap) o[p] = (uint8_t)v; p++; };
        auto put16 = [&](int v) { put8(v >> 8); put8(v & 0xff); };
        put16(0xffd8);
        put16(0xfffe); put16(comment_len + 3);
        for (int i = 0; i < comment_len; i++) put8(comment[i]);
        put8(0);
        put16(0xffdb); put16(2 + 65); put8(0);
        for (int i = 0; i < 64; i++) put8(T->dqt_zz[i]);
        put16(0xffc4);
        int size = 2;
        for (int t = 0; t < 4; t++) size += 17 + T->nvals[t];
        put16(size);
        for (int t = 0; t < 4; t++) {
            const int order[4] = {0, 1, 2, 3};  // DC luma, DC chroma, AC luma, AC chroma
            const int tt = order[t];
            put8(((tt >> 1) << 4) | (tt & 1));
            for (int i = 1; i <= 16; i++) put8(T->bits[tt][i]);
            for (int i = 0; i < T->nvals[tt]; i++) put8(T->vals[tt][i]);
        }
        put16(0xffc0); put16(17); put8(8); put16(L.h); put16(L.w); put8(3);
        put8(1); put8(0x22); put8(0);
        put8(2); put8(0x11); put8(0);
        put8(3); put8(0x11); put8(0);
        put16(0xffda); put16(12); put8(3);
        put8(1); put8(0x00);
        put8(2); put8(0x11);
        put8(3); put8(0x11);
        put8(0); put8(63); put8(0);
        T->header_bytes = p;
        if (p > out_cap) T->status = -4;
    }
}

// ------------------------------------------------------------------------------------------------
// K4: entropy coding (mjpegenc.c record_block + ff_mjpeg_encode_picture_frame).
// A tile is 128 consecutive blocks of one frame.  Phase 1: every thread walks the non-zero mask of its
// block and sums code lengths; block-wide exclusive scan.  Phase 2: the tile publishes (length, trailing
// bits) and resolves its exclusive prefix by decoupled look-back over the frame's earlier tiles.  Phase 3:
// threads emit their codes into the tile's shared-memory bit buffer.  Phase 4: the buffer is shifted to
// the tile's global bit position and stored; a 32-bit word is written by the tile that holds its last bit,
// with the bits of earlier tiles arriving through the look-back payload — no atomics, no pre-zeroed output.
//
// Descriptor (one 64-bit word, so a single relaxed store/load carries everything):
//   [63:62] status (0 invalid, 1 tile aggregate, 2 inclusive prefix)   [61:31] bit length   [30:0] last 31 bits
// ------------------------------------------------------------------------------------------------
struct BitRun { unsigned int len; unsigned int tail; };  // tail: the last min(len,31) bits, right aligned

__device__ __forceinline__ BitRun bitrun_concat(BitRun x, BitRun y)  // x then y
{
    BitRun r;
    r.len = x.len + y.len;
    r.tail = (y.len >= 31) ? y.tail : (((x.tail << y.len) | y.tail) & 0x7fffffffu);
    return r;
}
__device__ __forceinline__ unsigned long long desc_pack(unsigned status, BitRun r)
{
    return ((unsigned long long)status << 62) | ((unsigned long long)r.len << 31) | (unsigned long long)(r.tail & 0x7fffffffu);
}
__device__ __forceinline__ unsigned desc_status(unsigned long long d) { return (unsigned)(d >> 62); }
__device__ __forceinline__ BitRun desc_run(unsigned long long d)
{
    BitRun r;
    r.len = (unsigned)((d >> 31) & 0x7fffffffu);
    r.tail = (unsigned)(d & 0x7fffffffu);
    return r;
}
__device__ __forceinline__ unsigned long long ld_desc(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_desc(unsigned long long *p, unsigned long long v)
{
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

struct BitSink {  // appends MSB-first into a zeroed shared-memory word array
    unsigned int *buf;
    unsigned long long acc;
    int fill;
    int widx;
    __device__ __forceinline__ void init(unsigned int *b, unsigned pos) { buf = b; acc = 0; fill = (int)(pos & 31); widx = (int)(pos >> 5); }
    __device__ __forceinline__ void put(unsigned bits, int len)
    {
        acc = (acc << len) | bits;
        fill += len;
        if (fill >= 32) {
            atomicOr(&buf[widx++], (unsigned)(acc >> (fill - 32)));
            fill -= 32;
        }
    }
    __device__ __forceinline__ void flush() { if (fill > 0) atomicOr(&buf[widx], (unsigned)(acc << (32 - fill))); }
};

// One pass over a block.  EMIT=false: returns the bit length.  EMIT=true: writes the bits.
template <bool EMIT>
__device__ __forceinline__ unsigned walk_block(const int16_t *__restrict__ cb, unsigned long long mask, int dc_diff,
                                               const uint32_t *__restrict__ hdc, const uint32_t *__restrict__ hac, BitSink *sink)
{
    unsigned total = 0;
    {
        const int nb = dc_diff ? 32 - __clz(abs(dc_diff)) : 0;
        const uint32_t e = hdc[nb];
        const int sz = e & 31;
        if (EMIT) {
            const unsigned mant = (unsigned)(dc_diff < 0 ? dc_diff - 1 : dc_diff) & ((1u << nb) - 1u);
            sink->put(((e >> 5) << nb) | mant, sz + nb);
        } else total += sz + nb;
    }
    unsigned long long mm = mask & ~1ull;
    int prev = 0;
    const uint32_t zrl = hac[0xf0];
    while (mm) {
        const int k = __ffsll((long long)mm) - 1;
        mm &= mm - 1;
        int run = k - prev - 1;
        prev = k;
        const int val = (int)__ldg(cb + k);
        const int nb = 32 - __clz(abs(val));
        while (run >= 16) {
            if (EMIT) sink->put(zrl >> 5, zrl & 31);
            else total += zrl & 31;
            run -= 16;
        }
        const uint32_t e = hac[(run << 4) | nb];
        const int sz = e & 31;
        if (EMIT) {
            const unsigned mant = (unsigned)(val < 0 ? val - 1 : val) & ((1u << nb) - 1u);
            sink->put(((e >> 5) << nb) | mant, sz + nb);
        } else total += sz + nb;
    }
    if (prev < 63) {
        const uint32_t e = hac[0];
        if (EMIT) sink->put(e >> 5, e & 31);
        else total += e & 31;
    }
    return total;
}

__global__ void __launch_bounds__(kEntropyThreads) entropy_kernel(FrameLayout L, FrameTab *__restrict__ tabs, FrameState *__restrict__ state,
                                                                  const int16_t *__restrict__ coefs,
                                                                  const unsigned long long *__restrict__ masks,
                                                                  const int16_t *__restrict__ dcs, long long blocks_cap,
                                                                  unsigned long long *__restrict__ descs,  // [frame][tiles_per_frame]
                                                                  unsigned int *__restrict__ ticket, int tiles_per_frame,
                                                                  uint32_t *__restrict__ scan, long long scan_cap_words)
{
    extern __shared__ __align__(16) unsigned int s_bits[];  // kEntropyBufWords (+1 guard in front)
    __shared__ uint32_t s_hdc[2][16];
    __shared__ uint32_t s_hac[2][256];
    __shared__ unsigned s_warp[kEntropyThreads / 32];
    __shared__ unsigned s_ticket;
    __shared__ BitRun s_excl;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_ticket = atomicAdd(ticket, 1u);
    __syncthreads();
    const int f = (int)(s_ticket / (unsigned)tiles_per_frame);
    const int tile = (int)(s_ticket % (unsigned)tiles_per_frame);
    const FrameTab *T = tabs + f;
    for (int i = tid; i < 512; i += kEntropyThreads) (&s_hac[0][0])[i] = T->hcode[2 + (i >> 8)][i & 255];
    if (tid < 32) (&s_hdc[0][0])[tid] = T->hcode[tid >> 4][tid & 15];
    __syncthreads();

    const int b = tile * kEntropyThreads + tid;
    const bool valid = b < L.n_blocks;
    const long long gb = (long long)f * blocks_cap + b;
    unsigned long long mask = 0;
    int dc_diff = 0;
    const int cls = (b % 6) < 4 ? 0 : 1;
    const int16_t *cb = coefs + gb * 64;
    unsigned len = 0;
    if (valid) {
        mask = masks[gb];
        const int pi = dc_pred_index(b);
        const int16_t *dc = dcs + (long long)f * blocks_cap;
        dc_diff = (int)dc[b] - (pi < 0 ? 128 : (int)dc[pi]);
        len = walk_block<false>(cb, mask, dc_diff, s_hdc[cls], s_hac[cls], nullptr);
    }
    // block-wide exclusive scan of len
    unsigned incl = len;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    unsigned warp_off = 0, tile_len = 0;
#pragma unroll
    for (int w = 0; w < kEntropyThreads / 32; w++) {
        if (w < warp) warp_off += s_warp[w];
        tile_len += s_warp[w];
    }
    const unsigned off = warp_off + incl - len;
    const int n_words = (int)((tile_len + 31) >> 5);
    // word [0] is a guard in front so that phase 4 can read "the word before the first"
    for (int i = tid; i <= n_words + 1; i += kEntropyThreads) s_bits[i] = 0;
    __syncthreads();
    if (valid) {
        BitSink sink;
        sink.init(s_bits + 1, off);
        walk_block<true>(cb, mask, dc_diff, s_hdc[cls], s_hac[cls], &sink);
        sink.flush();
    }
    __syncthreads();

    // ---- publish aggregate, resolve exclusive prefix (warp 0) --------------------------------------
    unsigned long long *D = descs + (long long)f * tiles_per_frame;
    if (warp == 0) {
        BitRun own;
        own.len = tile_len;
        if (tile_len == 0) own.tail = 0;
        else {
            // last min(len,31) bits of the local stream
            const unsigned endw = (tile_len - 1) >> 5;          // word holding the last bit
            const unsigned used = ((tile_len - 1) & 31) + 1;    // bits used in it
            const unsigned long long two = ((unsigned long long)(endw ? s_bits[endw] : 0u) << 32) | s_bits[endw + 1];
            const unsigned last32 = (unsigned)(two >> (32 - used));
            own.tail = tile_len >= 31 ? (last32 & 0x7fffffffu) : (last32 & ((1u << tile_len) - 1u));
        }
        if (lane == 0 && tile > 0) st_desc(&D[tile], desc_pack(1, own));
        BitRun excl;
        excl.len = 0;
        excl.tail = 0;
        if (tile > 0) {
            BitRun running;
            running.len = 0;
            running.tail = 0;
            int basei = tile - 1;
            while (true) {
                const int idx = basei - lane;
                unsigned long long d;
                if (idx >= 0) {
                    do { d = ld_desc(&D[idx]); } while (desc_status(d) == 0);
                } else {
                    BitRun z; z.len = 0; z.tail = 0;
                    d = desc_pack(2, z);
                }
                const unsigned pm = __ballot_sync(0xffffffffu, desc_status(d) == 2);
                const int stop = pm ? (__ffs(pm) - 1) : 31;
                BitRun acc = running;
                for (int l = 0; l <= stop; l++) {
                    const unsigned long long dl = __shfl_sync(0xffffffffu, d, l);
                    acc = bitrun_concat(desc_run(dl), acc);
                }
                running = acc;
                if (pm) break;
                basei -= 32;
            }
            excl = running;
        }
        if (lane == 0) {
            st_desc(&D[tile], desc_pack(2, bitrun_concat(excl, own)));
            s_excl = excl;
            if (tile == tiles_per_frame - 1) state[f].scan_bits = (unsigned long long)excl.len + tile_len;
        }
    }
    __syncthreads();

    // ---- phase 4: shift to the global bit position and store big-endian words ----------------------
    const unsigned P = s_excl.len;
    const unsigned s = P & 31;
    const long long W0 = P >> 5;
    const unsigned long long endbit = (unsigned long long)P + tile_len;
    long long Wend = (long long)(endbit >> 5);
    if (tile == tiles_per_frame - 1 && (endbit & 31)) Wend++;  // the frame's last, incomplete word
    if (tid == 0) s_bits[0] = s_excl.tail;                      // bits of earlier tiles living in word W0
    __syncthreads();
    uint32_t *gs = scan + (long long)f * scan_cap_words;
    bool overflow = false;
    for (long long W = W0 + tid; W < Wend; W += kEntropyThreads) {
        const int j = (int)(W - W0);
        const unsigned hi = s_bits[j], lo = s_bits[j + 1];       // local words j-1 and j
        const unsigned v = s ? ((hi << (32 - s)) | (lo >> s)) : lo;
        if (W < scan_cap_words) gs[W] = __byte_perm(v, 0, 0x0123);
        else overflow = true;
    }
    if (overflow) tabs[f].status = -4;
}

// ------------------------------------------------------------------------------------------------
// K5: ff_mjpeg_escape_FF + picture trailer.  One CTA per frame walks the frame's scan bytes in chunks,
// counts 0xFF bytes, scans, and writes every byte (and a 0x00 after each 0xFF) to its final position
// behind the header; the last byte is padded with ones first.  Then EOI and the final size.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kStuffThreads) stuff_kernel(FrameTab *__restrict__ tabs, const FrameState *__restrict__ state,
                                                              const uint32_t *__restrict__ scan, long long scan_cap_words,
                                                              uint8_t *__restrict__ out, long long out_cap)
{
    const int f = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    __shared__ unsigned s_warp[kStuffThreads / 32];
    __shared__ unsigned long long s_carry;
    FrameTab *T = tabs + f;
    const long long bits = (long long)state[f].scan_bits;
    const long long nbytes = (bits + 7) >> 3;
    const int pad = (int)(nbytes * 8 - bits);
    const unsigned padmask = (1u << pad) - 1u;
    const long long hdr = T->header_bytes;
    const uint32_t *gs = scan + (long long)f * scan_cap_words;
    uint8_t *o = out + (long long)f * out_cap;
    if (tid == 0) s_carry = 0;
    __syncthreads();
    const long long nwords = (nbytes + 3) >> 2;
    for (long long base = 0; base < nwords; base += kStuffThreads) {
        const long long wi = base + tid;
        unsigned w = 0;
        int nb = 0;
        if (wi < nwords && wi < scan_cap_words) {
            w = gs[wi];
            const long long rem = nbytes - wi * 4;
            nb = rem >= 4 ? 4 : (int)rem;
            if (wi * 4 + nb == nbytes && pad) w |= padmask << (8 * (nb - 1));
        }
        unsigned cnt = 0;
#pragma unroll
        for (int j = 0; j < 4; j++) cnt += (j < nb && ((w >> (8 * j)) & 0xff) == 0xff) ? 1u : 0u;
        unsigned incl = cnt;
#pragma unroll
        for (int ofs = 1; ofs < 32; ofs <<= 1) {
            const unsigned t = __shfl_up_sync(0xffffffffu, incl, ofs);
            if (lane >= ofs) incl += t;
        }
        if (lane == 31) s_warp[warp] = incl;
        __syncthreads();
        unsigned woff = 0, chunk_total = 0;
#pragma unroll
        for (int k = 0; k < kStuffThreads / 32; k++) {
            if (k < warp) woff += s_warp[k];
            chunk_total += s_warp[k];
        }
        const unsigned long long carry = s_carry;
        long long pos = hdr + wi * 4 + (long long)carry + woff + incl - cnt;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            if (j < nb) {
                const unsigned byte = (w >> (8 * j)) & 0xff;
                if (pos < out_cap) o[pos] = (uint8_t)byte;
                pos++;
                if (byte == 0xff) {
                    if (pos < out_cap) o[pos] = 0;
                    pos++;
                }
            }
        }
        __syncthreads();
        if (tid == 0) s_carry = carry + chunk_total;
        __syncthreads();
    }
    if (tid == 0) {
        const long long ff = (long long)s_carry;
        const long long end = hdr + nbytes + ff;
        if (end + 2 <= out_cap) { o[end] = 0xff; o[end + 1] = 0xd9; }
        else T->status = -4;
        T->scan_bits = bits;
        T->stuffed_ff = ff;
        T->jpeg_bytes = end + 2;
    }
}

// ------------------------------------------------------------------------------------------------
// K6: pack the JPEGs of a batch back to back (so one D2H copy moves exactly the bytes produced).
// offsets[n+1] is computed by block 0 of pack_offsets_kernel.
// ------------------------------------------------------------------------------------------------
__global__ void pack_offsets_kernel(const FrameTab *__restrict__ tabs, int n, long long out_cap, unsigned long long *__restrict__ offsets,
                                    int *__restrict__ status)
{
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        unsigned long long acc = 0;
        for (int i = 0; i < n; i++) {
            offsets[i] = acc;
            const int st = tabs[i].status;
            status[i] = st;
            long long sz = tabs[i].jpeg_bytes;
            if (st != 0 || sz > out_cap) sz = 0;
            acc += (unsigned long long)sz;
        }
        offsets[n] = acc;
    }
}

__global__ void __launch_bounds__(256) pack_kernel(const uint8_t *__restrict__ out, long long out_cap,
                                                   const unsigned long long *__restrict__ offsets, uint8_t *__restrict__ packed)
{
    const int f = blockIdx.y;
    const long long size = (long long)(offsets[f + 1] - offsets[f]);
    const uint8_t *src = out + (long long)f * out_cap;
    uint8_t *dst = packed + offsets[f];
    // 16 source bytes per thread; destination alignment is arbitrary, so the stores are byte wide
    for (long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 16; i < size; i += (long long)gridDim.x * blockDim.x * 16) {
        if (i + 16 <= size) {
            const uint4 v = *reinterpret_cast<const uint4 *>(src + i);
            const unsigned w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int k = 0; k < 16; k++) dst[i + k] = (uint8_t)(w[k >> 2] >> (8 * (k & 3)));
        } else {
            for (long long k = i; k < size; k++) dst[k] = src[k];
        }
    }
}

// ------------------------------------------------------------------------------------------------
// kernel 1 on its own: range conversion + MCU edge replication into padded planes (h2j_convert_pad).
// One thread per 16 output bytes; 128-bit loads when the source row is 16-byte aligned.
// grid (ceil(padded_w/16 / 128), padded_h, 3)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) convert_pad_kernel(const uint8_t *__restrict__ frame, FrameLayout L, int range_mode,
                                                          uint8_t *__restrict__ oy, uint8_t *__restrict__ ou, uint8_t *__restrict__ ov)
{
    __shared__ uint8_t s_lut[256];
    const int plane = blockIdx.z;
    const int pw = plane ? L.cw : L.w, ph = plane ? L.ch : L.h;
    const int padw = plane ? L.mcu_w * 8 : L.mcu_w * 16, padh = plane ? L.mcu_h * 8 : L.mcu_h * 16;
    const int pitch = plane ? L.c_pitch : L.y_pitch;
    const uint8_t *P = frame + (plane == 0 ? 0 : (plane == 1 ? L.u_off : L.v_off));
    uint8_t *O = plane == 0 ? oy : (plane == 1 ? ou : ov);
    for (int i = threadIdx.x; i < 256; i += blockDim.x) s_lut[i] = range_mode ? c_range_lut[plane ? 1 : 0][i] : (uint8_t)i;
    __syncthreads();
    const int y = blockIdx.y;
    if (y >= padh) return;
    const int x0 = (blockIdx.x * blockDim.x + threadIdx.x) * 16;
    if (x0 >= padw) return;
    const uint8_t *row = P + (long long)min(y, ph - 1) * pitch;
    uint8_t px[16];
    if (L.aligned16 && x0 + 16 <= pw) {
        const uint4 v = ldg128(row + x0);
        const unsigned w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int k = 0; k < 16; k++) px[k] = (uint8_t)(w[k >> 2] >> (8 * (k & 3)));
    } else {
#pragma unroll
        for (int k = 0; k < 16; k++) px[k] = row[min(x0 + k, pw - 1)];
    }
    unsigned w[4] = {0, 0, 0, 0};
#pragma unroll
    for (int k = 0; k < 16; k++) w[k >> 2] |= (unsigned)s_lut[px[k]] << (8 * (k & 3));
    *reinterpret_cast<uint4 *>(O + (long long)y * padw + x0) = make_uint4(w[0], w[1], w[2], w[3]);
}

}  // namespace h2j
