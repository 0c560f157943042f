// K2: pixels -> quantised zigzag coefficients, DC differences and symbol statistics.
//
// One thread owns one 8x8 block with all 64 values in registers.  A tile is 16 consecutive MCUs in coding
// order (it may wrap to the next MCU row):
//   warp 0: Y0/Y1 of the 16 MCUs  (32 horizontally adjacent blocks -> 256 contiguous bytes per pixel row)
//   warp 1: Y2/Y3
//   warp 2: Cb of the 16 MCUs (lanes 0-15), Cr (lanes 16-31)
// The first CTA of a frame also evaluates the rate control (K1's variance sum -> qscale) and publishes the
// quantiser tables; every CTA recomputes them for itself (64 threads, a few dozen instructions) instead of
// waiting for a separate set-up launch.
//
// DC prediction needs the previous MCU's Y3/Cb/Cr DC levels.  Inside a tile they come from shared memory; for
// the first MCU of a CTA's first tile they are recomputed from the pixels: the DC output of ff_fdct_sse2 is
// exactly the sum of the 64 samples (8 * column sums, then (8*S*16384 + 65536) >> 17 == S), so no second FDCT
// is needed.
#pragma once
#include "h2j_common.cuh"

namespace h2j {

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

// plane / position of block n (0..3 luma, 4 Cb, 5 Cr) of MCU m
struct BlockGeom {
    const uint8_t *P;
    int pitch, pw, ph, bx, by;
};
__device__ __forceinline__ BlockGeom block_geom(const uint8_t *base, const FrameLayout &L, int m, int n)
{
    BlockGeom g;
    const int my = m / L.mcu_w, mx = m - my * L.mcu_w;
    if (n < 4) {
        g.P = base; g.pitch = L.y_pitch; g.pw = L.w; g.ph = L.h;
        g.bx = mx * 16 + (n & 1) * 8; g.by = my * 16 + (n >> 1) * 8;
    } else {
        g.P = base + (n == 4 ? L.u_off : L.v_off); g.pitch = L.c_pitch; g.pw = L.cw; g.ph = L.ch;
        g.bx = mx * 8; g.by = my * 8;
    }
    return g;
}

__global__ void __launch_bounds__(kFdctThreads) fdct_quant_kernel(const uint8_t *__restrict__ frames, FrameLayout L,
                                                                  FrameState *__restrict__ state,
                                                                  const uint8_t *__restrict__ qscale_lut,
                                                                  FrameTab *__restrict__ tabs,
                                                                  uint32_t *__restrict__ images,          // [frame][images_cap] tile images
                                                                  unsigned long long *__restrict__ masks, // [frame][blocks_cap]
                                                                  long long images_cap, long long blocks_cap, int tiles_per_cta)
{
    __shared__ __align__(16) uint32_t s_img[kTileImageWords];
    __shared__ __align__(16) int s_q[64];
    __shared__ __align__(16) int s_bq[64];
    __shared__ unsigned int s_hist[2][256];
    __shared__ unsigned int s_dchist[2][16];
    __shared__ int s_dc[kTileBlocks];
    __shared__ unsigned long long s_mask[kTileBlocks];
    __shared__ int s_prev[3];        // DC levels of the MCU in front of the current tile: Y3, Cb, Cr
    __shared__ int s_psum[3];
    __shared__ int s_qs;

    const int f = blockIdx.y;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint8_t *base = frames + (long long)f * L.frame_stride;
    const int n_tiles = (L.n_mcu + kTileMcus - 1) / kTileMcus;
    const int tile0 = blockIdx.x * tiles_per_cta;
    if (tile0 >= n_tiles) return;

    // ---- rate control + quantiser set-up (ratecontrol.c first I picture, mpegvideo_enc.c encode_picture) ----
    if (tid == 0) {
        const long long var = (long long)state[f].var_sum;
        int q;
        if (L.fixed_qscale > 0) q = L.fixed_qscale;
        else {
            // predict_size(): the IEEE-exact part; the pow()/rounding tail is folded into qscale_lut by the host
            const double bits = __ddiv_rn(__dmul_rn(826.0, sqrt((double)var)), 236.0);
            int n = (int)bits;
            n = n > kQscaleLutSize - 1 ? kQscaleLutSize - 1 : (n < 0 ? 0 : n);
            q = qscale_lut[n];
        }
        s_qs = q;
        if (blockIdx.x == 0) {
            tabs[f].qscale = q;
            tabs[f].mb_var_sum = var;
            tabs[f].status = 0;
        }
    }
    if (tid < 3) s_psum[tid] = 0;
    for (int i = tid; i < 512; i += kFdctThreads) (&s_hist[0][0])[i] = 0;
    if (tid < 32) (&s_dchist[0][0])[tid] = 0;
    __syncthreads();
    if (tid < 64) {
        uint8_t m;
        uint32_t pk;
        quant_entry(s_qs, c_mpeg1_intra[tid], tid, &m, &pk);
        s_q[tid] = (int)(pk & 0xffffu);
        s_bq[tid] = (int)(pk >> 16);
        if (blockIdx.x == 0) {
            tabs[f].qpack[tid] = pk;
            tabs[f].intra[tid] = m;
            uint8_t mk;
            uint32_t pk2;
            quant_entry(s_qs, c_mpeg1_intra[c_zigzag[tid]], c_zigzag[tid], &mk, &pk2);  // DQT is stored in zigzag order
            tabs[f].dqt_zz[tid] = mk;
        }
    }
    // ---- predictor DCs in front of the CTA's first tile: pixel sums of Y3/Cb/Cr of the previous MCU ----
    if (tile0 > 0 && tid < 24) {
        const int which = tid >> 3, r = tid & 7;  // 0: Y3, 1: Cb, 2: Cr
        const BlockGeom g = block_geom(base, L, tile0 * kTileMcus - 1, which == 0 ? 3 : 3 + which);
        const uint8_t *lutp = L.range_mode ? c_range_lut[which ? 1 : 0] : nullptr;
        const uint8_t *row = g.P + (long long)min(g.by + r, g.ph - 1) * g.pitch;
        int s = 0;
#pragma unroll
        for (int c = 0; c < 8; c++) {
            const int p = row[min(g.bx + c, g.pw - 1)];
            s += lutp ? lutp[p] : p;
        }
        atomicAdd(&s_psum[which], s);
    }
    __syncthreads();
    if (tid < 3) s_prev[tid] = tile0 > 0 ? quant_dc(s_psum[tid]) : 128;  // 128 = the encoder's initial last_dc (1024 >> 3)
    // (visibility of s_prev is covered by the barrier in front of its first use below)

    const int mcu_l = warp < 2 ? (lane >> 1) : (lane & 15);
    const int n = warp < 2 ? (warp * 2 + (lane & 1)) : (4 + (lane >> 4));
    const int slot = mcu_l * 6 + n;
    const int cls = n < 4 ? 0 : 1;
    const uint8_t *lut = L.range_mode ? c_range_lut[cls] : nullptr;
    int16_t *img16 = reinterpret_cast<int16_t *>(s_img);

    for (int t = 0; t < tiles_per_cta; t++) {
        const int tile = tile0 + t;
        if (tile >= n_tiles) break;
        const int m = tile * kTileMcus + mcu_l;
        const bool valid = m < L.n_mcu;
        unsigned mask_lo = 0, mask_hi = 0;
        uint32_t word0_hi = 0;
        int dc = 0;
        if (valid) {
            const BlockGeom g = block_geom(base, L, m, n);
            int v[64];
            load_block_pixels(g.P, g.pitch, g.pw, g.ph, g.bx, g.by, L.aligned8 != 0, lut, v);
            fdct_8x8(v);
            dc = quant_dc(v[0]);
#pragma unroll
            for (int i = 1; i < 64; i++) v[i] = quant_ac2(v[i], s_q[i], s_bq[i]);
            // zigzag; word j of the record holds levels j (low half) and j + 32 (high half), so that the non-zero
            // flags of 32 levels fall out of 16 packed min(x, 1) results shifted into place (VIMNMX.U16x2)
            uint32_t *dst = s_img + slot * kBlkWords;
            unsigned fa = 0, fb = 0;
#pragma unroll
            for (int j = 0; j < 32; j++) {
                const uint32_t w = (uint32_t)(v[zz_of(j)] & 0xffff) | ((uint32_t)v[zz_of(j + 32)] << 16);
                const unsigned t = __vminu2(w, 0x00010001u);
                if (j < 16) fa += t << j;
                else fb += t << (j - 16);
                if (j == 0) word0_hi = w & 0xffff0000u;  // the low half becomes the DC difference below
                else dst[j] = w;
            }
            mask_lo = ((fa & 0xffffu) | (fb << 16)) & ~1u;  // bit 0 is the DC position: always coded, never in the mask
            mask_hi = (fa >> 16) | (fb & 0xffff0000u);
        }
        s_dc[slot] = dc;
        s_mask[slot] = ((unsigned long long)mask_hi << 32) | mask_lo;
        __syncthreads();

        if (valid) {
            // ---- DC difference to the previous block of the same component (mjpegenc.c encode_block) ----
            int pred;
            if (n >= 1 && n <= 3) pred = s_dc[slot - 1];
            else if (n == 0) pred = mcu_l > 0 ? s_dc[slot - 3] : s_prev[0];
            else pred = mcu_l > 0 ? s_dc[slot - 6] : s_prev[n - 3];
            const int diff = dc - pred;
            s_img[slot * kBlkWords] = (uint32_t)(diff & 0xffff) | word0_hi;
            atomicAdd(&s_dchist[cls][mag_bits(diff)], 1u);
            // ---- AC symbol statistics (ff_mjpeg_encode_coef / record_block, AC part) ----
            const int16_t *lv = img16 + slot * kBlkHalf;
            unsigned int *hist = s_hist[cls];
            int prev = 0;
#pragma unroll
            for (int half = 0; half < 2; half++) {
                unsigned mm = half ? mask_hi : mask_lo;
                while (mm) {
                    const int bp = __ffs((int)mm) - 1, k = half * 32 + bp;
                    mm &= mm - 1;
                    const int run = k - prev - 1;
                    prev = k;
                    const int nb = mag_bits((int)lv[2 * bp + half]);
                    if (run >= 16) atomicAdd(&hist[0xf0], (unsigned)(run >> 4));
                    atomicAdd(&hist[((run & 15) << 4) | nb], 1u);
                }
            }
            if (prev < 63) atomicAdd(&hist[0], 1u);
        }
        __syncthreads();

        // ---- contiguous copy-out: the tile image (128-bit stores) and the masks ----
        const int blocks_here = min(kTileBlocks, L.n_blocks - tile * kTileBlocks);
        {
            uint4 *gdst = reinterpret_cast<uint4 *>(images + ((long long)f * images_cap + tile) * kTileImageWords);
            const uint4 *ssrc = reinterpret_cast<const uint4 *>(s_img);
            const int n16 = (blocks_here * kBlkWords * 4 + 15) >> 4;
            for (int c = tid; c < n16; c += kFdctThreads) gdst[c] = ssrc[c];
        }
        if (tid < blocks_here) masks[(long long)f * blocks_cap + (long long)tile * kTileBlocks + tid] = s_mask[tid];
        // the last MCU of this tile predicts the first of the next
        if (tid < 3) s_prev[tid] = s_dc[(kTileMcus - 1) * 6 + 3 + tid];
        __syncthreads();
    }

    for (int i = tid; i < 512; i += kFdctThreads) {
        const unsigned c = (&s_hist[0][0])[i];
        if (c) atomicAdd(&state[f].hist[2 + (i >> 8)][i & 255], c);
    }
    if (tid < 32) {
        const unsigned c = (&s_dchist[0][0])[tid];
        if (c) atomicAdd(&state[f].hist[tid >> 4][tid & 15], c);
    }
}

}  // namespace h2j
