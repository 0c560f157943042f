// K2: pixels -> quantised zigzag coefficients, DC differences and symbol statistics.
//
// One thread owns one 8x8 block with all 64 values in registers.  A tile is 16 consecutive MCUs in coding order
// (it may wrap to the next MCU row); `tiles_per_cta` consecutive tiles are walked by three single-warp CTAs, one per
// role (blockIdx.x = tile group * 3 + role):
//   role 0: the four luma blocks of MCUs 0..7   (lane = mcu * 4 + n: lane order IS coding order)
//   role 1: the four luma blocks of MCUs 8..15
//   role 2: Cb of the 16 MCUs (lanes 0-15), Cr (lanes 16-31)
// A CTA is ONE warp because the warps share nothing: with three-warp CTAs the registers of a finished warp stayed
// allocated until the slowest of the three (luma and chroma blocks differ in their number of non-zero levels) reached
// the closing barrier -- 8.5 % of the stall samples -- and 128 registers x 96 threads fits five times per SM (15 warps),
// while 128 x 32 fits sixteen times.
// With this mapping every DC predictor (mjpegenc.c encode_block: the previous block of the same component) is
// the neighbouring lane's DC, so prediction is one shuffle and the warps never wait for each other.  The first
// lane of a run needs the block in front of the warp's range; its DC is recomputed from the pixels: the DC
// output of ff_fdct_sse2 is exactly the sum of the 64 samples (8 * column sums, then (8*S*16384 + 65536) >> 17
// == S), so eight lanes add one pixel row each -- no second FDCT.
//
// What leaves the kernel is the role's COMPACT sub-image of the tile (h2j_common.cuh): a header word per block and one
// entry per non-zero AC level.  The quantised levels are first written to a scratch record per lane in shared memory
// (registers cannot be indexed by a position found at run time); the statistics walk, which visits every non-zero level
// anyway to count its (run, size) symbol, appends the level's entry -- level, run, size -- to the lane's list in the
// staging buffer, at the offset a warp scan of the non-zero counts gave it.  The sub-image goes to the place its
// (tile, role) owns in the frame's region with ONE bulk (TMA) store of exactly its length, and the (tile, role) directory
// entry records that length.  A sub-image whose lists outgrow the staging buffer (more than 30 non-zero levels per block on
// average) leaves in two pieces, sixteen blocks each.
// Pixel rows of the next tile are requested before the statistics of the current one are taken, so their latency
// is covered by work.
//
// The frame's qscale and quantiser tables are set up once per frame by the last CTA of K1 (h2j_k_planes.cuh).
#pragma once
#include "h2j_common.cuh"

// (the walk takes two positions per trip from the low end.  K2 per 2048 frames, r1: two-ended 5.538 ms, 2 per trip 5.511,
// 3 per trip 5.592, 4 per trip 5.626 -- absent positions are predicated off)
#ifndef H2J_FDCT_MIN_CTAS
#define H2J_FDCT_MIN_CTAS 16  // resident single-warp CTAs per SM the register allocation is bounded for (16 -> 128 registers)
#endif

namespace h2j {

// What a thread needs to know about the plane its block lives in: fixed for the whole kernel.
struct PlaneRef {
    const uint8_t *P;
    int pitch, pw, ph;   // row pitch in bytes, width in BYTES / height in rows the encoder reads (edges are replicated beyond)
    int step, xstep;     // block spacing per MCU in rows (16 luma, 8 chroma) and in bytes (the same, but 16 for NV12 chroma)
    int xoff, yoff;      // block offset inside the MCU (bytes / rows)
    bool can_fast;       // rows start 8-byte aligned: 64-bit loads
};
// NV12 (template parameter of the kernel): the chroma plane holds Cb/Cr PAIRS, so an MCU's two chroma blocks are the 16
// bytes [16 mx, 16 mx + 16) of eight rows; the Cb lane fetches the first eight bytes of every row, the Cr lane (16 lanes
// up) the other eight, and they trade halves before the transform (nv12_trade).
template <bool NV12> __device__ __forceinline__ PlaneRef plane_ref(const uint8_t *base, const FrameLayout &L, int n)
{
    PlaneRef r;
    if (n < 4) {
        r.P = base; r.pitch = L.y_pitch; r.pw = L.w; r.ph = L.h; r.step = 16; r.xstep = 16;
        r.xoff = (n & 1) * 8; r.yoff = (n >> 1) * 8;
    } else if (NV12) {
        r.P = base + L.u_off; r.pitch = L.c_pitch; r.pw = 2 * L.cw; r.ph = L.ch; r.step = 8; r.xstep = 16;
        r.xoff = (n - 4) * 8; r.yoff = 0;
    } else {
        r.P = base + (n == 4 ? L.u_off : L.v_off); r.pitch = L.c_pitch; r.pw = L.cw; r.ph = L.ch; r.step = 8; r.xstep = 8;
        r.xoff = 0; r.yoff = 0;
    }
    r.can_fast = L.aligned8 != 0;
    return r;
}

// Top-left sample of a thread's block in its plane plus the MCU index; advances a tile (16 MCUs) at a time without
// dividing.  mx = -1, my = 0 stands for "the MCU in front of MCU 0" (never dereferenced) and advances correctly.
struct BlockPos {
    int m, bx, by;
    __device__ __forceinline__ void init(int m_, const PlaneRef &R, int mcu_w)
    {
        m = m_;
        const int my = m_ / mcu_w, mx = m_ - my * mcu_w;  // m_ = -1: my = 0, mx = -1
        bx = mx * R.xstep + R.xoff;
        by = my * R.step + R.yoff;
    }
    __device__ __forceinline__ void advance(const PlaneRef &R, int mcu_w)
    {
        const int row_w = mcu_w * R.xstep;  // bx = mx * xstep + xoff with xoff < xstep: bx >= row_w <=> mx >= mcu_w
        m += kTileMcus;
        bx += kTileMcus * R.xstep;
        while (bx >= row_w) { bx -= row_w; by += R.step; }
    }
};

// The pixel rows a thread has in flight for its block of the coming tile.  The loads are ALWAYS issued and ALWAYS unpacked
// (blocks that do not exist, unaligned rows and rows cut by the right edge read the frame's first bytes instead and are
// then re-read bytewise): with loads and consumers under matching branches, ptxas could not tell that the registers were
// free again, and the first instruction touching them waited on a scoreboard shared with the load issued just before
// (5 % of the stall samples); flags kept across the transform were spilled to local memory (8 %).
struct BlockFetch {
    uint2 rows[8];
    uint2 prow;          // lanes that help with a predecessor DC: one row of that block
};
// C = the block is an NV12 chroma block: the vector path needs the MCU's whole 16 bytes inside the row (both lanes of a
// Cb/Cr pair take the same path, they trade halves), and sample c of the block is byte 2 * (x + c) + component.
template <bool C> __device__ __forceinline__ bool fetch_is_fast(const PlaneRef &R, int bx)
{
    return R.can_fast && (C ? bx - R.xoff + 16 : bx + 8) <= R.pw;
}
template <bool C> __device__ __forceinline__ int sample_at(const uint8_t *row, const PlaneRef &R, int bx, int c)
{
    if (C) return row[2 * min(((bx - R.xoff) >> 1) + c, (R.pw >> 1) - 1) + (R.xoff >> 3)];
    return row[min(bx + c, R.pw - 1)];
}

template <bool C, bool CQ>
__device__ __forceinline__ void fetch_issue(BlockFetch &F, const uint8_t *safe, const PlaneRef &R, const BlockPos &bp, bool valid,
                                            const PlaneRef &Q, const BlockPos &pp, bool phelp, int prow_idx)
{
    const bool pf = !CQ && phelp && fetch_is_fast<false>(Q, pp.bx);  // (NV12 chroma: the predecessor rows are read bytewise)
    const bool rf = valid && fetch_is_fast<C>(R, bp.bx);
    const uint8_t *pq = pf ? Q.P + (long long)min(pp.by + prow_idx, Q.ph - 1) * Q.pitch + pp.bx : safe;
    F.prow = ldg64(pq);
    if (!rf || bp.by + 8 <= R.ph) {  // interior: one address, then a pitch per row
        const uint8_t *p = rf ? R.P + (long long)bp.by * R.pitch + bp.bx : safe;
        const int pitch = rf ? R.pitch : 0;
#pragma unroll
        for (int r = 0; r < 8; r++) {
            F.rows[r] = ldg64(p);
            p += pitch;
        }
    } else {
#pragma unroll
        for (int r = 0; r < 8; r++) F.rows[r] = ldg64(R.P + (long long)min(bp.by + r, R.ph - 1) * R.pitch + bp.bx);
    }
}

// NV12 chroma warp, all 32 lanes: every row arrives as four Cb/Cr pairs; the Cb lane keeps the four Cb samples and hands
// the four Cr samples to the Cr lane of its MCU (16 lanes up), which hands back the Cb samples of ITS four pairs.
__device__ __forceinline__ void nv12_trade(BlockFetch &F, bool cr_lane)
{
#pragma unroll
    for (int r = 0; r < 8; r++) {
        const unsigned cb4 = __byte_perm(F.rows[r].x, F.rows[r].y, 0x6420), cr4 = __byte_perm(F.rows[r].x, F.rows[r].y, 0x7531);
        const unsigned keep = cr_lane ? cr4 : cb4;
        const unsigned got = __shfl_xor_sync(0xffffffffu, cr_lane ? cb4 : cr4, 16);
        F.rows[r] = cr_lane ? make_uint2(got, keep) : make_uint2(keep, got);  // Cb lane had pairs 0-3, Cr lane pairs 4-7
    }
}

template <bool C>
__device__ __forceinline__ void fetch_consume(const BlockFetch &F, const PlaneRef &R, const BlockPos &bp, const uint8_t *lut, int (&v)[64])
{
#pragma unroll
    for (int r = 0; r < 8; r++) {
        // one byte permute per sample (selector 4 = a zero byte of the second operand)
        v[r * 8 + 0] = (int)__byte_perm(F.rows[r].x, 0u, 0x4440);
        v[r * 8 + 1] = (int)__byte_perm(F.rows[r].x, 0u, 0x4441);
        v[r * 8 + 2] = (int)__byte_perm(F.rows[r].x, 0u, 0x4442);
        v[r * 8 + 3] = (int)__byte_perm(F.rows[r].x, 0u, 0x4443);
        v[r * 8 + 4] = (int)__byte_perm(F.rows[r].y, 0u, 0x4440);
        v[r * 8 + 5] = (int)__byte_perm(F.rows[r].y, 0u, 0x4441);
        v[r * 8 + 6] = (int)__byte_perm(F.rows[r].y, 0u, 0x4442);
        v[r * 8 + 7] = (int)__byte_perm(F.rows[r].y, 0u, 0x4443);
    }
    if (!fetch_is_fast<C>(R, bp.bx)) {
#pragma unroll
        for (int r = 0; r < 8; r++) {
            const uint8_t *row = R.P + (long long)min(bp.by + r, R.ph - 1) * R.pitch;
#pragma unroll
            for (int c = 0; c < 8; c++) v[r * 8 + c] = sample_at<C>(row, R, bp.bx, c);
        }
    }
    if (lut) {
#pragma unroll
        for (int i = 0; i < 64; i++) v[i] = lut[v[i]];
    }
}

// this lane's share (one pixel row) of the predecessor block's sample sum
template <bool CQ>
__device__ __forceinline__ int fetch_pred_rowsum(const BlockFetch &F, const PlaneRef &Q, const BlockPos &pp, bool phelp, int prow_idx,
                                                 const uint8_t *lut)
{
    int s = (int)__dp4a(F.prow.y, 0x01010101u, __dp4a(F.prow.x, 0x01010101u, 0u));
    if (!phelp) return 0;
    if (CQ || !fetch_is_fast<false>(Q, pp.bx) || lut) {
        const uint8_t *row = Q.P + (long long)min(pp.by + prow_idx, Q.ph - 1) * Q.pitch;
        s = 0;
#pragma unroll
        for (int c = 0; c < 8; c++) {
            const int p = sample_at<CQ>(row, Q, pp.bx, c);
            s += lut ? lut[p] : p;
        }
    }
    return s;
}

constexpr int kFdctStageWords = 1088;  // staging buffer of a warp: 32 headers + up to 1056 entries (4.25 KiB); half a sub-image
                                       // at its densest (16 blocks x 63 entries + the headers) always fits
static_assert(kSubHdrWords + 16 * 63 <= kFdctStageWords, "half a sub-image must fit the staging buffer");

template <bool NV12>
__global__ void __launch_bounds__(kFdctThreads, H2J_FDCT_MIN_CTAS) fdct_quant_kernel(const uint8_t *__restrict__ frames, FrameLayout L,
                                                                     FrameState *__restrict__ state,
                                                                     const FrameTab *__restrict__ tabs,
                                                                     uint32_t *__restrict__ images,  // [frame][img_words_cap] coefficient regions
                                                                     long long img_words_cap,
                                                                     unsigned *__restrict__ dir,     // [frame][images_cap][kDirPerTile] words of each sub-image
                                                                     long long images_cap, int tiles_per_cta)
{
    __shared__ __align__(128) uint32_t s_out[kFdctStageWords];
    __shared__ uint32_t s_rec[kSubRecs * kBlkWords];
    __shared__ __align__(16) int s_q[64];
    __shared__ __align__(16) int s_bq[64];
    __shared__ unsigned int s_hist[256];
    __shared__ unsigned int s_dchist[16];

    const int f = blockIdx.y;
    const int lane = threadIdx.x;
    const int group = blockIdx.x / 3, warp = blockIdx.x - group * 3;  // `warp` = the role of this single-warp CTA
    const uint8_t *base = frames + (long long)f * L.frame_stride;
    const int n_tiles = (L.n_mcu + kTileMcus - 1) / kTileMcus;
    const int tile0 = group * tiles_per_cta;
    if (tile0 >= n_tiles) return;
    const int tile_end = min(tile0 + tiles_per_cta, n_tiles);

    // lane -> block of the tile
    const bool luma = warp < 2;
    const int mcu_l = luma ? warp * 8 + (lane >> 2) : (lane & 15);
    const int n = luma ? (lane & 3) : 4 + (lane >> 4);
    const int cls = luma ? 0 : 1;
    const uint8_t *lut = L.range_mode ? c_range_lut[cls] : nullptr;
    // predecessor block of the warp's first lane(s): luma -> Y3 of the MCU in front of the warp's eight;
    // chroma -> Cb (lanes 0-7 help) and Cr (lanes 8-15 help) of the MCU in front of the tile (first tile only)
    const bool phelp_lane = luma ? lane < 8 : lane < 16;
    const int mcu_first = luma ? warp * 8 : 0;
    const PlaneRef R = plane_ref<NV12>(base, L, n), Q = plane_ref<NV12>(base, L, luma ? 3 : 4 + (lane >> 3));
    const bool nvc = NV12 && !luma;  // this warp's blocks are NV12 chroma blocks (CTA-uniform)

    BlockFetch F;
    const uint8_t *safe = reinterpret_cast<const uint8_t *>(tabs);  // aligned, always readable: what skipped loads read
    BlockPos bp, pp;  // this thread's block / the predecessor block in front of the warp's range, for the tile being fetched
    bp.init(tile0 * kTileMcus + mcu_l, R, L.mcu_w);
    pp.init(tile0 * kTileMcus + mcu_first - 1, Q, L.mcu_w);
    // lanes that add a row of the predecessor block: chroma only needs it for its first tile (then the DC is carried)
    auto phelp_at = [&](int tile) { return phelp_lane && pp.m >= 0 && pp.m < L.n_mcu && (luma || tile == tile0); };
    if (nvc) fetch_issue<true, true>(F, safe, R, bp, bp.m < L.n_mcu, Q, pp, phelp_at(tile0), lane & 7);
    else fetch_issue<false, false>(F, safe, R, bp, bp.m < L.n_mcu, Q, pp, phelp_at(tile0), lane & 7);

    // ---- the frame's quantiser (set up once per frame by K1's last CTA), cleared statistics ----
    {
        const uint2 pk = reinterpret_cast<const uint2 *>(tabs[f].qpack)[lane];
        s_q[2 * lane] = (int)(pk.x & 0xffffu);
        s_bq[2 * lane] = (int)(pk.x >> 16);
        s_q[2 * lane + 1] = (int)(pk.y & 0xffffu);
        s_bq[2 * lane + 1] = (int)(pk.y >> 16);
    }
#pragma unroll
    for (int i = 0; i < 8; i++) s_hist[i * 32 + lane] = 0;
    if (lane < 16) s_dchist[lane] = 0;
    __syncwarp();

    int chroma_carry = 128;  // chroma warp, lanes 0 / 16: DC of the previous tile's last Cb / Cr block
    uint32_t *gimg = images + (long long)f * img_words_cap;

    for (int tile = tile0; tile < tile_end; tile++) {
        const bool valid = bp.m < L.n_mcu;

        // ---- predecessor DC for the first lane(s) of the warp, from pixel sums ----
        int psum = nvc ? fetch_pred_rowsum<true>(F, Q, pp, phelp_at(tile), lane & 7, lut) : fetch_pred_rowsum<false>(F, Q, pp, phelp_at(tile), lane & 7, lut);
        if (nvc) nv12_trade(F, lane >= 16);  // (all 32 lanes, outside `valid`: the lanes of a pair share their MCU)
        psum += __shfl_xor_sync(0xffffffffu, psum, 1);
        psum += __shfl_xor_sync(0xffffffffu, psum, 2);
        psum += __shfl_xor_sync(0xffffffffu, psum, 4);
        const int psum_cr = __shfl_sync(0xffffffffu, psum, 8);  // chroma: lanes 8-15 summed Cr
        const int pm = tile * kTileMcus + mcu_first - 1;
        int pred_first;  // meaningful in lane 0 (and lane 16 of the chroma warp)
        if (luma) pred_first = pm >= 0 ? quant_dc(psum) : 128;  // 128 = the encoder's initial last_dc (1024 >> 3)
        else if (tile == tile0) pred_first = pm >= 0 ? quant_dc(lane < 16 ? psum : psum_cr) : 128;
        else pred_first = chroma_carry;

        unsigned mask_lo = 0, mask_hi = 0;
        int dc = 0;
        uint32_t *rec = s_rec + lane * kBlkWords;  // scratch record (this lane writes it, this lane reads it back)
        if (valid) {
            int v[64];
            if (nvc) fetch_consume<true>(F, R, bp, lut, v);
            else fetch_consume<false>(F, R, bp, lut, v);
            fdct_8x8(v);
            dc = quant_dc(v[0]);
            // quantise without the final >> 16: the level is the upper half of the 32-bit product, so two of them
            // are packed into one record word with a single byte permute
#pragma unroll
            for (int i = 1; i < 64; i++) v[i] = quant_ac2_hi(v[i], s_q[i], s_bq[i]);
            // zigzag; word j of the record holds levels j (low half) and j + 32 (high half), so that the non-zero
            // flags of 32 levels fall out of 16 packed min(x, 1) results shifted into place (VIMNMX.U16x2)
            unsigned fa = 0, fb = 0;
#pragma unroll
            for (int j = 0; j < 32; j++) {
                const uint32_t w = j == 0 ? __byte_perm(0u, (uint32_t)v[zz_of(32)], 0x7610) : __byte_perm((uint32_t)v[zz_of(j)], (uint32_t)v[zz_of(j + 32)], 0x7632);
                const unsigned t = __vminu2(w, 0x00010001u);
                if (j < 16) fa += t << j;
                else fb += t << (j - 16);
                rec[j] = w;  // (word 0: low half unused, high half = level 32)
            }
            mask_lo = ((fa & 0xffffu) | (fb << 16)) & ~1u;  // bit 0 is the DC position: always coded, never in the mask
            mask_hi = (fa >> 16) | (fb & 0xffff0000u);
        }

        // ---- request the next tile's pixels: nothing of this tile's 64-value block is live any more ----
        if (tile + 1 < tile_end) {
            bp.advance(R, L.mcu_w);
            pp.advance(Q, L.mcu_w);
            if (nvc) fetch_issue<true, true>(F, safe, R, bp, bp.m < L.n_mcu, Q, pp, phelp_at(tile + 1), lane & 7);
            else fetch_issue<false, false>(F, safe, R, bp, bp.m < L.n_mcu, Q, pp, phelp_at(tile + 1), lane & 7);
        }

        // ---- where the lists go: a warp scan of the non-zero counts ----
        const int cnt = __popc(mask_lo) + __popc(mask_hi);  // (0 for blocks that do not exist)
        int incl = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        const int total = __shfl_sync(0xffffffffu, incl, 31);
        // the (tile, role)'s own place in the frame's region.  (A bump allocator -- one atomicAdd per sub-image, for densely
        // packed regions -- measured slower: all warps working on a frame queue up on one address, and lane 0 then waits for
        // the answer when it issues the store: 5 % of the kernel's stall samples.)
        uint32_t *gsub = gimg + ((long long)tile * kTileRoles + warp) * kSubMaxWords;

        // ---- DC difference to the previous block of the same component (mjpegenc.c encode_block) ----
        const int up = __shfl_up_sync(0xffffffffu, dc, 1);
        const int pred = (luma ? lane == 0 : (lane & 15) == 0) ? pred_first : up;
        const int last = __shfl_sync(0xffffffffu, dc, (lane & 16) | 15);  // chroma: this tile's last Cb / Cr
        chroma_carry = last;
        const int diff = dc - pred;
        if (valid) {
            atomicAdd(&s_dchist[mag_bits(diff)], 1u);
            if (!(mask_hi >> 31)) atomicAdd(&s_hist[0], 1u);  // position 63 is zero: the block ends with an EOB
        }

        // The block's entry list (ff_mjpeg_encode_coef / record_block, AC part): one entry per non-zero level.  The
        // (run, size) symbol of a non-zero level needs only its position, the position of the non-zero level below it and
        // its value.  Positions are taken in ascending order, two at a time: two independent bit-scan -> load -> size
        // chains per iteration instead of one, half the trips.
        auto walk_into = [&](uint32_t *dst) {
            const int16_t *lv = reinterpret_cast<const int16_t *>(rec);
            uint32_t *dst_hi = dst + __popc(mask_lo);
            auto put = [&](uint32_t *at, int k, int below, int val) {  // val != 0
                const int run = k - below - 1;
                unsigned top;  // size - 1
                asm("bfind.u32 %0, %1;" : "=r"(top) : "r"(abs(val)));
                *at = ((unsigned)val << 16) | (unsigned)((run << 4) + (int)top);  // entry: level | run << 4 | (size - 1)
            };
            unsigned lo = mask_lo;
            int below = 0;
            while (lo) {
                const unsigned b0 = lo & (0u - lo);
                lo ^= b0;
                const unsigned b1 = lo & (0u - lo);
                lo ^= b1;
                const int k0 = 31 - __clz(b0), k1 = 31 - __clz(b1);  // k1 = -1: absent (then this was the last trip)
                const int v0 = (int)lv[2 * k0], v1 = (int)lv[2 * max(k1, 0)];
                put(dst, k0, below, v0);
                if (b1) put(dst + 1, k1, k0, v1);
                dst += 2;
                below = k1;
            }
            int prev = mask_lo ? 31 - __clz(mask_lo) : 0;  // highest non-zero position below 32 (0: none but the DC)
            unsigned hi = mask_hi;
            while (hi) {
                const int bp = __ffs((int)hi) - 1, k = 32 + bp;
                hi &= hi - 1;
                put(dst_hi++, k, prev, (int)lv[2 * bp + 1]);
                prev = k;
            }
        };
        // Symbol statistics of finished lists, a lane per entry (every lane busy, unlike in the walk): (run & 15, size)
        // counts and 16-zero runs (symbol 0xF0)
        auto count_entries = [&](int e0, int e1) {
#pragma unroll 1
            for (int i = e0 + lane; i < e1; i += 32) {
                const unsigned e = s_out[i];
                atomicAdd(&s_hist[1 + (e & 0xffu)], 1u);  // (the symbol's size nibble is the entry's + 1: rides in the address)
                if (e & 0x300u) atomicAdd(&s_hist[0xf0], (e >> 8) & 3u);  // runs of 16 and more are rare
            }
        };

        // the store of the previous tile must have read the staging buffer out (lane 0 issued it, lane 0 waits; it had a
        // whole transform's time to do so)
        if (lane == 0) bulk_wait_read_all();
        __syncwarp();
        int words;
        if (kSubHdrWords + total <= kFdctStageWords) {
            // ---- the usual case: the whole sub-image is assembled in the staging buffer and leaves with one bulk store ----
            const int first = incl - cnt;
            // header: everything in it is known before the walk (an EOB is coded unless position 63 is non-zero)
            s_out[lane] = valid ? sub_hdr_pack(diff, (int)(~mask_hi >> 31), cnt, first) : 0u;
            if (valid) walk_into(s_out + kSubHdrWords + first);
            __syncwarp();
            count_entries(kSubHdrWords, kSubHdrWords + total);
            words = kSubHdrWords + total;
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) bulk_s2g(gsub, s_out, (uint32_t)((words + 3) & ~3) * 4u);
        } else {
            // ---- dense content (more than 30 non-zero levels per block on average): two pieces, lanes 0-15 (with the
            //      headers), then lanes 16-31, whose lists start on the next 16-byte boundary (the bulk copy's granule);
            //      half a sub-image always fits the buffer ----
            const int half_a = __shfl_sync(0xffffffffu, incl, 15);  // entries of lanes 0-15
            const int split = (half_a + 3) & ~3;                    // where the second piece's entries start
            const int first = incl - cnt + (lane >= 16 ? split - half_a : 0);
            s_out[lane] = valid ? sub_hdr_pack(diff, (int)(~mask_hi >> 31), cnt, first) : 0u;
            if (valid && lane < 16) walk_into(s_out + kSubHdrWords + first);
            __syncwarp();
            count_entries(kSubHdrWords, kSubHdrWords + half_a);
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
                bulk_s2g(gsub, s_out, (uint32_t)(kSubHdrWords + split) * 4u);
                bulk_wait_read_all();  // the second piece reuses the buffer
            }
            __syncwarp();
            if (valid && lane >= 16) walk_into(s_out + (first - split));
            __syncwarp();
            count_entries(0, total - half_a);
            words = kSubHdrWords + split + (total - half_a);
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) bulk_s2g(gsub + kSubHdrWords + split, s_out, (uint32_t)((total - half_a + 3) & ~3) * 4u);
        }
        if (lane == 0) dir[((long long)f * images_cap + tile) * kDirPerTile + warp] = (unsigned)words;
    }
    // the stores only have to be done READING shared memory before the CTA retires; they complete on their own and the
    // kernel boundary orders them before K4a
    if (lane == 0) bulk_wait_read_all();
    __syncwarp();

#pragma unroll
    for (int i = 0; i < 8; i++) {
        const unsigned c = s_hist[i * 32 + lane];
        if (c) atomicAdd(&state[f].hist[2 + cls][i * 32 + lane], c);
    }
    if (lane < 16) {
        const unsigned c = s_dchist[lane];
        if (c) atomicAdd(&state[f].hist[cls][lane], c);
    }
}

}  // namespace h2j
