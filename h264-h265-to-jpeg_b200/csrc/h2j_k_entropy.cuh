// K4: entropy coding (mjpegenc.c record_block + ff_mjpeg_encode_picture_frame).
//
// A tile is 192 consecutive blocks of one frame (two K2 tile images).  One elected thread pulls the images (levels
// and non-zero masks) into shared memory with a single bulk copy (cp.async.bulk -> SASS UBLKCP) signalled on an
// mbarrier while the other threads fetch the frame's code tables and clear the bit window.
//   1. every thread walks the non-zero mask of its block and sums code lengths; block-wide exclusive scan.  The
//      tile's bit length is published at once, and warp 0 issues its look-back loads (tiles are handed out through
//      an atomic ticket, so every predecessor is already running).
//   2. threads emit their codes into the tile's shared-memory bit window.
//   3. warp 0 publishes the tile's trailing 31 bits, resolves the exclusive prefix from the loads issued in 1
//      (decoupled look-back; by now the predecessors have long published their lengths) and fetches the trailing
//      bits of the tile in front -- the only thing that depends on a neighbour's emission.
//   4. the window is shifted to the tile's global bit position and stored as big-endian words; a 32-bit word is
//      written by the tile that holds its last bit, with the bits of the tile in front taken from its published
//      tail -- no atomics on the scan, no pre-zeroed output.  While storing, the 0xFF bytes of every word are
//      counted into per-chunk counters for K5.
// The bit window holds 64 Kibit.  A tile that needs more (> 341 bits per block on average; the reference's own
// 2 MiB output cap is hit first at 1080p) takes the windowed path: the same emission clipped to one window of the
// tile's bit range at a time.
//
// Descriptors (64-bit words, so a single relaxed store/load carries everything), two per tile:
//   length word  [63:62] status (0 invalid, 1 tile length, 2 inclusive prefix length)   [31:0] bits
//   tail word    [63] valid   [30:0] the tile's last 31 bits (every tile but a frame's last is longer than that:
//                192 blocks of at least two bits each)
#pragma once
#include <cstddef>

#include "h2j_common.cuh"

namespace h2j {

__device__ __forceinline__ unsigned long long desc_pack(unsigned status, unsigned len) { return ((unsigned long long)status << 62) | len; }
__device__ __forceinline__ unsigned desc_status(unsigned long long d) { return (unsigned)(d >> 62); }
__device__ __forceinline__ unsigned desc_len(unsigned long long d) { return (unsigned)d; }
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

// ---- bit sinks ------------------------------------------------------------------------------------
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

struct BitSinkClip {  // same stream, but only the bits inside [lo, hi) are kept, at position (bit - lo)
    unsigned int *buf;
    unsigned pos, lo, hi;
    __device__ __forceinline__ void init(unsigned int *b, unsigned start, unsigned lo_, unsigned hi_) { buf = b; pos = start; lo = lo_; hi = hi_; }
    __device__ __forceinline__ void put(unsigned bits, int len)
    {
        unsigned s = pos;
        const unsigned e = pos + (unsigned)len;
        pos = e;
        if (e <= lo || s >= hi) return;
        if (s < lo) {
            len = (int)(e - lo);
            bits &= (1u << len) - 1u;
            s = lo;
        }
        if (e > hi) {
            const int cut = (int)(e - hi);
            bits >>= cut;
            len -= cut;
        }
        const unsigned r = s - lo;
        const unsigned long long v = (unsigned long long)bits << (64 - (int)(r & 31) - len);
        atomicOr(&buf[r >> 5], (unsigned)(v >> 32));
        if ((unsigned)v) atomicOr(&buf[(r >> 5) + 1], (unsigned)v);
    }
    __device__ __forceinline__ void flush() {}
};

// One pass over a block.  EMIT=false: returns the bit length.  EMIT=true: writes the bits into `sink`.
// cb = the block's 66-halfword record in shared memory ([0] = DC difference, level k at [2*(k&31) + (k>>5)]).
template <bool EMIT, typename Sink>
__device__ __forceinline__ unsigned walk_block(const int16_t *cb, unsigned mask_lo, unsigned mask_hi, const uint32_t *__restrict__ hdc,
                                               const uint32_t *__restrict__ hac, Sink *sink)
{
    unsigned total = 0;
    {
        const int diff = (int)cb[0];
        const int nb = mag_bits(diff);
        const uint32_t e = hdc[nb];
        const int sz = e & 31;
        if (EMIT) {
            const unsigned mant = (unsigned)(diff < 0 ? diff - 1 : diff) & ((1u << nb) - 1u);
            sink->put(((e >> 5) << nb) | mant, sz + nb);
        } else total += sz + nb;
    }
    int prev = 0;
    const uint32_t zrl = hac[0xf0];
#pragma unroll
    for (int half = 0; half < 2; half++) {
        unsigned mm = half ? mask_hi : mask_lo;
        while (mm) {
            const int bp = __ffs((int)mm) - 1, k = half * 32 + bp;
            mm &= mm - 1;
            int run = k - prev - 1;
            prev = k;
            const int val = (int)cb[2 * bp + half];
            const int nb = mag_bits(val);
            if (run >= 16) {
                if (EMIT) {
                    for (int z = run >> 4; z > 0; z--) sink->put(zrl >> 5, zrl & 31);
                } else total += (unsigned)(run >> 4) * (zrl & 31);
                run &= 15;
            }
            const uint32_t e = hac[(run << 4) | nb];
            const int sz = e & 31;
            if (EMIT) {
                const unsigned mant = (unsigned)(val < 0 ? val - 1 : val) & ((1u << nb) - 1u);
                sink->put(((e >> 5) << nb) | mant, sz + nb);
            } else total += sz + nb;
        }
    }
    if (prev < 63) {
        const uint32_t e = hac[0];
        if (EMIT) sink->put(e >> 5, e & 31);
        else total += e & 31;
    }
    return total;
}

// last min(len,31) bits of a stream of `len` bits held MSB-first in words[1..]; words[0] must be readable
__device__ __forceinline__ unsigned stream_tail(const unsigned int *words, unsigned len)
{
    if (len == 0) return 0;
    const unsigned endw = (len - 1) >> 5;        // word holding the last bit
    const unsigned used = ((len - 1) & 31) + 1;  // bits used in it
    const unsigned long long two = ((unsigned long long)(endw ? words[endw] : 0u) << 32) | words[endw + 1];
    const unsigned last32 = (unsigned)(two >> (32 - used));
    return len >= 31 ? (last32 & 0x7fffffffu) : (last32 & ((1u << len) - 1u));
}

// Phase 4 for a run of `len` bits that starts at global bit position P of frame f's scan: the bits are in
// s_bits[1..] (MSB first), s_bits[0] holds the bits in front of P (right aligned).  Words whose last bit lies in
// the run are stored; with `final_run` the frame's last, incomplete word is stored too, padded with ones up to the
// byte boundary (ff_mjpeg_encode_picture_trailer / put_bits padding).
__device__ __forceinline__ bool store_run(const unsigned int *s_bits, unsigned P, unsigned len, bool final_run, uint32_t *__restrict__ gs,
                                          long long scan_cap_words, unsigned int *__restrict__ chunk_ff, int tid)
{
    const unsigned s = P & 31;
    const long long W0 = P >> 5;
    const unsigned long long endbit = (unsigned long long)P + len;
    long long Wend = (long long)(endbit >> 5);
    const unsigned used = (unsigned)(endbit & 31);
    if (final_run && used) Wend++;
    bool overflow = false;
    for (long long W = W0 + tid; W < Wend; W += kEntThreads) {
        const int j = (int)(W - W0);
        const unsigned hi = s_bits[j], lo = s_bits[j + 1];  // local words j-1 and j
        unsigned v = s ? ((hi << (32 - s)) | (lo >> s)) : lo;
        if (final_run && used && W == Wend - 1) {
            const unsigned padn = (8 - (used & 7)) & 7;
            v |= ((1u << padn) - 1u) << (32 - used - padn);
        }
        if (W < scan_cap_words) {
            gs[W] = __byte_perm(v, 0, 0x0123);
            const unsigned c = count_ff_bytes(v);
            if (c) atomicAdd(&chunk_ff[W >> kChunkShift], c);
        } else overflow = true;
    }
    return overflow;
}

constexpr int kEntTabBytes = (2 * 16 + 2 * 256) * 4;  // DC luma/chroma (16 entries each), AC luma/chroma of FrameTab::hcode
constexpr int kEntSmemBytes = kEntFdctTiles * kTileImageBytes + kEntTabBytes + (kEntWinWords + 4) * 4;
static_assert(offsetof(FrameTab, hcode) % 16 == 0 && sizeof(FrameTab) % 16 == 0, "hcode must be bulk-copyable");
constexpr int kEntPreZeroWords = 768;  // cleared while the bulk load is in flight; covers 128 bits per block

__global__ void __launch_bounds__(kEntThreads) entropy_kernel(FrameLayout L, FrameTab *__restrict__ tabs, FrameState *__restrict__ state,
                                                              const uint32_t *__restrict__ images, long long images_cap,
                                                              unsigned long long *__restrict__ descs,  // [frame][2][tiles_per_frame]
                                                              unsigned int *__restrict__ ticket, int tiles_per_frame,
                                                              uint32_t *__restrict__ scan, long long scan_cap_words,
                                                              unsigned int *__restrict__ chunk_ff, int chunks_cap)
{
    extern __shared__ __align__(128) unsigned char ent_smem[];
    uint32_t *s_img = reinterpret_cast<uint32_t *>(ent_smem);                                   // two tile images
    uint32_t *s_hdc = reinterpret_cast<uint32_t *>(ent_smem + kEntFdctTiles * kTileImageBytes);  // [2][16] DC code tables
    uint32_t *s_hac = s_hdc + 32;                                                                // [2][256] AC code tables
    unsigned int *s_bits = reinterpret_cast<unsigned int *>(ent_smem + kEntFdctTiles * kTileImageBytes + kEntTabBytes);  // [0] guard, [1..] bits
    __shared__ unsigned s_warp[kEntThreads / 32];
    __shared__ unsigned s_ticket;
    __shared__ unsigned s_excl_len;
    __shared__ __align__(8) unsigned long long s_bar;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        s_ticket = atomicAdd(ticket, 1u);
        mbar_init(&s_bar, 1);
    }
    for (int i = tid; i < kEntPreZeroWords; i += kEntThreads) s_bits[i] = 0;  // while the ticket is on its way
    __syncthreads();
    const int f = (int)(s_ticket / (unsigned)tiles_per_frame);
    const int tile = (int)(s_ticket % (unsigned)tiles_per_frame);
    if (tid == 0) {
        // images_cap is even, so both images of the tile exist in the buffer even when the second holds no block
        const uint32_t *src = images + ((long long)f * images_cap + (long long)tile * kEntFdctTiles) * kTileImageWords;
        mbar_expect_tx(&s_bar, kEntFdctTiles * kTileImageBytes + kEntTabBytes);
        bulk_g2s(s_img, src, kEntFdctTiles * kTileImageBytes, &s_bar);
        bulk_g2s(s_hdc, tabs[f].hcode[0], 64, &s_bar);
        bulk_g2s(s_hdc + 16, tabs[f].hcode[1], 64, &s_bar);
        bulk_g2s(s_hac, tabs[f].hcode[2], 2048, &s_bar);
    }

    const int b = tile * kEntBlocks + tid;
    const bool valid = b < L.n_blocks;
    const int img_i = tid >= kTileBlocks ? 1 : 0, rec_i = tid - img_i * kTileBlocks;
    const uint32_t *rec = s_img + img_i * kTileImageWords + rec_i * kBlkWords;
    const int cls = (tid % 6) < 4 ? 0 : 1;  // 192 is a multiple of 6: the block's position in its MCU is tid % 6
    const int16_t *cb = reinterpret_cast<const int16_t *>(rec);
    mbar_wait(&s_bar, 0); // coefficient images and code tables landed
    unsigned mask_lo = 0, mask_hi = 0;
    if (valid) {
        mask_lo = rec[kMaskLoWord];
        mask_hi = s_img[img_i * kTileImageWords + kMaskHiOff + rec_i];
    }

    // ---- 1. lengths ----
    unsigned len = 0;
    if (valid) len = walk_block<false, BitSink>(cb, mask_lo, mask_hi, s_hdc + 16 * cls, s_hac + 256 * cls, nullptr);
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
    for (int w = 0; w < kEntThreads / 32; w++) {
        if (w < warp) warp_off += s_warp[w];
        tile_len += s_warp[w];
    }
    const unsigned off = warp_off + incl - len;
    const bool windowed = tile_len > (unsigned)kEntWinBits;

    // the length is all a successor's look-back needs: publish it now, and start our own look-back loads
    unsigned long long *D = descs + (long long)f * tiles_per_frame * 2;  // length words
    unsigned long long *TW = D + tiles_per_frame;                        // tail words
    unsigned long long d_look = 0;
    if (warp == 0) {
        if (lane == 0) st_desc(&D[tile], desc_pack(tile > 0 ? 1 : 2, tile_len));
        const int idx = tile - 1 - lane;
        d_look = idx >= 0 ? ld_desc(&D[idx]) : desc_pack(2, 0);
    }

    // ---- 2. bits of the tile (or, windowed, only its last 64 bits: enough for the tail word) ----
    unsigned own_tail;
    if (!windowed) {
        const int n_words = (int)((tile_len + 31) >> 5);
        if (n_words + 2 > kEntPreZeroWords) {  // uniform across the CTA
            for (int i = kEntPreZeroWords + tid; i <= n_words + 1; i += kEntThreads) s_bits[i] = 0;
            __syncthreads();
        }
        if (valid) {
            BitSink sink;
            sink.init(s_bits + 1, off);
            walk_block<true, BitSink>(cb, mask_lo, mask_hi, s_hdc + 16 * cls, s_hac + 256 * cls, &sink);
            sink.flush();
        }
        __syncthreads();
        own_tail = stream_tail(s_bits, tile_len);
    } else {
        const unsigned lo = tile_len - 64;
        if (valid && off + len > lo) {
            BitSinkClip sink;
            sink.init(s_bits + 1, off, lo, tile_len);
            walk_block<true, BitSinkClip>(cb, mask_lo, mask_hi, s_hdc + 16 * cls, s_hac + 256 * cls, &sink);
        }
        __syncthreads();
        own_tail = s_bits[2] & 0x7fffffffu;
    }

    // ---- 3. publish the tail, resolve the exclusive prefix, fetch the tail of the tile in front (warp 0) ----
    if (warp == 0) {
        if (lane == 0) st_desc(&TW[tile], (1ull << 63) | own_tail);
        unsigned excl = 0;
        if (tile > 0) {
            int basei = tile - 1;
            unsigned long long d = d_look;
            while (true) {
                const int idx = basei - lane;
                if (idx >= 0) {
                    while (desc_status(d) == 0) d = ld_desc(&D[idx]);
                } else d = desc_pack(2, 0);
                const unsigned pm = __ballot_sync(0xffffffffu, desc_status(d) == 2);
                const int stop = pm ? (__ffs(pm) - 1) : 31;
                unsigned part = lane <= stop ? desc_len(d) : 0u;
#pragma unroll
                for (int o = 16; o; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
                excl += part;
                if (pm) break;
                basei -= 32;
                const int idx2 = basei - lane;
                d = idx2 >= 0 ? ld_desc(&D[idx2]) : desc_pack(2, 0);
            }
        }
        if (lane == 0) {
            if (tile > 0) st_desc(&D[tile], desc_pack(2, excl + tile_len));
            s_excl_len = excl;
            if (tile == tiles_per_frame - 1) state[f].scan_bits = (unsigned long long)excl + tile_len;
        }
    }
    __syncthreads();

    // ---- 4. shift to the global bit position, store, count 0xFF bytes ----
    const unsigned P = s_excl_len;
    // Only the tile's first output word needs bits of the tile in front, and only thread 0 builds that word: it alone
    // waits for the neighbour's tail, everybody else is already storing.
    if (tid == 0) {
        unsigned tail_in = 0;
        if (tile > 0 && (P & 31)) {
            unsigned long long t;
            do { t = ld_desc(&TW[tile - 1]); } while ((t >> 63) == 0);
            tail_in = (unsigned)t & 0x7fffffffu;
        }
        s_bits[0] = tail_in;
    }
    uint32_t *gs = scan + (long long)f * scan_cap_words;
    unsigned int *cff = chunk_ff + (long long)f * chunks_cap;
    const bool last_tile = tile == tiles_per_frame - 1;
    bool overflow = false;
    if (!windowed) {
        overflow = store_run(s_bits, P, tile_len, last_tile, gs, scan_cap_words, cff, tid);
    } else {
        __syncthreads();
        unsigned tail_in = s_bits[0];
        for (unsigned lo = 0; lo < tile_len; lo += kEntWinBits) {
            const unsigned hi = min(lo + (unsigned)kEntWinBits, tile_len);
            __syncthreads();  // previous window fully stored (first round: everybody has read s_bits[0])
            for (int i = tid; i <= kEntWinWords + 2; i += kEntThreads) s_bits[i] = 0;
            __syncthreads();
            if (valid && off < hi && off + len > lo) {
                BitSinkClip sink;
                sink.init(s_bits + 1, off, lo, hi);
                walk_block<true, BitSinkClip>(cb, mask_lo, mask_hi, s_hdc + 16 * cls, s_hac + 256 * cls, &sink);
            }
            if (tid == 0) s_bits[0] = tail_in;
            __syncthreads();
            overflow |= store_run(s_bits, P + lo, hi - lo, last_tile && hi == tile_len, gs, scan_cap_words, cff, tid);
            tail_in = stream_tail(s_bits, hi - lo);  // a full window is longer than 31 bits, so it alone decides the tail
        }
    }
    if (overflow) tabs[f].status = -4;
}

}  // namespace h2j
