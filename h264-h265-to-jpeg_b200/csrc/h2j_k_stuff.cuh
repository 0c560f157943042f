// K5: ff_mjpeg_escape_FF + picture trailer.  The scan of a frame is cut into chunks of 1024 words; K4 left the
// number of 0xFF bytes of every chunk in chunk_ff, so a chunk's output position is known from a short sum and
// chunks are independent: grid (ctas per frame, frames), each CTA strides over the frame's chunks.  A chunk is
// expanded into shared memory (every byte, and a 0x00 after each 0xFF) and leaves as aligned 32-bit stores.
// The CTA that owns the last chunk appends EOI and publishes the frame's size.
#pragma once
#include "h2j_common.cuh"

namespace h2j {

__global__ void __launch_bounds__(kStuffThreads) stuff_kernel(FrameTab *__restrict__ tabs, const FrameState *__restrict__ state,
                                                              const uint32_t *__restrict__ scan, long long scan_cap_words,
                                                              const unsigned int *__restrict__ chunk_ff, int chunks_cap,
                                                              uint8_t *__restrict__ out, long long out_cap)
{
    __shared__ __align__(16) uint8_t s_out[2 * kChunkWords * 4 + 32];
    __shared__ unsigned s_warp[kStuffThreads / 32];
    __shared__ unsigned s_red[kStuffThreads / 32];

    const int f = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    FrameTab *T = tabs + f;
    const long long bits = (long long)state[f].scan_bits;
    const long long nbytes = (bits + 7) >> 3;
    const long long nwords = (nbytes + 3) >> 2;
    const int nchunks = (int)((nwords + kChunkWords - 1) >> kChunkShift);
    const long long hdr = T->header_bytes;
    const uint32_t *gs = scan + (long long)f * scan_cap_words;
    const unsigned int *cff = chunk_ff + (long long)f * chunks_cap;
    uint8_t *o = out + (long long)f * out_cap;
    const uint32_t *s_out_w = reinterpret_cast<const uint32_t *>(s_out);

    for (int c = blockIdx.x; c < nchunks; c += gridDim.x) {
        // ---- 0xFF bytes in front of this chunk ----
        unsigned part = 0;
        for (int i = tid; i < c; i += kStuffThreads) part += cff[i];
#pragma unroll
        for (int ofs = 16; ofs; ofs >>= 1) part += __shfl_xor_sync(0xffffffffu, part, ofs);
        if (lane == 0) s_red[warp] = part;
        // ---- this thread's 4 words ----
        const long long w0 = (long long)c * kChunkWords + tid * 4;
        uint4 q = make_uint4(0, 0, 0, 0);
        if (w0 < nwords && w0 + 4 <= scan_cap_words) q = *reinterpret_cast<const uint4 *>(gs + w0);
        unsigned wv[4] = {q.x, q.y, q.z, q.w};  // memory byte order == stream order: byte j is (wv[j >> 2] >> (8 * (j & 3))) & 0xff
        long long rem = nbytes - w0 * 4;  // valid bytes from here on
        const int nb = rem >= 16 ? 16 : (rem > 0 ? (int)rem : 0);
        if (nb < 16) {  // the frame's last bytes: everything behind them counts as (and is copied as) zero
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const int left = nb - 4 * k;
                if (left <= 0) wv[k] = 0;
                else if (left < 4) wv[k] &= (1u << (8 * left)) - 1u;
            }
        }
        const unsigned cnt = count_ff_bytes(wv[0]) + count_ff_bytes(wv[1]) + count_ff_bytes(wv[2]) + count_ff_bytes(wv[3]);
        unsigned incl = cnt;
#pragma unroll
        for (int ofs = 1; ofs < 32; ofs <<= 1) {
            const unsigned t = __shfl_up_sync(0xffffffffu, incl, ofs);
            if (lane >= ofs) incl += t;
        }
        if (lane == 31) s_warp[warp] = incl;
        __syncthreads();
        unsigned woff = 0, chunk_total = 0, before = 0;
#pragma unroll
        for (int k = 0; k < kStuffThreads / 32; k++) {
            if (k < warp) woff += s_warp[k];
            chunk_total += s_warp[k];
            before += s_red[k];
        }
        // ---- expand into shared memory ----
        unsigned p = (unsigned)tid * 16 + woff + incl - cnt;
        if (cnt == 0 && nb == 16) {
            // no 0xFF among this thread's 16 bytes (19 threads in 20): they move as a block.  The three aligned words inside
            // [p, p + 16) are whole ours; the 4 bytes at the ragged ends share their words with the neighbours.
            const unsigned a = p & 3;
            uint32_t *sw = reinterpret_cast<uint32_t *>(s_out + (p & ~3u));
            if (a == 0) {
                sw[0] = wv[0]; sw[1] = wv[1]; sw[2] = wv[2]; sw[3] = wv[3];
            } else {
                const unsigned sh = a * 8;
                sw[1] = __funnelshift_l(wv[0], wv[1], sh);  // bytes 4-a .. 8-a of ours
                sw[2] = __funnelshift_l(wv[1], wv[2], sh);
                sw[3] = __funnelshift_l(wv[2], wv[3], sh);
                for (unsigned j = 0; j < 4 - a; j++) s_out[p + j] = (uint8_t)(wv[0] >> (8 * j));
                for (unsigned j = 0; j < a; j++) s_out[p + 16 - a + j] = (uint8_t)(wv[3] >> (8 * (4 - a + j)));
            }
        } else {
#pragma unroll
            for (int j = 0; j < 16; j++) {
                if (j < nb) {
                    const unsigned byte = (wv[j >> 2] >> (8 * (j & 3))) & 0xff;
                    s_out[p++] = (uint8_t)byte;
                    if (byte == 0xff) s_out[p++] = 0;
                }
            }
        }
        __syncthreads();
        // ---- copy out: dst is byte aligned at best, so head bytes, aligned words, tail bytes ----
        const long long chunk_bytes_in = min((long long)kChunkWords * 4, nbytes - (long long)c * kChunkWords * 4);
        const long long total = chunk_bytes_in + chunk_total;
        const long long dst0 = hdr + (long long)c * kChunkWords * 4 + before;
        const int head = (int)min(total, (long long)((4 - (dst0 & 3)) & 3));
        if (tid < head && dst0 + tid < out_cap) o[dst0 + tid] = s_out[tid];
        const long long nw = (total - head) >> 2;
        uint32_t *ow = reinterpret_cast<uint32_t *>(o + dst0 + head);
        for (long long i = tid; i < nw; i += kStuffThreads) {
            const int s = head + (int)i * 4;
            const uint32_t lo = s_out_w[s >> 2], hi = s_out_w[(s >> 2) + 1];
            if (dst0 + head + i * 4 + 4 <= out_cap) ow[i] = __funnelshift_r(lo, hi, (s & 3) * 8);
        }
        const int tail = (int)((total - head) & 3);
        if (tid < tail) {
            const long long at = dst0 + head + nw * 4 + tid;
            if (at < out_cap) o[at] = s_out[head + nw * 4 + tid];
        }
        if (c == nchunks - 1 && tid == 0) {
            const long long ff = (long long)before + chunk_total;
            const long long end = hdr + nbytes + ff;
            if (end + 2 <= out_cap) { o[end] = 0xff; o[end + 1] = 0xd9; }
            else T->status = -4;
            T->scan_bits = bits;
            T->stuffed_ff = ff;
            T->jpeg_bytes = end + 2;
        }
        __syncthreads();
    }
}

}  // namespace h2j
