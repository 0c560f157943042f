// K3: optimal Huffman tables (mjpegenc_huffman.c) + code tables + JPEG header (mjpegenc_common.c).
// One CTA of 128 threads per table (DC luma, DC chroma, AC luma, AC chroma) and frame.
//
// What is sequential and why: AV_QSORT is not stable and the order of equal counts / equal lengths decides the
// DHT bytes, so both sorts are replayed step for step by one thread of the group (in shared memory).
// What is parallel: the package-merge.  Level t of ff_mjpegenc_huffman_compute_bits is a merge of two sorted
// lists -- the symbols and the pairwise sums ("packages") of level t-1 -- with packages winning ties.  Every
// item finds its rank in the other list with a binary search, so a level costs a handful of dependent
// shared-memory reads instead of ~500 serial steps.  Per level only the item probabilities and, for every
// prefix of the merged list, the number of symbols in it are kept; that is enough to recover the code lengths
// (symbols are consumed in sorted order, packages in pairs) and gives the same answer as the reference's
// list-copying formulation.
#pragma once
#include "h2j_common.cuh"

namespace h2j {

struct HuffPair { int a, b; };  // {value, prob} or {code, length}

// tools/microbench/k3_phases.cu: cycle stamps of one table's phases (thread 0 of the CTA that builds table 2 of frame 0)
#ifdef H2J_K3_CLOCKS
__device__ long long g_k3_clocks[16];
#define K3_STAMP(i) do { if (gt == 0 && S->stamp_on) g_k3_clocks[i] = clock64(); } while (0)
#else
#define K3_STAMP(i) do { } while (0)
#endif

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
                // The two scans below visit the elements in the reference's order and stop where it stops; only the LOADS
                // differ: four keys are requested at once (the walk is a chain of shared-memory round trips otherwise --
                // this thread is the frame's critical path).  Reading past `right` / below `left` is harmless: the keys
                // are only looked at under the reference's own bounds test, and the addresses stay inside the array
                // (start <= left, right <= end - 2, elements up to p[num + 2] exist).
                const int pivot = end[-1].b;  // end[-1] is not touched inside the partition loop
                while (left <= right) {
                    for (;;) {
                        const int k0 = left[0].b, k1 = left[1].b, k2 = left[2].b, k3 = left[3].b;
                        if (!(left <= right && k0 < pivot)) break;
                        left++;
                        if (!(left <= right && k1 < pivot)) break;
                        left++;
                        if (!(left <= right && k2 < pivot)) break;
                        left++;
                        if (!(left <= right && k3 < pivot)) break;
                        left++;
                    }
                    for (;;) {
                        HuffPair *r1 = right - 1 < start ? start : right - 1, *r2 = right - 2 < start ? start : right - 2,
                                 *r3 = right - 3 < start ? start : right - 3;
                        const int k0 = right[0].b, k1 = r1->b, k2 = r2->b, k3 = r3->b;
                        if (!(left <= right && k0 > pivot)) break;
                        right--;
                        if (!(left <= right && k1 > pivot)) break;
                        right--;
                        if (!(left <= right && k2 > pivot)) break;
                        right--;
                        if (!(left <= right && k3 > pivot)) break;
                        right--;
                    }
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

constexpr int kPmMaxItems = 516;  // a level holds at most 257 + 258/2 items; rounded up
struct HuffScratch {
    HuffPair sorted[260];                        // {value, count}, then sorted by count
    HuffPair distinct[260];                      // {value, code length}, then sorted by length
    int prob[2][kPmMaxItems];
    unsigned short leaves[17][kPmMaxItems + 2];  // leaves[t][p] = number of symbols among the first p items of level t
    int nl[17];                                  // symbols counted at each level by the back-trace
    unsigned char nbits_by_value[260];
    int warp_tot[4];
    int first_code[18], first_index[18];
    unsigned int bits_cnt[17];
#ifdef H2J_K3_CLOCKS
    int stamp_on;
#endif
};

__device__ __forceinline__ void group_sync(int group) { asm volatile("bar.sync %0, %1;" ::"r"(group + 1), "n"(kHuffGroup) : "memory"); }

// exclusive prefix sum of `v` over the 128 threads of a group; *total receives the group sum
__device__ __forceinline__ int group_excl_scan(int v, int gt, int group, int *warp_tot, int *total)
{
    const int lane = gt & 31, w = gt >> 5;
    int incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    group_sync(group);  // protects warp_tot against the previous use
    if (lane == 31) warp_tot[w] = incl;
    group_sync(group);
    int off = 0, tot = 0;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        if (i < w) off += warp_tot[i];
        tot += warp_tot[i];
    }
    *total = tot;
    return off + incl - v;
}

// One table, executed by the 128 threads of `group` (gt = thread index inside the group).
__device__ void build_one_table(const unsigned int *__restrict__ hist, HuffScratch *S, int gt, int group, uint8_t *g_bits, uint8_t *g_vals,
                                int *g_nvals, uint32_t *g_hcode)
{
    // ---- ff_mjpeg_encode_huffman_close: used symbols in increasing value order, plus the dummy (256, 0) ----
    K3_STAMP(0);
    const int h0 = (int)hist[2 * gt], h1 = (int)hist[2 * gt + 1];
    int nval;
    int pos = group_excl_scan((h0 != 0) + (h1 != 0), gt, group, S->warp_tot, &nval);
    if (h0) { S->sorted[pos].a = 2 * gt; S->sorted[pos].b = h0; pos++; }
    if (h1) { S->sorted[pos].a = 2 * gt + 1; S->sorted[pos].b = h1; }
    if (gt == 0) { S->sorted[nval].a = 256; S->sorted[nval].b = 0; }
    const int size = nval + 1;
    group_sync(group);
    K3_STAMP(1);
    if (gt == 0) av_qsort_pairs(S->sorted, size);
    group_sync(group);
    K3_STAMP(2);

    // ---- ff_mjpegenc_huffman_compute_bits, max_length 16: levels 0..15 take symbols, level 16 only packages ----
    for (int k = gt; k <= size; k += kHuffGroup) {
        S->leaves[0][k] = (unsigned short)k;
        if (k < size) S->prob[0][k] = S->sorted[k].b;
    }
    int n_prev = size, cur = 0;
    for (int t = 1; t <= 16; t++) {
        group_sync(group);
        const int *pp = S->prob[cur];
        int *np = S->prob[cur ^ 1];
        const int L = t < 16 ? size : 0, npk = n_prev >> 1;
        for (int idx = gt; idx < L + npk; idx += kHuffGroup) {
            if (idx < L) {
                // symbol idx: preceded by its idx predecessors and by every package whose sum is <= its count
                const int key = S->sorted[idx].b;
                int lo = 0, hi = npk;
                while (lo < hi) {
                    const int mid = (lo + hi) >> 1;
                    if (pp[2 * mid] + pp[2 * mid + 1] <= key) lo = mid + 1;
                    else hi = mid;
                }
                const int p = idx + lo;
                np[p] = key;
                S->leaves[t][p + 1] = (unsigned short)(idx + 1);
            } else {
                // package m: preceded by its m predecessors and by every symbol whose count is < its sum
                const int m = idx - L;
                const int key = pp[2 * m] + pp[2 * m + 1];
                int lo = 0, hi = L;
                while (lo < hi) {
                    const int mid = (lo + hi) >> 1;
                    if (S->sorted[mid].b < key) lo = mid + 1;
                    else hi = mid;
                }
                const int p = m + lo;
                np[p] = key;
                S->leaves[t][p + 1] = (unsigned short)lo;
            }
        }
        if (gt == 0) S->leaves[t][0] = 0;
        n_prev = L + npk;
        cur ^= 1;
    }
    group_sync(group);
    K3_STAMP(3);
    // ---- back-trace: how many symbols each level contributes to the first min(size-1, nitems) items ----
    if (gt == 0) {
        int p = size - 1 < n_prev ? size - 1 : n_prev;
        for (int t = 16; t >= 0; t--) {
            const int nl = p > 0 ? (int)S->leaves[t][p] : 0;
            S->nl[t] = nl;
            p = p > 0 ? 2 * (p - nl) : 0;
        }
    }
    for (int i = gt; i < 260; i += kHuffGroup) S->nbits_by_value[i] = 0;
    if (gt < 17) S->bits_cnt[gt] = 0;
    group_sync(group);
    for (int r = gt; r < size; r += kHuffGroup) {
        int nb = 0;
#pragma unroll
        for (int t = 0; t <= 16; t++) nb += S->nl[t] > r ? 1 : 0;
        S->nbits_by_value[S->sorted[r].a] = (unsigned char)nb;
    }
    group_sync(group);
    // ---- symbols with a code, in increasing value order (the dummy 256 is left out), sorted by length ----
    {
        const int b0 = S->nbits_by_value[2 * gt], b1 = S->nbits_by_value[2 * gt + 1];
        int tot;
        int p = group_excl_scan((b0 != 0) + (b1 != 0), gt, group, S->warp_tot, &tot);
        if (b0) { S->distinct[p].a = 2 * gt; S->distinct[p].b = b0; p++; }
        if (b1) { S->distinct[p].a = 2 * gt + 1; S->distinct[p].b = b1; }
        // tot == nval: every used symbol receives a code
    }
    group_sync(group);
    K3_STAMP(4);
    if (gt == 0) av_qsort_pairs(S->distinct, nval);
    group_sync(group);
    K3_STAMP(5);
    // ---- BITS / HUFFVAL, then ff_mjpeg_build_huffman_codes ----
    for (int i = gt; i < 256; i += kHuffGroup) {
        g_vals[i] = i < nval ? (uint8_t)S->distinct[i].a : 0;
        g_hcode[i] = 0;
        if (i < nval) atomicAdd(&S->bits_cnt[S->distinct[i].b], 1u);
    }
    group_sync(group);
    if (gt == 0) {
        int code = 0, k = 0;
        for (int i = 1; i <= 16; i++) {
            S->first_code[i] = code;
            S->first_index[i] = k;
            code = (code + (int)S->bits_cnt[i]) << 1;
            k += (int)S->bits_cnt[i];
        }
        *g_nvals = nval;
    }
    if (gt < 17) g_bits[gt] = gt ? (uint8_t)S->bits_cnt[gt] : 0;
    group_sync(group);
    for (int i = gt; i < nval; i += kHuffGroup) {
        const int len = S->distinct[i].b;
        const int code = S->first_code[len] + (i - S->first_index[len]);
        // K4a appends nb mantissa bits behind the code of a symbol whose low nibble is nb (DC: the symbol itself), so the
        // table holds the code already shifted into place and the total length: (code << nb) << 5 | (len + nb)
        const int sym = S->distinct[i].a, nb = sym & 15;
        g_hcode[sym] = (((uint32_t)code << nb) << 5) | (uint32_t)(len + nb);
    }
    K3_STAMP(6);
}

// Byte i of the picture header: SOI, COM, DQT, DHT (4 tables), SOF0, SOS (ff_mjpeg_encode_picture_header).
struct HeaderPlan {
    int off_dqt, off_dht, off_tab[4], off_sof, total;
};
__device__ __forceinline__ HeaderPlan header_plan(const int nvals[4], int comment_len)
{
    HeaderPlan h;
    h.off_dqt = 2 + 4 + comment_len + 1;
    h.off_dht = h.off_dqt + 4 + 1 + 64;
    int p = h.off_dht + 4;
    for (int t = 0; t < 4; t++) {
        h.off_tab[t] = p;
        p += 17 + nvals[t];
    }
    h.off_sof = p;
    h.total = p + 19 + 14;
    return h;
}
__device__ uint8_t header_byte(int i, const HeaderPlan &h, const FrameTab *T, const FrameLayout &L, const char *comment, int comment_len)
{
    if (i < h.off_dqt) {
        if (i < 6) {
            const int len = comment_len + 3;
            const uint8_t fixed[6] = {0xff, 0xd8, 0xff, 0xfe, (uint8_t)(len >> 8), (uint8_t)len};
            return fixed[i];
        }
        return i - 6 < comment_len ? (uint8_t)comment[i - 6] : 0;
    }
    if (i < h.off_dht) {
        const int k = i - h.off_dqt;
        if (k < 5) {
            const uint8_t fixed[5] = {0xff, 0xdb, 0, 67, 0};
            return fixed[k];
        }
        return T->dqt_zz[k - 5];
    }
    if (i < h.off_sof) {
        const int k = i - h.off_dht;
        if (k < 4) {
            const int size = h.off_sof - h.off_dht - 2;
            const uint8_t fixed[4] = {0xff, 0xc4, (uint8_t)(size >> 8), (uint8_t)size};
            return fixed[k];
        }
        int t = 3;
        while (i < h.off_tab[t]) t--;
        const int j = i - h.off_tab[t];
        if (j == 0) return (uint8_t)(((t >> 1) << 4) | (t & 1));  // class << 4 | id: DC0, DC1, AC0, AC1
        if (j <= 16) return T->bits[t][j];
        return T->vals[t][j - 17];
    }
    const int k = i - h.off_sof;
    const uint8_t tail[33] = {0xff, 0xc0, 0, 17, 8, (uint8_t)(L.h >> 8), (uint8_t)L.h, (uint8_t)(L.w >> 8), (uint8_t)L.w, 3,
                              1, 0x22, 0, 2, 0x11, 0, 3, 0x11, 0,
                              0xff, 0xda, 0, 12, 3, 1, 0x00, 2, 0x11, 3, 0x11, 0, 63, 0};
    return tail[k];
}

// grid 4 * frames CTAs of one group each: one table per CTA.  blockIdx.x / frames picks the table, AC tables first (they
// take ~10x longer than the DC ones and decide the kernel's duration: with one CTA per FRAME the four scratch areas
// allowed two CTAs per SM, and 512 frames needed two waves of them).  The frame's CTA that finishes last writes the header.
__global__ void __launch_bounds__(kHuffGroup) huffman_kernel(FrameLayout L, FrameTab *__restrict__ tabs, FrameState *__restrict__ state,
                                                             int n_frames, uint8_t *__restrict__ out, long long out_cap,
                                                             const char *__restrict__ comment, int comment_len)
{
    __shared__ HuffScratch scratch;
    __shared__ int s_last;
    const int tid = threadIdx.x;
    const int slot = blockIdx.x / n_frames, f = blockIdx.x - slot * n_frames;
    const int table = slot ^ 2;  // 2, 3 (AC luma, AC chroma), then 0, 1 (DC)
    FrameTab *T = tabs + f;
#ifdef H2J_K3_CLOCKS
    if (tid == 0) scratch.stamp_on = (table == 2 && f == 0);
    __syncthreads();
#endif
    build_one_table(state[f].hist[table], &scratch, tid, 0, T->bits[table], T->vals[table], &T->nvals[table], T->hcode[table]);
    __threadfence();
    __syncthreads();
    if (tid == 0) s_last = atomicAdd(&state[f].k3_done, 1u) == 3u;
    __syncthreads();
    if (!s_last) return;
    __threadfence();  // the other three tables were written by other CTAs
    const volatile int *nv = T->nvals;
    const int nvals[4] = {nv[0], nv[1], nv[2], nv[3]};
    const HeaderPlan h = header_plan(nvals, comment_len);
    uint8_t *o = out + (long long)f * out_cap;
    for (int i = tid; i < h.total; i += kHuffGroup)
        if (i < out_cap) o[i] = header_byte(i, h, T, L, comment, comment_len);
    if (tid == 0) {
        T->header_bytes = h.total;
        if (h.total > out_cap) T->status = -4;
    }
}

}  // namespace h2j
