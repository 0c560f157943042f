// K6: packing of a batch's JPEGs for one D2H copy.
#pragma once
#include "h2j_common.cuh"

namespace h2j {

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

}  // namespace h2j
