// Issue cost of IMAD.HI against IMAD + SHF on sm_100a: is (a * (c << 16)) >> 32 in ONE instruction cheaper than
// (a * c) >> 16 in two?   nvcc -arch=sm_100a -o imad_hi imad_hi.cu && ./imad_hi
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE> __global__ void k(int *out, int a0, int c, long long *cycles)
{
    int v[16];
#pragma unroll
    for (int i = 0; i < 16; i++) v[i] = a0 + threadIdx.x + i;
    const long long t0 = clock64();
    for (int it = 0; it < 1024; it++) {
#pragma unroll
        for (int i = 0; i < 16; i++) {
            if (MODE == 0) v[i] = (v[i] * c) >> 16;                 // IMAD + SHF
            else if (MODE == 1) v[i] = __mulhi(v[i], c << 16);      // IMAD.HI
            else v[i] = v[i] * c + i;                                // IMAD alone
        }
    }
    const long long t1 = clock64();
    int s = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) s += v[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}
int main()
{
    int *out; long long *cyc, h;
    cudaMalloc(&out, 148 * 4 * 512 * 4); cudaMalloc(&cyc, 8);
    for (int warps = 4; warps <= 16; warps *= 2) {
        for (int mode = 0; mode < 3; mode++) {
            for (int rep = 0; rep < 2; rep++) {
                if (mode == 0) k<0><<<148, warps * 32>>>(out, 3, 23170, cyc);
                else if (mode == 1) k<1><<<148, warps * 32>>>(out, 3, 23170, cyc);
                else k<2><<<148, warps * 32>>>(out, 3, 23170, cyc);
                cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
            }
            const char *names[3] = {"IMAD+SHF", "IMAD.HI ", "IMAD    "};
            printf("%2d warps/SM  %s  %.2f cycles per op per warp-scheduler slot (%lld cycles / %d ops)\n", warps, names[mode], (double)h / (1024.0 * 16 * warps / 4), h, 1024 * 16);
        }
    }
    return 0;
}
