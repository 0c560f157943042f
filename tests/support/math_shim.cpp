// Host-side build of csrc/h2j_math.cuh so the exact integer code the kernels run can be compared with
// the oracle without a GPU (tests/test_math_host.py).
#include "../../h264-h265-to-jpeg_b200/csrc/h2j_math.cuh"
extern "C" {
void shim_fdct(const int16_t *in, int16_t *out, int n)
{
    for (int b = 0; b < n; b++) {
        int v[64];
        for (int i = 0; i < 64; i++) v[i] = in[64 * b + i];
        h2j::fdct_8x8(v);
        for (int i = 0; i < 64; i++) out[64 * b + i] = (int16_t)v[i];
    }
}
// raster-order quantisation of fdct output with the matrices for `qscale`
void shim_quant(const int16_t *in, int16_t *out, int n, int qscale, const uint16_t *mpeg1_intra)
{
    uint32_t pk[64]; uint8_t dqt[64];
    for (int i = 0; i < 64; i++) h2j::quant_entry(qscale, mpeg1_intra[i], i, &dqt[i], &pk[i]);
    for (int b = 0; b < n; b++) {
        out[64 * b] = (int16_t)h2j::quant_dc(in[64 * b]);
        for (int i = 1; i < 64; i++) out[64 * b + i] = (int16_t)h2j::quant_ac(in[64 * b + i], pk[i]);
    }
}
void shim_matrix(int qscale, const uint16_t *mpeg1_intra, uint8_t *dqt, uint32_t *pk)
{
    for (int i = 0; i < 64; i++) h2j::quant_entry(qscale, mpeg1_intra[i], i, &dqt[i], &pk[i]);
}
int shim_lambda_to_qscale(int l) { return h2j::lambda_to_qscale(l); }
}
