// Integer arithmetic of the hot path, written once and usable from device code and (for the CPU-side
// unit test tests/test_math_host.py builds) from host code.  Every function here is bit-exact with the
// x86 code the reference ends up running inside libavcodec 58.117.101:
//   ff_fdct_sse2            libavcodec/x86/fdct.c              (dct_algo = FF_DCT_AUTO on x86)
//   dct_quantize_ssse3      libavcodec/x86/mpegvideoenc_template.c
// for 8-bit input samples.  The SSE2 code uses saturating 16-bit adds (paddsw/psubsw/packssdw); for
// samples in [0,255] no intermediate can saturate (bounds in DESIGN.md §4.2, checked exhaustively on
// extreme patterns by tests/test_fdct_bounds.py), so plain 32-bit adds are used here.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define H2J_HD __host__ __device__ __forceinline__
#else
#define H2J_HD static inline
#endif

namespace h2j {

// pmulhw: signed 16x16 -> high 16 bits.
H2J_HD int mulh16(int a, int c) { return (a * c) >> 16; }

// Column pass of ff_fdct_sse2 on one column (x[0..7] = rows), in place.
// Constants: tg_1_16 = 13036, tg_2_16 = 27146, tg_3_16 = -21746 (tg3 - 1), ocos_4_16 = 23170.
// The SSE2 code shifts the eight butterfly outputs left (psllw 3, or 4 for t5/t6) before anything else.  Here the
// shifts are folded into the constants of the multiplies that follow -- ((a << s) * c) >> 16 == (a * (c << s)) >> 16
// exactly, nothing overflows 32 bits for 8-bit samples -- and into the adds (a * 8 + b is one IMAD/LEA).
H2J_HD void fdct_col(int &x0, int &x1, int &x2, int &x3, int &x4, int &x5, int &x6, int &x7)
{
    const int s0 = x0 + x7, s1 = x1 + x6, s2 = x2 + x5, s3 = x3 + x4;
    const int d0 = x0 - x7, d1 = x1 - x6, d2 = x2 - x5, d3 = x3 - x4;
    const int um12 = s1 - s2, up12 = s1 + s2, um03 = s0 - s3, up03 = s0 + s3;   // tm12 = um12 << 3, ...
    // rows 0 and 4 leave WITHOUT their << 3: fdct_row<0>, the only consumer of these two rows, carries the factor 8 in its
    // constants instead (exact: (a * 8) * c == a * (8 * c) modulo 2^32, and nothing here exceeds 31 bits anyway)
    const int y0 = up03 + up12;
    const int y4 = up03 - up12;
    const int y2 = (((um12 * (27146 * 8)) >> 16) + um03 * 8) | 1;
    const int y6 = (((um03 * (27146 * 8)) >> 16) - um12 * 8) | 1;
    const int tp65 = (((d1 + d2) * (23170 * 16)) >> 16) | 1;                     // t5, t6 carry << 4
    const int tm65 = ((d1 - d2) * (23170 * 16)) >> 16;
    const int tp465 = d3 * 8 + tm65, tm465 = d3 * 8 - tm65;
    const int tm765 = d0 * 8 - tp65, tp765 = d0 * 8 + tp65;
    const int y1 = (mulh16(tp465, 13036) + tp765) | 1;
    // pmulhw(a, tg3 - 1) + a == (a * (tg3 - 1 + 65536)) >> 16 exactly (a * 65536 has no low bits): one add less each
    const int y3 = tm765 - mulh16(tm465, 65536 - 21746);
    const int y5 = mulh16(tm765, 65536 - 21746) + tm465;
    const int y7 = mulh16(tp765, 13036) - tp465;
    x0 = y0; x1 = y1; x2 = y2; x3 = y3; x4 = y4; x5 = y5; x6 = y6; x7 = y7;
}

// Row pass of ff_fdct_sse2 on one row, in place.  TAB selects the coefficient set
// (rows 0/4 -> 0, 1/7 -> 1, 2/6 -> 2, 3/5 -> 3); the four pmaddwd dot products per output are regrouped
// algebraically (exact: all sums are taken modulo 2^32 before the arithmetic shift).
template <int TAB>
H2J_HD void fdct_row(int &x0, int &x1, int &x2, int &x3, int &x4, int &x5, int &x6, int &x7)
{
    // TAB 0 (rows 0 and 4) is scaled by 8: fdct_col hands those two rows over without their << 3
    constexpr int C1 = TAB == 0 ? 22725 * 8 : TAB == 1 ? 31521 : TAB == 2 ? 29692 : 26722;
    constexpr int C2 = TAB == 0 ? 21407 * 8 : TAB == 1 ? 29692 : TAB == 2 ? 27969 : 25172;
    constexpr int C3 = TAB == 0 ? 19266 * 8 : TAB == 1 ? 26722 : TAB == 2 ? 25172 : 22654;
    constexpr int C4 = TAB == 0 ? 16384 * 8 : TAB == 1 ? 22725 : TAB == 2 ? 21407 : 19266;
    constexpr int C5 = TAB == 0 ? 12873 * 8 : TAB == 1 ? 17855 : TAB == 2 ? 16819 : 15137;
    constexpr int C6 = TAB == 0 ? 8867 * 8 : TAB == 1 ? 12299 : TAB == 2 ? 11585 : 10426;
    constexpr int C7 = TAB == 0 ? 4520 * 8 : TAB == 1 ? 6270 : TAB == 2 ? 5906 : 5315;
    constexpr int RND = 1 << 16;
    const int a0 = x0 + x7, a1 = x1 + x6, a2 = x2 + x5, a3 = x3 + x4;
    const int b0 = x0 - x7, b1 = x1 - x6, b2 = x2 - x5, b3 = x3 - x4;
    const int s03 = a0 + a3, s12 = a1 + a2, d03 = a0 - a3, d12 = a1 - a2;
    if (TAB == 0) {
        // C4 == 2^17 here, and |s03 +- s12| < 2^14 for 8-bit samples: (k * 2^17 + 2^16) >> 17 == k
        x0 = s03 + s12;
        x4 = s03 - s12;
    } else {
        x0 = ((s03 + s12) * C4 + RND) >> 17;
        x4 = ((s03 - s12) * C4 + RND) >> 17;
    }
    x2 = (d03 * C2 + d12 * C6 + RND) >> 17;
    x6 = (d03 * C6 - d12 * C2 + RND) >> 17;
    x1 = (b0 * C1 + b1 * C3 + b2 * C5 + b3 * C7 + RND) >> 17;
    x3 = (b0 * C3 - b1 * C7 - b2 * C1 - b3 * C5 + RND) >> 17;
    x5 = (b0 * C5 - b1 * C1 + b2 * C7 + b3 * C3 + RND) >> 17;
    x7 = (b0 * C7 - b1 * C5 + b2 * C3 - b3 * C1 + RND) >> 17;
}

// Whole 8x8 block held in 64 scalars, raster order, in place.
H2J_HD void fdct_8x8(int (&v)[64])
{
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int c = 0; c < 8; c++)
        fdct_col(v[c], v[8 + c], v[16 + c], v[24 + c], v[32 + c], v[40 + c], v[48 + c], v[56 + c]);
    fdct_row<0>(v[0], v[1], v[2], v[3], v[4], v[5], v[6], v[7]);
    fdct_row<1>(v[8], v[9], v[10], v[11], v[12], v[13], v[14], v[15]);
    fdct_row<2>(v[16], v[17], v[18], v[19], v[20], v[21], v[22], v[23]);
    fdct_row<3>(v[24], v[25], v[26], v[27], v[28], v[29], v[30], v[31]);
    fdct_row<0>(v[32], v[33], v[34], v[35], v[36], v[37], v[38], v[39]);
    fdct_row<3>(v[40], v[41], v[42], v[43], v[44], v[45], v[46], v[47]);
    fdct_row<2>(v[48], v[49], v[50], v[51], v[52], v[53], v[54], v[55]);
    fdct_row<1>(v[56], v[57], v[58], v[59], v[60], v[61], v[62], v[63]);
}

// ---- quantiser --------------------------------------------------------------------------------
// dct_quantize_ssse3 on an intra MJPEG block:
//   AC: level = sign(x) * (((|x| + bias16) * qmat16) >> 16)
// rewritten without the abs/sign round trip:  x >= 0: (x*q + bq) >> 16,  x < 0: (x*q + (65535 - bq)) >> 16
// with bq = bias16*qmat16 (< 65536).  One packed constant per coefficient: q in the low 16 bits,
// bq in the high 16 bits.
H2J_HD uint32_t quant_pack(uint32_t qmat16, uint32_t bias16) { return qmat16 | ((bias16 * qmat16) << 16); }

H2J_HD int quant_ac2(int x, int q, int bq)
{
    const int s = x >> 31;                 // 0 or -1
    const int c = bq ^ (s & 0xffff);       // bq or 65535 - bq
    return (x * q + c) >> 16;
}
// the same product before the final shift: the level is its upper 16 bits (|level| < 2^15 for FDCT outputs)
H2J_HD int quant_ac2_hi(int x, int q, int bq)
{
    const int s = x >> 31;
    return x * q + (bq ^ (s & 0xffff));
}
H2J_HD int quant_ac(int x, uint32_t packed) { return quant_ac2(x, (int)(packed & 0xffffu), (int)(packed >> 16)); }
// DC: ((block[0] >> 2) + q) * ff_inverse[2q] >> 32 with q = 8  ==  ((x >> 2) + 8) >> 4  (x >= 0)
H2J_HD int quant_dc(int x) { return ((x >> 2) + 8) >> 4; }

// MJPEG matrix set-up (mpegvideo_enc.c encode_picture + ff_convert_matrix, SIMD-fdct branch) for
// raster index i: returns the DQT byte and the packed quantiser constant.
H2J_HD void quant_entry(int qscale, int mpeg1_intra_i, int i, uint8_t *dqt_byte, uint32_t *packed)
{
    int m = (mpeg1_intra_i * qscale) >> 3;
    if (m > 255) m = 255;
    if (i == 0) m = 8;                      // ff_mpeg2_dc_scale_table[0][8]
    int q16 = (2 << 16) / (16 * m);         // qscale2 = 8 << 1
    if (q16 == 0 || q16 == 128 * 256) q16 = 128 * 256 - 1;
    const int b16 = (96 * 256 + (q16 >> 1)) / q16; // ROUNDED_DIV(intra_quant_bias << 8, q16), bias = 3<<5
    *dqt_byte = (uint8_t)m;
    *packed = quant_pack((uint32_t)q16, (uint32_t)b16);
}

// update_qscale(): lambda -> qscale, clipped to the codec context's qmin/qmax (2, 31)
H2J_HD int lambda_to_qscale(int lambda)
{
    int q = (lambda * 139 + 128 * 64) >> 14;
    return q < 2 ? 2 : (q > 31 ? 31 : q);
}

}  // namespace h2j
