/*
 * TEST INFRASTRUCTURE — not product code.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load this.
 *
 * CPU restatement (plain C, scalar) of the YUV -> JPEG stage the reference runs:
 *   reference src/Encoder.cpp:104-308  Encoder::yuv2Jpeg() opens libavcodec's "mjpeg" encoder with
 *   pix_fmt = AV_PIX_FMT_YUVJ420P (Encoder.cpp:162), time_base 1/25 (Encoder.cpp:201), every other
 *   option at its default, sends ONE frame (Encoder.cpp:250) and writes the packet through the raw
 *   "mjpeg" muxer (Encoder.cpp:278).  No sws_scale call exists in the reference: the decoded planes are
 *   handed to the encoder as they are.
 *
 * The algorithm itself lives in a third-party dependency that is NOT in /root/reference as source:
 *   FFmpeg git-2021-01-28-6fd0116 (libavcodec 58.117.101, libavutil 56.63.101, libswscale 5.8.100),
 *   vendored by the reference only as binaries under lib/ffmpeg/x86_64_shared.  What is restated here
 *   is that version's published algorithm, as selected at run time on x86-64 with default options:
 *     libavcodec/mpegvideo_enc.c   load_input_picture (edge replication), mb_var_thread, encode_picture
 *                                  (MJPEG matrix set-up), ff_convert_matrix, update_qscale
 *     libavcodec/ratecontrol.c     ff_rate_estimate_qscale / get_qscale / modify_qscale (first frame)
 *     libavcodec/x86/fdct.c        ff_fdct_sse2   (dct_algo=FF_DCT_AUTO on x86 -> NOT jpeg_fdct_islow)
 *     libavcodec/x86/mpegvideoenc_template.c  dct_quantize_{sse2,ssse3}
 *     libavcodec/mjpegenc.c        record_block / ff_mjpeg_encode_coef / ff_mjpeg_encode_picture_frame
 *     libavcodec/mjpegenc_huffman.c ff_mjpegenc_huffman_compute_bits (package-merge) / ..._close
 *     libavutil/qsort.h            AV_QSORT (unstable; tie order matters for the DHT bytes)
 *     libavcodec/mjpegenc_common.c ff_mjpeg_encode_picture_header, ff_mjpeg_escape_FF, trailer
 *     libswscale/swscale.c         lumRangeToJpeg_c / chrRangeToJpeg_c + output dither (range_mode 1)
 *   The FDCT constants and instruction order were read back from the vendored binary
 *   (objdump of libavcodec.so.58 at the address AVDCT.fdct resolves to) rather than from memory.
 *
 * PARITY IS PINNED: tests/test_oracle_vs_reference.py runs this file against the real reference
 * (oracle/_ref/libh2j_ref.so = the reference's own Encoder.cpp + its vendored libavcodec) on the
 * reference's test/img fixtures and on seeded random / extreme frames, and tests/golden/ holds vectors
 * generated from the reference by tests/golden/make_golden.py.
 */
#define _GNU_SOURCE
#include "mjpeg_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------------ */
/* tables                                                                                           */
/* ------------------------------------------------------------------------------------------------ */
static const uint8_t zigzag_direct[64] = {
    0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48,
    41, 34, 27, 20, 13, 6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,
    30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};

/* ff_mpeg1_default_intra_matrix (mpeg12data.c), raster order */
static const uint16_t mpeg1_default_intra_matrix[64] = {
    8,  16, 19, 22, 26, 27, 29, 34, 16, 16, 22, 24, 27, 29, 34, 37, 19, 22, 26, 27, 29, 34,
    34, 38, 22, 22, 26, 27, 29, 34, 37, 40, 22, 26, 27, 29, 32, 35, 40, 48, 26, 27, 29, 32,
    35, 40, 48, 58, 26, 27, 29, 34, 38, 46, 56, 69, 27, 29, 35, 38, 46, 56, 69, 83};

/* ------------------------------------------------------------------------------------------------ */
/* ff_fdct_sse2 (libavcodec/x86/fdct.c), lane semantics of the SSE2 instructions                   */
/* ------------------------------------------------------------------------------------------------ */
static inline int16_t sat16(int v) { return (int16_t)(v > 32767 ? 32767 : (v < -32768 ? -32768 : v)); }
static inline int16_t adds(int16_t a, int16_t b) { return sat16((int)a + b); }            /* paddsw */
static inline int16_t subs(int16_t a, int16_t b) { return sat16((int)a - b); }            /* psubsw */
static inline int16_t mulh(int16_t a, int16_t b) { return (int16_t)(((int)a * b) >> 16); } /* pmulhw */
static inline int16_t shl(int16_t a, int n) { return (int16_t)((uint16_t)a << n); }        /* psllw  */

#define TG_1_16 13036
#define TG_2_16 27146
#define TG_3_16 (-21746)
#define OCOS_4_16 23170
#define SHIFT_FRW_COL 3
#define SHIFT_FRW_ROW 17
#define RND_FRW_ROW (1 << (SHIFT_FRW_ROW - 1))

/* tab_frw_01234567_sse2: rows 0/4, 1/7, 2/6, 3/5 (dumped from the vendored libavcodec.so.58 .rodata) */
static const int16_t tab_frw_sse2[4][32] = {
    {16384, 16384, 22725, 19266, -8867, -21407, -22725, -12873, 16384, 16384, 12873, 4520, 21407, 8867, 19266, -4520,
     -16384, 16384, 4520, 19266, 8867, -21407, 4520, -12873, 16384, -16384, 12873, -22725, 21407, -8867, 19266, -22725},
    {22725, 22725, 31521, 26722, -12299, -29692, -31521, -17855, 22725, 22725, 17855, 6270, 29692, 12299, 26722, -6270,
     -22725, 22725, 6270, 26722, 12299, -29692, 6270, -17855, 22725, -22725, 17855, -31521, 29692, -12299, 26722, -31521},
    {21407, 21407, 29692, 25172, -11585, -27969, -29692, -16819, 21407, 21407, 16819, 5906, 27969, 11585, 25172, -5906,
     -21407, 21407, 5906, 25172, 11585, -27969, 5906, -16819, 21407, -21407, 16819, -29692, 27969, -11585, 25172, -29692},
    {19266, 19266, 26722, 22654, -10426, -25172, -26722, -15137, 19266, 19266, 15137, 5315, 25172, 10426, 22654, -5315,
     -19266, 19266, 5315, 22654, 10426, -25172, 5315, -15137, 19266, -19266, 15137, -26722, 25172, -10426, 22654, -26722}};
static const int row_table_of[8] = {0, 1, 2, 3, 0, 3, 2, 1};

static void fdct_col_sse2(const int16_t *in, int16_t *out) /* one column, stride 8 */
{
    const int16_t x0 = in[0], x1 = in[8], x2 = in[16], x3 = in[24], x4 = in[32], x5 = in[40], x6 = in[48], x7 = in[56];
    int16_t t0 = shl(adds(x0, x7), SHIFT_FRW_COL);
    int16_t t1 = shl(adds(x1, x6), SHIFT_FRW_COL);
    int16_t t2 = shl(adds(x5, x2), SHIFT_FRW_COL);
    int16_t t3 = shl(adds(x3, x4), SHIFT_FRW_COL);
    int16_t tm12 = subs(t1, t2), tp12 = adds(t1, t2);
    int16_t tm03 = subs(t0, t3), tp03 = adds(t0, t3);
    int16_t y2 = adds(mulh(TG_2_16, tm12), tm03) | 1;
    int16_t y4 = subs(tp03, tp12);
    int16_t y0 = adds(tp03, tp12);
    int16_t y6 = subs(mulh(tm03, TG_2_16), tm12) | 1;
    int16_t t6 = shl(subs(x1, x6), SHIFT_FRW_COL + 1);
    int16_t t5 = shl(subs(x2, x5), SHIFT_FRW_COL + 1);
    int16_t t4 = shl(subs(x3, x4), SHIFT_FRW_COL);
    int16_t t7 = shl(subs(x0, x7), SHIFT_FRW_COL);
    int16_t tp65 = mulh(adds(t6, t5), OCOS_4_16) | 1;
    int16_t tm65 = mulh(subs(t6, t5), OCOS_4_16);
    int16_t tp465 = adds(t4, tm65), tm465 = subs(t4, tm65);
    int16_t tm765 = subs(t7, tp65), tp765 = adds(t7, tp65);
    int16_t y1 = adds(mulh(TG_1_16, tp465), tp765) | 1;
    int16_t a3 = adds(mulh(TG_3_16, tm465), tm465); /* tm465 * tg3 (tg3 is stored minus one) */
    int16_t b3 = adds(mulh(TG_3_16, tm765), tm765);
    int16_t y3 = subs(tm765, a3);
    int16_t y5 = adds(b3, tm465);
    int16_t y7 = subs(mulh(tp765, TG_1_16), tp465);
    out[0] = y0; out[8] = y1; out[16] = y2; out[24] = y3; out[32] = y4; out[40] = y5; out[48] = y6; out[56] = y7;
}

static void fdct_row_sse2(const int16_t *in, int16_t *out, const int16_t *T)
{
    int16_t w1[8], w2[8];
    /* a_k = x_k + x_{7-k}, b_k = x_k - x_{7-k}; punpckldq / pshufd 0x4e word layout */
    int16_t a[4], b[4];
    for (int k = 0; k < 4; k++) { a[k] = adds(in[k], in[7 - k]); b[k] = subs(in[k], in[7 - k]); }
    w1[0] = a[0]; w1[1] = a[1]; w1[2] = b[0]; w1[3] = b[1]; w1[4] = a[2]; w1[5] = a[3]; w1[6] = b[2]; w1[7] = b[3];
    for (int k = 0; k < 8; k++) w2[k] = w1[(k + 4) & 7];
    for (int d = 0; d < 4; d++) {
        /* pmaddwd + paddd wrap modulo 2^32 */
        uint32_t lo = (uint32_t)((int)w1[2 * d] * T[2 * d]) + (uint32_t)((int)w1[2 * d + 1] * T[2 * d + 1]) +
                      (uint32_t)((int)w2[2 * d] * T[8 + 2 * d]) + (uint32_t)((int)w2[2 * d + 1] * T[8 + 2 * d + 1]) +
                      (uint32_t)RND_FRW_ROW;
        uint32_t hi = (uint32_t)((int)w2[2 * d] * T[16 + 2 * d]) + (uint32_t)((int)w2[2 * d + 1] * T[16 + 2 * d + 1]) +
                      (uint32_t)((int)w1[2 * d] * T[24 + 2 * d]) + (uint32_t)((int)w1[2 * d + 1] * T[24 + 2 * d + 1]) +
                      (uint32_t)RND_FRW_ROW;
        out[d] = sat16((int32_t)lo >> SHIFT_FRW_ROW);     /* psrad + packssdw */
        out[4 + d] = sat16((int32_t)hi >> SHIFT_FRW_ROW);
    }
}

void orc_fdct_sse2(int16_t blk[64])
{
    int16_t tmp[64];
    for (int c = 0; c < 8; c++) fdct_col_sse2(blk + c, tmp + c);
    for (int r = 0; r < 8; r++) fdct_row_sse2(tmp + 8 * r, blk + 8 * r, tab_frw_sse2[row_table_of[r]]);
}

/* ------------------------------------------------------------------------------------------------ */
/* jpeg_fdct_islow_8 (libavcodec/jfdctint_template.c) — dct_algo=FF_DCT_INT, kept for completeness  */
/* ------------------------------------------------------------------------------------------------ */
void orc_fdct_islow(int16_t data[64])
{
    enum { CONST_BITS = 13, PASS1_BITS = 4 };
#define FIX_0_298631336 2446
#define FIX_0_390180644 3196
#define FIX_0_541196100 4433
#define FIX_0_765366865 6270
#define FIX_0_899976223 7373
#define FIX_1_175875602 9633
#define FIX_1_501321110 12299
#define FIX_1_847759065 15137
#define FIX_1_961570560 16069
#define FIX_2_053119869 16819
#define FIX_2_562915447 20995
#define FIX_3_072711026 25172
#define DESCALE(x, n) (((x) + (1 << ((n)-1))) >> (n))
    int tmp0, tmp1, tmp2, tmp3, tmp4, tmp5, tmp6, tmp7, tmp10, tmp11, tmp12, tmp13, z1, z2, z3, z4, z5;
    int16_t *p = data;
    for (int ctr = 0; ctr < 8; ctr++, p += 8) {
        tmp0 = p[0] + p[7]; tmp7 = p[0] - p[7]; tmp1 = p[1] + p[6]; tmp6 = p[1] - p[6];
        tmp2 = p[2] + p[5]; tmp5 = p[2] - p[5]; tmp3 = p[3] + p[4]; tmp4 = p[3] - p[4];
        tmp10 = tmp0 + tmp3; tmp13 = tmp0 - tmp3; tmp11 = tmp1 + tmp2; tmp12 = tmp1 - tmp2;
        p[0] = (int16_t)((tmp10 + tmp11) * (1 << PASS1_BITS));
        p[4] = (int16_t)((tmp10 - tmp11) * (1 << PASS1_BITS));
        z1 = (tmp12 + tmp13) * FIX_0_541196100;
        p[2] = (int16_t)DESCALE(z1 + tmp13 * FIX_0_765366865, CONST_BITS - PASS1_BITS);
        p[6] = (int16_t)DESCALE(z1 + tmp12 * (-FIX_1_847759065), CONST_BITS - PASS1_BITS);
        z1 = tmp4 + tmp7; z2 = tmp5 + tmp6; z3 = tmp4 + tmp6; z4 = tmp5 + tmp7;
        z5 = (z3 + z4) * FIX_1_175875602;
        tmp4 *= FIX_0_298631336; tmp5 *= FIX_2_053119869; tmp6 *= FIX_3_072711026; tmp7 *= FIX_1_501321110;
        z1 *= -FIX_0_899976223; z2 *= -FIX_2_562915447; z3 *= -FIX_1_961570560; z4 *= -FIX_0_390180644;
        z3 += z5; z4 += z5;
        p[7] = (int16_t)DESCALE(tmp4 + z1 + z3, CONST_BITS - PASS1_BITS);
        p[5] = (int16_t)DESCALE(tmp5 + z2 + z4, CONST_BITS - PASS1_BITS);
        p[3] = (int16_t)DESCALE(tmp6 + z2 + z3, CONST_BITS - PASS1_BITS);
        p[1] = (int16_t)DESCALE(tmp7 + z1 + z4, CONST_BITS - PASS1_BITS);
    }
    p = data;
    for (int ctr = 0; ctr < 8; ctr++, p++) {
        tmp0 = p[0] + p[56]; tmp7 = p[0] - p[56]; tmp1 = p[8] + p[48]; tmp6 = p[8] - p[48];
        tmp2 = p[16] + p[40]; tmp5 = p[16] - p[40]; tmp3 = p[24] + p[32]; tmp4 = p[24] - p[32];
        tmp10 = tmp0 + tmp3; tmp13 = tmp0 - tmp3; tmp11 = tmp1 + tmp2; tmp12 = tmp1 - tmp2;
        p[0] = (int16_t)DESCALE(tmp10 + tmp11, PASS1_BITS);
        p[32] = (int16_t)DESCALE(tmp10 - tmp11, PASS1_BITS);
        z1 = (tmp12 + tmp13) * FIX_0_541196100;
        p[16] = (int16_t)DESCALE(z1 + tmp13 * FIX_0_765366865, CONST_BITS + PASS1_BITS);
        p[48] = (int16_t)DESCALE(z1 + tmp12 * (-FIX_1_847759065), CONST_BITS + PASS1_BITS);
        z1 = tmp4 + tmp7; z2 = tmp5 + tmp6; z3 = tmp4 + tmp6; z4 = tmp5 + tmp7;
        z5 = (z3 + z4) * FIX_1_175875602;
        tmp4 *= FIX_0_298631336; tmp5 *= FIX_2_053119869; tmp6 *= FIX_3_072711026; tmp7 *= FIX_1_501321110;
        z1 *= -FIX_0_899976223; z2 *= -FIX_2_562915447; z3 *= -FIX_1_961570560; z4 *= -FIX_0_390180644;
        z3 += z5; z4 += z5;
        p[56] = (int16_t)DESCALE(tmp4 + z1 + z3, CONST_BITS + PASS1_BITS);
        p[40] = (int16_t)DESCALE(tmp5 + z2 + z4, CONST_BITS + PASS1_BITS);
        p[24] = (int16_t)DESCALE(tmp6 + z2 + z3, CONST_BITS + PASS1_BITS);
        p[8] = (int16_t)DESCALE(tmp7 + z1 + z4, CONST_BITS + PASS1_BITS);
    }
#undef DESCALE
}

/* ------------------------------------------------------------------------------------------------ */
/* padded picture (mpegvideo_enc.c load_input_picture + draw_edges)                                 */
/* ------------------------------------------------------------------------------------------------ */
/* Pixel (x, y) of the encoder's internal picture: the copied w x h region, then the right edge
 * replicated from column w-1 and the bottom edge replicated from row h-1.  For planes 1/2 the encoder
 * copies w>>1 x h>>1 (floor) — the odd last chroma column/row of the decoder's frame is never read. */
static inline uint8_t pad_px(const uint8_t *p, int stride, int w, int h, int x, int y)
{
    if (x >= w) x = w - 1;
    if (y >= h) y = h - 1;
    return p[(long)y * stride + x];
}

/* mb_var_thread (mpegvideo_enc.c): sum over macroblocks of the luma variance term */
int64_t orc_mb_var_sum(const uint8_t *y, int ystride, int w, int h)
{
    const int mbw = (w + 15) >> 4, mbh = (h + 15) >> 4;
    int64_t total = 0;
    for (int my = 0; my < mbh; my++)
        for (int mx = 0; mx < mbw; mx++) {
            int sum = 0, norm = 0;
            for (int r = 0; r < 16; r++)
                for (int c = 0; c < 16; c++) {
                    int v = pad_px(y, ystride, w, h, mx * 16 + c, my * 16 + r);
                    sum += v;      /* pix_sum   */
                    norm += v * v; /* pix_norm1 */
                }
            int varc = (int)((unsigned)norm - (((unsigned)sum * (unsigned)sum) >> 8) + 500 + 128) >> 8;
            total += varc;
        }
    return total;
}

/* ------------------------------------------------------------------------------------------------ */
/* ratecontrol.c, first frame of a fresh context (the reference builds a new Encoder per image)      */
/* ------------------------------------------------------------------------------------------------ */
#define FF_QP2LAMBDA 118
#define FF_LAMBDA_SHIFT 7
#define FF_LAMBDA_SCALE (1 << FF_LAMBDA_SHIFT)
#define FF_LAMBDA_MAX (256 * 128 - 1)
static int clipi(int a, int lo, int hi) { return a < lo ? lo : (a > hi ? hi : a); }

int orc_rate_control_qscale(int64_t mb_var_sum, int64_t pts, int *lambda_out)
{
    /* AVCodecContext defaults (libavcodec/options_table.h) */
    const int64_t bit_rate = 200000;
    const int bit_rate_tolerance = 200000 * 20;
    const float qcompress = 0.5f, qblur = 0.5f;
    const float i_quant_factor = -0.8f, i_quant_offset = 0.0f;
    const int avctx_qmin = 2, avctx_qmax = 31;
    const int lmin = 2 * FF_QP2LAMBDA, lmax = 31 * FF_QP2LAMBDA; /* mpegvideo lmin/lmax option defaults */
    const double fps = 1.0 / (1.0 / 25.0) / 1.0;                  /* get_fps(): time_base {1,25} */

    /* The frame's pts does not take part: measured against the vendored libavcodec, the JPEG is
     * byte-identical for pts = AV_NOPTS_VALUE, 0, 1 .. 100000 and negative values (tests pin this), i.e.
     * wanted_bits is 0 for the first picture of a fresh context.  (Raw .h264/.h265 demuxing, the only
     * input the reference build supports, yields AV_NOPTS_VALUE anyway.) */
    (void)pts;
    pts = 0;

    /* get_qminmax() for an I picture */
    int qmin = (int)(lmin * fabs((double)i_quant_factor) + i_quant_offset + 0.5);
    int qmax = (int)(lmax * fabs((double)i_quant_factor) + i_quant_offset + 0.5);
    qmin = clipi(qmin, 1, FF_LAMBDA_MAX);
    qmax = clipi(qmax, 1, FF_LAMBDA_MAX);
    if (qmax < qmin) qmax = qmin;

    /* ff_rate_estimate_qscale() */
    int64_t wanted_bits = (int64_t)(uint64_t)(bit_rate * (double)pts / fps);
    double diff = 0 /* total_bits */ - wanted_bits;
    float br_compensation = (bit_rate_tolerance - diff) / bit_rate_tolerance;
    if (br_compensation <= 0.0) br_compensation = 0.001;

    const float rce_qscale = FF_QP2LAMBDA * 2;
    const double pred_coeff = FF_QP2LAMBDA * 7.0, pred_count = 1.0;
    double bits = pred_coeff * sqrt((double)mb_var_sum) / (rce_qscale * pred_count); /* predict_size() */
    int i_tex_bits = (int)bits;
    const int p_tex_bits = 0;

    double rate_factor = 0.001 /* pass1_wanted_bits */ / 0.001 /* pass1_rc_eq_output_sum */ * br_compensation;

    /* get_qscale(): rc_eq = "tex^qComp" */
    double tex = (i_tex_bits + p_tex_bits) * (double)rce_qscale;
    bits = pow(tex, (double)qcompress);
    if (isnan(bits)) return -1;
    bits *= rate_factor;
    if (bits < 0.0) bits = 0.0;
    bits += 1.0;
    double qd = rce_qscale * (double)(i_tex_bits + p_tex_bits + 1) / bits; /* bits2qp() */
    qd = -qd * i_quant_factor + i_quant_offset;                            /* i_quant_factor < 0 */
    if (qd < 1) qd = 1;
    float q = qd;

    /* get_diff_limited_q(): nothing applies to the first I picture */
    q = (double)q;

    /* intra_only: short-term blur */
    double short_term_qsum = 0.001, short_term_qcount = 0.001;
    short_term_qsum *= qblur;
    short_term_qcount *= qblur;
    short_term_qsum += q;
    short_term_qcount++;
    q = short_term_qsum / short_term_qcount;

    /* modify_qscale(): no vbv buffer, rc_qsquish == 0 -> clip (in double) */
    double qm = q;
    if (qm < qmin) qm = qmin;
    else if (qm > qmax) qm = qmax;
    q = qm;

    /* back in ff_rate_estimate_qscale: clip again, then (no adaptive quant) q = (int)(q + 0.5) */
    if (q < qmin) q = qmin;
    else if (q > qmax) q = qmax;
    q = (int)(q + 0.5);

    /* estimate_qp(): int quality = q;  update_qscale() */
    int lambda = (int)q;
    int qscale = (lambda * 139 + FF_LAMBDA_SCALE * 64) >> (FF_LAMBDA_SHIFT + 7);
    qscale = clipi(qscale, avctx_qmin, avctx_qmax);
    if (lambda_out) *lambda_out = lambda;
    return qscale;
}

/* ------------------------------------------------------------------------------------------------ */
/* MJPEG matrices (encode_picture) + ff_convert_matrix (SIMD fdct branch)                           */
/* ------------------------------------------------------------------------------------------------ */
void orc_build_matrices(int qscale, uint8_t intra_matrix[64], uint16_t qmat16[64], uint16_t bias16[64])
{
    const int intra_quant_bias = 3 << (8 - 3); /* QUANT_BIAS_SHIFT 8; MJPEG shares the MPEG-2 bias */
    for (int i = 1; i < 64; i++) {
        int v = (mpeg1_default_intra_matrix[i] * qscale) >> 3;
        intra_matrix[i] = (uint8_t)(v > 255 ? 255 : v);
    }
    intra_matrix[0] = 8; /* ff_mpeg2_dc_scale_table[0][8] */
    /* ff_convert_matrix(..., qmin = qmax = 8, intra = 1): qscale2 = 8 << 1 */
    for (int i = 0; i < 64; i++) {
        int64_t den = (int64_t)16 * intra_matrix[i];
        int q16 = (int)((2 << 16) / den);
        if (q16 == 0 || q16 == 128 * 256) q16 = 128 * 256 - 1;
        qmat16[i] = (uint16_t)q16;
        int a = intra_quant_bias * (1 << (16 - 8));
        bias16[i] = (uint16_t)((a + (q16 >> 1)) / q16); /* ROUNDED_DIV, a >= 0 */
    }
}

/* dct_quantize_{sse2,ssse3} (mpegvideoenc_template.c), intra block of an MJPEG picture:
 *   DC:  level = ((block[0] >> 2) + q) * ff_inverse[q << 1] >> 32   with q = y/c_dc_scale = 8
 *   AC:  level = sign(x) * (pmulhw(paddusw(|x|, bias16), qmat16))                               */
int orc_quantize(const int16_t in[64], int16_t out_zz[64], const uint16_t qmat16[64], const uint16_t bias16[64])
{
    int16_t nat[64];
    const unsigned q = 8;
    const uint32_t inverse_2q = (uint32_t)(((1ULL << 32) + 2 * q - 1) / (2 * q)); /* ff_inverse[16] */
    nat[0] = (int16_t)(((uint64_t)(uint32_t)((in[0] >> 2) + (int)q) * inverse_2q) >> 32);
    for (int i = 1; i < 64; i++) {
        int x = in[i];
        uint16_t ax = (uint16_t)(x < 0 ? -x : x);                     /* pabsw / sign trick */
        unsigned s = (unsigned)ax + bias16[i];
        if (s > 65535) s = 65535;                                     /* paddusw */
        int16_t lv = (int16_t)(((int)(int16_t)s * (int)(int16_t)qmat16[i]) >> 16); /* pmulhw (signed) */
        nat[i] = (int16_t)(x < 0 ? -lv : lv);
    }
    int last = 0;
    for (int k = 0; k < 64; k++) {
        out_zz[k] = nat[zigzag_direct[k]];
        if (k && out_zz[k]) last = k;
    }
    return last;
}

/* ------------------------------------------------------------------------------------------------ */
/* AV_QSORT (libavutil/qsort.h) on (key, payload) pairs — unstable, order of equal keys matters      */
/* ------------------------------------------------------------------------------------------------ */
typedef struct { int a; int b; } pair_t; /* PTable {value, prob} / HuffTable {code, length} */
#define SWAPP(x, y) do { pair_t t_ = (x); (x) = (y); (y) = t_; } while (0)
static void av_qsort_pairs(pair_t *p, int num) /* cmp = x->b - y->b */
{
#define CMP(x, y) ((x)->b - (y)->b)
    pair_t *stack[64][2];
    int sp = 1;
    stack[0][0] = p;
    stack[0][1] = p + num - 1;
    while (sp) {
        pair_t *start = stack[--sp][0];
        pair_t *end = stack[sp][1];
        while (start < end) {
            if (start < end - 1) {
                int checksort = 0;
                pair_t *right = end - 2;
                pair_t *left = start + 1;
                pair_t *mid = start + ((end - start) >> 1);
                if (CMP(start, end) > 0) {
                    if (CMP(end, mid) > 0) SWAPP(*start, *mid);
                    else SWAPP(*start, *end);
                } else {
                    if (CMP(start, mid) > 0) SWAPP(*start, *mid);
                    else checksort = 1;
                }
                if (CMP(mid, end) > 0) {
                    SWAPP(*mid, *end);
                    checksort = 0;
                }
                if (start == end - 2) break;
                SWAPP(end[-1], *mid);
                while (left <= right) {
                    while (left <= right && CMP(left, end - 1) < 0) left++;
                    while (left <= right && CMP(right, end - 1) > 0) right--;
                    if (left <= right) {
                        SWAPP(*left, *right);
                        left++;
                        right--;
                    }
                }
                SWAPP(end[-1], *left);
                if (checksort && (mid == left - 1 || mid == left)) {
                    mid = start;
                    while (mid < end && CMP(mid, mid + 1) <= 0) mid++;
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
                if (CMP(start, end) > 0) SWAPP(*start, *end);
                break;
            }
        }
    }
#undef CMP
}

/* ------------------------------------------------------------------------------------------------ */
/* mjpegenc_huffman.c: package-merge length-limited code (max 16 bits)                              */
/* ------------------------------------------------------------------------------------------------ */
typedef struct {
    int nitems;
    int item_idx[515];
    int probability[514];
    int items[257 * 16];
} merger_list_t;

static void huffman_compute_bits(pair_t *prob_table /* {value, prob} */, pair_t *distincts /* {code, length} */,
                                 int size, int max_length)
{
    merger_list_t *la = calloc(1, sizeof *la), *lb = calloc(1, sizeof *lb);
    merger_list_t *to = la, *from = lb, *temp;
    int times, i = 0, j, k;
    int nbits[257] = {0};
    int min;

    to->nitems = 0;
    from->nitems = 0;
    to->item_idx[0] = 0;
    from->item_idx[0] = 0;
    av_qsort_pairs(prob_table, size);

    for (times = 0; times <= max_length; times++) {
        to->nitems = 0;
        to->item_idx[0] = 0;
        j = 0;
        k = 0;
        if (times < max_length) i = 0;
        while (i < size || j + 1 < from->nitems) {
            to->nitems++;
            to->item_idx[to->nitems] = to->item_idx[to->nitems - 1];
            if (i < size && (j + 1 >= from->nitems ||
                             prob_table[i].b < from->probability[j] + from->probability[j + 1])) {
                to->items[to->item_idx[to->nitems]++] = prob_table[i].a;
                to->probability[to->nitems - 1] = prob_table[i].b;
                i++;
            } else {
                for (k = from->item_idx[j]; k < from->item_idx[j + 2]; k++)
                    to->items[to->item_idx[to->nitems]++] = from->items[k];
                to->probability[to->nitems - 1] = from->probability[j] + from->probability[j + 1];
                j += 2;
            }
        }
        temp = to; to = from; from = temp;
    }
    min = (size - 1 < from->nitems) ? size - 1 : from->nitems;
    for (i = 0; i < from->item_idx[min]; i++) nbits[from->items[i]]++;
    j = 0;
    for (i = 0; i < 256; i++)
        if (nbits[i] > 0) {
            distincts[j].a = i;
            distincts[j].b = nbits[i];
            j++;
        }
    free(la);
    free(lb);
}

void orc_huffman_table(const uint32_t hist[256], uint8_t bits[17], uint8_t vals[256], int *nvals)
{
    pair_t val_counts[257], distincts[256];
    int nval = 0, j = 0;
    for (int i = 0; i < 256; i++)
        if (hist[i]) nval++;
    for (int i = 0; i < 256; i++)
        if (hist[i]) {
            val_counts[j].a = i;
            val_counts[j].b = (int)hist[i];
            j++;
        }
    val_counts[j].a = 256;
    val_counts[j].b = 0;
    huffman_compute_bits(val_counts, distincts, nval + 1, 16);
    av_qsort_pairs(distincts, nval);
    memset(bits, 0, 17);
    for (int i = 0; i < nval; i++) {
        vals[i] = (uint8_t)distincts[i].a;
        bits[distincts[i].b]++;
    }
    *nvals = nval;
}

/* ------------------------------------------------------------------------------------------------ */
/* libswscale range conversion yuv420p -> yuvj420p, same size (range_mode 1)                         */
/* ------------------------------------------------------------------------------------------------ */
/* swscale.c: the unscaled horizontal pass lifts 8-bit samples to 15 bits (hScale8To15_c with the
 * identity filter 1<<14: (src*16384)>>7 = src<<7), lumRangeToJpeg_c / chrRangeToJpeg_c run on that, and
 * the vertical pass is yuv2plane1_8_c.  Source and destination are both 8 bit, so the "dither" row is
 * the constant ff_sws_pb_64, i.e. plain round-to-nearest: the whole conversion is a 256-entry LUT per
 * plane kind.  Measured against the vendored libswscale.so.5 for SWS_POINT / SWS_BILINEAR / SWS_BICUBIC
 * (with and without SWS_ACCURATE_RND); SWS_FAST_BILINEAR takes a different (MMX) path and is not covered. */
static inline uint8_t clip_u8(int v) { return (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v)); }

static inline uint8_t range_luma_px(uint8_t s)
{
    int v = s << 7;
    if (v > 30189) v = 30189;
    v = (v * 19077 - 39057361) >> 14; /* lumRangeToJpeg_c */
    return clip_u8((v + 64) >> 7);
}
static inline uint8_t range_chroma_px(uint8_t s)
{
    int v = s << 7;
    if (v > 30775) v = 30775;
    v = (v * 4663 - 9289992) >> 12; /* chrRangeToJpeg_c */
    return clip_u8((v + 64) >> 7);
}

void orc_range_luma(const uint8_t *src, int sstride, uint8_t *dst, int dstride, int w, int h)
{
    for (int yy = 0; yy < h; yy++)
        for (int x = 0; x < w; x++) dst[(long)yy * dstride + x] = range_luma_px(src[(long)yy * sstride + x]);
}

void orc_range_chroma(const uint8_t *src, int sstride, uint8_t *dst, int dstride, int w, int h)
{
    for (int yy = 0; yy < h; yy++)
        for (int x = 0; x < w; x++) dst[(long)yy * dstride + x] = range_chroma_px(src[(long)yy * sstride + x]);
}

/* ------------------------------------------------------------------------------------------------ */
/* bit writer + JPEG container (mjpegenc_common.c)                                                  */
/* ------------------------------------------------------------------------------------------------ */
typedef struct {
    uint8_t *buf;
    long cap, pos; /* bytes */
    uint64_t acc;
    int nacc;      /* bits pending in acc (msb-first) */
    int overflow;
} bitw_t;
static void bw_byte(bitw_t *b, uint8_t v)
{
    if (b->pos < b->cap) b->buf[b->pos] = v;
    else b->overflow = 1;
    b->pos++;
}
static void bw_put(bitw_t *b, int n, uint32_t v)
{
    if (!n) return;
    b->acc = (b->acc << n) | (v & ((n == 32) ? 0xFFFFFFFFu : ((1u << n) - 1)));
    b->nacc += n;
    while (b->nacc >= 8) {
        bw_byte(b, (uint8_t)(b->acc >> (b->nacc - 8)));
        b->nacc -= 8;
    }
}
static void bw_marker(bitw_t *b, int code) { bw_put(b, 8, 0xff); bw_put(b, 8, code); }

enum { M_SOF0 = 0xc0, M_DHT = 0xc4, M_SOI = 0xd8, M_EOI = 0xd9, M_SOS = 0xda, M_DQT = 0xdb, M_COM = 0xfe };

static int put_huffman_table(bitw_t *b, int table_class, int table_id, const uint8_t *bits, const uint8_t *vals)
{
    int n = 0;
    bw_put(b, 4, table_class);
    bw_put(b, 4, table_id);
    for (int i = 1; i <= 16; i++) { n += bits[i]; bw_put(b, 8, bits[i]); }
    for (int i = 0; i < n; i++) bw_put(b, 8, vals[i]);
    return n + 17;
}

static void build_codes(uint8_t *huff_size, uint16_t *huff_code, const uint8_t *bits, const uint8_t *vals)
{
    int k = 0, code = 0;
    for (int i = 1; i <= 16; i++) {
        int nb = bits[i];
        for (int j = 0; j < nb; j++) {
            int sym = vals[k++];
            huff_size[sym] = (uint8_t)i;
            huff_code[sym] = (uint16_t)code;
            code++;
        }
        code <<= 1;
    }
}

static inline int log2_16(unsigned v) { int n = 0; while (v >>= 1) n++; return n; } /* av_log2_16bit */

/* ------------------------------------------------------------------------------------------------ */
/* whole frame                                                                                      */
/* ------------------------------------------------------------------------------------------------ */
typedef struct { uint8_t table_id, code; uint16_t mant; } huffsym_t; /* MJpegHuffmanCode */

static inline void rec_code(huffsym_t *hb, long *n, uint32_t hist[4][256], int table_id, int code, int mant)
{
    hb[*n].table_id = (uint8_t)table_id;
    hb[*n].code = (uint8_t)code;
    hb[*n].mant = (uint16_t)mant;
    (*n)++;
    hist[table_id][code]++;
}
static inline void rec_coef(huffsym_t *hb, long *n, uint32_t hist[4][256], int table_id, int val, int run)
{
    if (val == 0) {
        rec_code(hb, n, hist, table_id, 0, 0);
    } else {
        int mant = val;
        if (val < 0) { val = -val; mant--; }
        int code = (run << 4) | (log2_16((unsigned)val) + 1);
        rec_code(hb, n, hist, table_id, code, mant);
    }
}

long orc_encode_frame(const uint8_t *y, int ys, const uint8_t *u, int us, const uint8_t *v, int vs,
                      const orc_params *p, uint8_t *out, long cap, orc_debug *dbg)
{
    const int w = p->width, h = p->height;
    if (w <= 0 || h <= 0 || w > 65500 || h > 65500) return 0;
    const int fmt = p->chroma_format;
    if (fmt < 0 || fmt > 2) return 0;
    /* av_pix_fmt_get_chroma_sub_sample: yuvj420p (1,1), yuvj422p (1,0), yuvj444p (0,0) */
    const int hshift = fmt == ORC_CHROMA_444 ? 0 : 1, vshift = fmt == ORC_CHROMA_420 ? 1 : 0;
    const int cw = w >> hshift, ch = h >> vshift; /* what load_input_picture copies for planes 1/2 */
    if (cw <= 0 || ch <= 0) return 0;
    const int mbw = (w + 15) >> 4, mbh = (h + 15) >> 4;
    const char *comment = p->comment ? p->comment : "Lavc58.117.101";

    uint8_t *cy = NULL, *cu = NULL, *cv = NULL;
    if (p->range_mode == 1) { /* optional swscale-exact limited->full conversion in front */
        const int fcw = (w + (1 << hshift) - 1) >> hshift, fch = (h + (1 << vshift) - 1) >> vshift;
        cy = malloc((size_t)w * h); cu = malloc((size_t)fcw * fch); cv = malloc((size_t)fcw * fch);
        orc_range_luma(y, ys, cy, w, w, h);
        orc_range_chroma(u, us, cu, fcw, fcw, fch);
        orc_range_chroma(v, vs, cv, fcw, fcw, fch);
        y = cy; ys = w; u = cu; us = fcw; v = cv; vs = fcw;
    }

    int lambda = 0, qscale;
    int64_t var = orc_mb_var_sum(y, ys, w, h);
    if (p->fixed_qscale > 0) qscale = p->fixed_qscale;
    else qscale = orc_rate_control_qscale(var, p->pts, &lambda);

    uint8_t intra_matrix[64];
    uint16_t qmat16[64], bias16[64];
    orc_build_matrices(qscale, intra_matrix, qmat16, bias16);

    /* ---- encode_thread / encode_mb: record symbols for every MCU ------------------------------ */
    /* mpegvideo_enc.c encode_mb_internal numbers the blocks of a 16x16 macroblock: 0..3 luma (TL TR BL BR), then
     *   4:2:0  4 Cb, 5 Cr
     *   4:2:2  4 Cb top, 5 Cr top, 6 Cb bottom, 7 Cr bottom
     *   4:4:4  4 Cb TL, 5 Cr TL, 6 Cb TR, 7 Cr TR, 8 Cb BL, 9 Cr BL, 10 Cb BR, 11 Cr BR
     * and mjpegenc.c ff_mjpeg_encode_mb codes them in JPEG MCU order:
     *   4:2:0  0 1 2 3 4 5                     (16x16 MCU, Y 2x2, Cb 1x1, Cr 1x1)
     *   4:2:2  0 1 2 3 4 6 5 7                 (16x16 MCU, Y 2x2, Cb 1x2, Cr 1x2)
     *   4:4:4  0 2 4 8 5 9, then 1 3 6 10 7 11 if 16*mb_x+8 < width   (two 8x16 MCUs, every component 1x2) */
    static const int order420[6] = {0, 1, 2, 3, 4, 5};
    static const int order422[8] = {0, 1, 2, 3, 4, 6, 5, 7};
    static const int order444[12] = {0, 2, 4, 8, 5, 9, 1, 3, 6, 10, 7, 11};
    const int *order = fmt == ORC_CHROMA_420 ? order420 : (fmt == ORC_CHROMA_422 ? order422 : order444);
    const int per_mb = fmt == ORC_CHROMA_420 ? 6 : (fmt == ORC_CHROMA_422 ? 8 : 12);
    const long nblocks = (long)mbw * mbh * per_mb;
    huffsym_t *hb = malloc(sizeof(huffsym_t) * (size_t)nblocks * 64 + 64);
    uint32_t hist[4][256];
    memset(hist, 0, sizeof hist);
    long ncode = 0;
    int last_dc[3] = {128, 128, 128};
    long blk = 0;
    for (int my = 0; my < mbh; my++)
        for (int mx = 0; mx < mbw; mx++)
            for (int k = 0; k < per_mb; k++) {
                if (fmt == ORC_CHROMA_444 && k >= 6 && !(16 * mx + 8 < w)) break; /* no right half */
                const int n = order[k];
                int16_t b[64], zz[64];
                const uint8_t *pl; int st, pw, ph, bx, by;
                if (n < 4) { pl = y; st = ys; pw = w; ph = h; bx = mx * 16 + (n & 1) * 8; by = my * 16 + (n >> 1) * 8; }
                else {
                    pl = (n & 1) ? v : u; st = (n & 1) ? vs : us; pw = cw; ph = ch;
                    const int pos = (n - 4) >> 1; /* which chroma block of the component inside the macroblock */
                    if (fmt == ORC_CHROMA_420) { bx = mx * 8; by = my * 8; }
                    else if (fmt == ORC_CHROMA_422) { bx = mx * 8; by = my * 16 + pos * 8; }
                    else { bx = mx * 16 + (pos & 1) * 8; by = my * 16 + (pos >> 1) * 8; }
                }
                for (int r = 0; r < 8; r++)
                    for (int c = 0; c < 8; c++) b[r * 8 + c] = pad_px(pl, st, pw, ph, bx + c, by + r); /* get_pixels */
                orc_fdct_sse2(b);
                int last_index = orc_quantize(b, zz, qmat16, bias16);
                if (dbg && dbg->coefs) memcpy(dbg->coefs + blk * 64, zz, sizeof zz);
                blk++;
                /* record_block() */
                int component = (n <= 3 ? 0 : (n & 1) + 1);
                int table_id = (n <= 3 ? 0 : 1);
                int dc = zz[0];
                rec_coef(hb, &ncode, hist, table_id, dc - last_dc[component], 0);
                last_dc[component] = dc;
                int run = 0;
                table_id |= 2;
                for (int i = 1; i <= last_index; i++) {
                    int val = zz[i];
                    if (val == 0) run++;
                    else {
                        while (run >= 16) { rec_code(hb, &ncode, hist, table_id, 0xf0, 0); run -= 16; }
                        rec_coef(hb, &ncode, hist, table_id, val, run);
                        run = 0;
                    }
                }
                if (last_index < 63 || run != 0) rec_code(hb, &ncode, hist, table_id, 0, 0);
            }

    /* ---- ff_mjpeg_build_optimal_huffman ------------------------------------------------------- */
    uint8_t bits[4][17], vals[4][256];
    int nvals[4];
    uint8_t hsize[4][256];
    uint16_t hcode[4][256];
    memset(hsize, 0, sizeof hsize);
    memset(hcode, 0, sizeof hcode);
    memset(vals, 0, sizeof vals);
    for (int t = 0; t < 4; t++) {
        orc_huffman_table(hist[t], bits[t], vals[t], &nvals[t]);
        build_codes(hsize[t], hcode[t], bits[t], vals[t]);
    }

    /* ---- ff_mjpeg_encode_picture_header ------------------------------------------------------- */
    bitw_t bw = {out, cap, 0, 0, 0, 0};
    bw_marker(&bw, M_SOI);
    /* jpeg_put_comments: sample_aspect_ratio is 0/1 -> no JFIF APP0; BITEXACT not set -> COM */
    bw_marker(&bw, M_COM);
    bw_put(&bw, 16, (uint32_t)strlen(comment) + 3);
    for (const char *c = comment; *c; c++) bw_put(&bw, 8, (uint8_t)*c);
    bw_put(&bw, 8, 0);
    /* pix_fmt is YUVJ420P -> no "CS=ITU601" comment */
    /* jpeg_table_header: luma == chroma matrix -> one DQT table */
    bw_marker(&bw, M_DQT);
    bw_put(&bw, 16, 2 + 1 * (1 + 64));
    bw_put(&bw, 4, 0);
    bw_put(&bw, 4, 0);
    for (int i = 0; i < 64; i++) bw_put(&bw, 8, intra_matrix[zigzag_direct[i]]);
    bw_marker(&bw, M_DHT);
    {
        int size = 2;
        for (int t = 0; t < 4; t++) size += nvals[t] + 17;
        bw_put(&bw, 16, size);
        put_huffman_table(&bw, 0, 0, bits[0], vals[0]);
        put_huffman_table(&bw, 0, 1, bits[1], vals[1]);
        put_huffman_table(&bw, 1, 0, bits[2], vals[2]);
        put_huffman_table(&bw, 1, 1, bits[3], vals[3]);
    }
    bw_marker(&bw, M_SOF0);
    bw_put(&bw, 16, 17);
    bw_put(&bw, 8, 8);
    bw_put(&bw, 16, h);
    bw_put(&bw, 16, w);
    bw_put(&bw, 8, 3);
    /* mjpegenc_common.c ff_mjpeg_init_hvsample: 4:4:4 is coded with every component at h = 1, v = 2 (8x16 MCUs), the
     * others with luma at 2x2 and chroma at (2 >> chroma shift) */
    const int hs0 = fmt == ORC_CHROMA_444 ? 1 : 2, vs0 = 2;
    const int hsc = fmt == ORC_CHROMA_444 ? 1 : 2 >> hshift, vsc = fmt == ORC_CHROMA_444 ? 2 : 2 >> vshift;
    bw_put(&bw, 8, 1); bw_put(&bw, 4, hs0); bw_put(&bw, 4, vs0); bw_put(&bw, 8, 0);
    bw_put(&bw, 8, 2); bw_put(&bw, 4, hsc); bw_put(&bw, 4, vsc); bw_put(&bw, 8, 0);
    bw_put(&bw, 8, 3); bw_put(&bw, 4, hsc); bw_put(&bw, 4, vsc); bw_put(&bw, 8, 0);
    bw_marker(&bw, M_SOS);
    bw_put(&bw, 16, 12);
    bw_put(&bw, 8, 3);
    bw_put(&bw, 8, 1); bw_put(&bw, 4, 0); bw_put(&bw, 4, 0);
    bw_put(&bw, 8, 2); bw_put(&bw, 4, 1); bw_put(&bw, 4, 1);
    bw_put(&bw, 8, 3); bw_put(&bw, 4, 1); bw_put(&bw, 4, 1);
    bw_put(&bw, 8, 0);
    bw_put(&bw, 8, 63);
    bw_put(&bw, 8, 0);
    const long header_bytes = bw.pos;

    /* ---- ff_mjpeg_encode_picture_frame + ff_mjpeg_escape_FF (stuffing done on the fly) --------- */
    int64_t scan_bits = 0;
    uint64_t acc = 0;
    int nacc = 0;
    for (long i = 0; i < ncode; i++) {
        int t = hb[i].table_id, code = hb[i].code, nb = code & 0xf;
        int len = hsize[t][code];
        uint32_t word = hcode[t][code];
        if (nb) { word = (word << nb) | (hb[i].mant & ((1u << nb) - 1)); len += nb; } /* put_sbits */
        scan_bits += len;
        acc = (acc << len) | word;
        nacc += len;
        while (nacc >= 8) {
            uint8_t byte = (uint8_t)(acc >> (nacc - 8));
            nacc -= 8;
            bw_byte(&bw, byte);
            if (byte == 0xff) bw_byte(&bw, 0);
        }
    }
    if (nacc) { /* pad with ones */
        int pad = 8 - nacc;
        uint8_t byte = (uint8_t)(((acc << pad) | ((1u << pad) - 1)) & 0xff);
        bw_byte(&bw, byte);
        if (byte == 0xff) bw_byte(&bw, 0);
    }
    /* ff_mjpeg_encode_picture_trailer */
    bw_byte(&bw, 0xff);
    bw_byte(&bw, M_EOI);

    if (dbg) {
        dbg->qscale = qscale; dbg->lambda = lambda; dbg->mb_var_sum = var;
        dbg->mcu_w = fmt == ORC_CHROMA_444 ? (w + 7) >> 3 : mbw; dbg->mcu_h = mbh;
        memcpy(dbg->intra_matrix, intra_matrix, 64);
        memcpy(dbg->qmat16, qmat16, sizeof qmat16);
        memcpy(dbg->bias16, bias16, sizeof bias16);
        memcpy(dbg->hist, hist, sizeof hist);
        memcpy(dbg->bits, bits, sizeof bits);
        memcpy(dbg->vals, vals, sizeof vals);
        memcpy(dbg->nvals, nvals, sizeof nvals);
        dbg->scan_bits = scan_bits;
        dbg->header_bytes = (int)header_bytes;
    }
    free(hb);
    free(cy); free(cu); free(cv);
    return bw.overflow ? -bw.pos : bw.pos;
}

/* ------------------------------------------------------------------------------------------------ */
/* multi-threaded batch (CPU baseline "port" leg)                                                   */
/* ------------------------------------------------------------------------------------------------ */
typedef struct {
    const uint8_t *frames; long frame_stride; int n; const orc_params *p;
    uint8_t *out; long cap; long *sizes; int tid, nthreads;
} mt_job_t;
static void *mt_worker(void *arg)
{
    mt_job_t *j = arg;
    const int w = j->p->width, h = j->p->height, fcw = (w + 1) >> 1, fch = (h + 1) >> 1;
    for (int i = j->tid; i < j->n; i += j->nthreads) {
        const uint8_t *y = j->frames + (long)i * j->frame_stride;
        const uint8_t *u = y + (long)w * h, *v = u + (long)fcw * fch;
        j->sizes[i] = orc_encode_frame(y, w, u, fcw, v, fcw, j->p, j->out + (long)i * j->cap, j->cap, NULL);
    }
    return NULL;
}
int orc_encode_batch_mt(const uint8_t *frames, long frame_stride, int n, const orc_params *p,
                        uint8_t *out, long cap_per_frame, long *sizes, int threads)
{
    if (threads < 1) threads = 1;
    if (threads > 256) threads = 256;
    pthread_t th[256];
    mt_job_t jobs[256];
    for (int t = 0; t < threads; t++) {
        jobs[t] = (mt_job_t){frames, frame_stride, n, p, out, cap_per_frame, sizes, t, threads};
        if (pthread_create(&th[t], NULL, mt_worker, &jobs[t])) return -1;
    }
    for (int t = 0; t < threads; t++) pthread_join(th[t], NULL);
    return 0;
}

/* ------------------------------------------------------------------------------------------------ */
/* Independent check: baseline JPEG entropy DECODER (ITU T.81 F.2.2), used by the round-trip tests   */
/* to recover the quantised levels from a finished JPEG without going through any encoder code.      */
/* out: levels in zigzag order per block, MCU order Y0 Y1 Y2 Y3 Cb Cr (DC as level, not difference).  */
/* Returns the number of blocks decoded, or a negative error.  info = {w, h, n_ff00, scan_bytes}.     */
/* ------------------------------------------------------------------------------------------------ */
typedef struct { int mincode[18], maxcode[18], valptr[18]; uint8_t vals[256]; int present; } dec_tab_t;
typedef struct { const uint8_t *p, *end; uint32_t acc; int n; long ff00; int hit_marker; } dec_bits_t;

static int dec_bit(dec_bits_t *b)
{
    if (b->n == 0) {
        if (b->p >= b->end) return -1;
        uint8_t c = *b->p++;
        if (c == 0xff) {
            if (b->p < b->end && *b->p == 0x00) { b->p++; b->ff00++; }
            else { b->hit_marker = 1; return -1; }
        }
        b->acc = c;
        b->n = 8;
    }
    b->n--;
    return (b->acc >> b->n) & 1;
}
static int dec_sym(dec_bits_t *b, const dec_tab_t *t)
{
    int code = 0;
    for (int l = 1; l <= 16; l++) {
        int bit = dec_bit(b);
        if (bit < 0) return -1;
        code = (code << 1) | bit;
        if (t->maxcode[l] >= 0 && code <= t->maxcode[l] && code >= t->mincode[l]) return t->vals[t->valptr[l] + code - t->mincode[l]];
    }
    return -2;
}
static int dec_receive_extend(dec_bits_t *b, int s)
{
    int v = 0;
    for (int i = 0; i < s; i++) { int bit = dec_bit(b); if (bit < 0) return 0x7fffffff; v = (v << 1) | bit; }
    if (s && v < (1 << (s - 1))) v -= (1 << s) - 1;
    return v;
}

long orc_jpeg_decode_coefs(const uint8_t *jpeg, long n, int16_t *out, long cap_blocks, long *info)
{
    dec_tab_t tabs[2][2];
    memset(tabs, 0, sizeof tabs);
    long i = 2;
    int w = 0, h = 0;
    int hv[3][2] = {{0, 0}, {0, 0}, {0, 0}}; /* sampling factors (h, v) of the three components */
    if (n < 4 || jpeg[0] != 0xff || jpeg[1] != 0xd8) return -1;
    while (i + 4 <= n) {
        if (jpeg[i] != 0xff) return -2;
        int m = jpeg[i + 1];
        long L = (jpeg[i + 2] << 8) | jpeg[i + 3];
        const uint8_t *seg = jpeg + i + 4;
        if (m == 0xc4) {
            long k = 0;
            while (k < L - 2) {
                int tc = seg[k] >> 4, th = seg[k] & 15;
                if (tc > 1 || th > 1) return -3;
                dec_tab_t *t = &tabs[tc][th];
                const uint8_t *bits = seg + k + 1;
                int total = 0, code = 0;
                for (int l = 1; l <= 16; l++) {
                    t->valptr[l] = total;
                    t->mincode[l] = code;
                    total += bits[l - 1];
                    code += bits[l - 1];
                    t->maxcode[l] = bits[l - 1] ? code - 1 : -1;
                    code <<= 1;
                }
                memcpy(t->vals, seg + k + 17, total);
                t->present = 1;
                k += 17 + total;
            }
        } else if (m == 0xc0) {
            h = (seg[1] << 8) | seg[2];
            w = (seg[3] << 8) | seg[4];
            if (seg[5] != 3) return -4;
            for (int c = 0; c < 3; c++) { hv[c][0] = seg[7 + 3 * c] >> 4; hv[c][1] = seg[7 + 3 * c] & 15; }
            if (hv[1][0] != hv[2][0] || hv[1][1] != hv[2][1] || hv[0][1] != 2) return -4;
        } else if (m == 0xda) {
            i += 2 + L;
            break;
        }
        i += 2 + L;
    }
    if (!w || !h) return -5;
    /* blocks of an MCU in scan order: h*v of every component; MCU = (8 * hmax) x (8 * vmax) pixels */
    const int per_mcu = hv[0][0] * hv[0][1] + 2 * hv[1][0] * hv[1][1];
    const int mbw = (w + 8 * hv[0][0] - 1) / (8 * hv[0][0]), mbh = (h + 15) >> 4;
    const long nblk = (long)mbw * mbh * per_mcu;
    if (per_mcu <= 0 || per_mcu > 12 || nblk > cap_blocks) return -6;
    const int n_luma = hv[0][0] * hv[0][1], n_cb = hv[1][0] * hv[1][1];
    dec_bits_t b = {jpeg + i, jpeg + n, 0, 0, 0, 0};
    int pred[3] = {128, 128, 128}; /* the encoder's last_dc starts at 128 (level shift folded into DC) */
    for (long blk = 0; blk < nblk; blk++) {
        const int nn = (int)(blk % per_mcu), comp = nn < n_luma ? 0 : (nn < n_luma + n_cb ? 1 : 2), th = comp ? 1 : 0;
        int16_t *o = out + blk * 64;
        memset(o, 0, 128);
        int s = dec_sym(&b, &tabs[0][th]);
        if (s < 0) return -7;
        int diff = dec_receive_extend(&b, s);
        if (diff == 0x7fffffff) return -8;
        pred[comp] += diff;
        o[0] = (int16_t)pred[comp];
        for (int k = 1; k < 64;) {
            int rs = dec_sym(&b, &tabs[1][th]);
            if (rs < 0) return -9;
            int r = rs >> 4, sz = rs & 15;
            if (sz == 0) {
                if (r == 15) { k += 16; continue; }
                break; /* EOB */
            }
            k += r;
            if (k > 63) return -10;
            int v = dec_receive_extend(&b, sz);
            if (v == 0x7fffffff) return -11;
            o[k++] = (int16_t)v;
        }
    }
    /* remaining bits must be 1-padding, then EOI */
    while (b.n > 0) { if (dec_bit(&b) != 1) return -12; }
    if (b.p + 2 > b.end || b.p[0] != 0xff || b.p[1] != 0xd9) return -13;
    if (info) { info[0] = w; info[1] = h; info[2] = b.ff00; info[3] = (long)(b.p - (jpeg + i)); }
    return nblk;
}
