/* TEST INFRASTRUCTURE — not product code.  See mjpeg_oracle.c. */
#ifndef H2J_MJPEG_ORACLE_H
#define H2J_MJPEG_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_NOPTS ((int64_t)0x8000000000000000ULL) /* AV_NOPTS_VALUE */

typedef struct orc_params {
    int width, height;
    int64_t pts;         /* AVFrame.pts the decoder produced; ORC_NOPTS for raw .h264/.h265 input */
    int fixed_qscale;    /* 0: the reference's rate control picks it; 1..31: force it */
    int range_mode;      /* 0: planes used as they are (what reference src/Encoder.cpp does);
                            1: swscale-exact yuv420p(limited) -> yuvj420p(full) first */
    const char *comment; /* COM payload; NULL = "Lavc58.117.101" (the x86_64_shared build) */
    int chroma_format;   /* 0: 4:2:0 (yuvj420p, what reference src/Encoder.cpp:162 opens); 1: 4:2:2 (yuvj422p); 2: 4:4:4
                            (yuvj444p) -- the same libavcodec encoder at its other MCU geometries */
} orc_params;

#define ORC_CHROMA_420 0
#define ORC_CHROMA_422 1
#define ORC_CHROMA_444 2

typedef struct orc_debug {
    int qscale;
    int lambda;
    int64_t mb_var_sum;
    int mcu_w, mcu_h;
    uint8_t intra_matrix[64];   /* natural (raster) order, [0] = 8 */
    uint16_t qmat16[64];        /* natural order */
    uint16_t bias16[64];
    uint32_t hist[4][256];      /* 0 DC luma, 1 DC chroma, 2 AC luma, 3 AC chroma */
    uint8_t bits[4][17];
    uint8_t vals[4][256];
    int nvals[4];
    int64_t scan_bits;          /* entropy coded bits before the 1-padding */
    int header_bytes;           /* SOI .. end of SOS */
    int16_t *coefs;             /* optional, caller-allocated [mcu_w*mcu_h*blocks_per_mcu*64]: quantised levels in
                                   ZIGZAG order per block, blocks in coding order (4:2:0: Y0 Y1 Y2 Y3 Cb Cr per 16x16 MCU;
                                   4:2:2: Y0 Y1 Y2 Y3 Cb0 Cb1 Cr0 Cr1 per 16x16 MCU; 4:4:4: Y0 Y1 Cb0 Cb1 Cr0 Cr1 per 8x16 MCU) */
} orc_debug;

/* ff_fdct_sse2 restated (in place, natural order). */
void orc_fdct_sse2(int16_t blk[64]);
/* jpeg_fdct_islow_8 restated (dct_algo=FF_DCT_INT; not what the reference selects on x86). */
void orc_fdct_islow(int16_t blk[64]);

int64_t orc_mb_var_sum(const uint8_t *y, int ystride, int w, int h);
/* first-frame rate control: returns qscale, writes lambda */
int orc_rate_control_qscale(int64_t mb_var_sum, int64_t pts, int *lambda_out);
void orc_build_matrices(int qscale, uint8_t intra_matrix[64], uint16_t qmat16[64], uint16_t bias16[64]);
/* dct_quantize (SSE2/SSSE3 template) restated: in = fdct output (natural), out = levels in zigzag order.
 * returns last non-zero zigzag index (0 if only DC). */
int orc_quantize(const int16_t in[64], int16_t out_zz[64], const uint16_t qmat16[64], const uint16_t bias16[64]);

/* optimal huffman table from a 256-bin histogram (mjpegenc_huffman.c restated) */
void orc_huffman_table(const uint32_t hist[256], uint8_t bits[17], uint8_t vals[256], int *nvals);

/* swscale yuv420p -> yuvj420p (same size) restated, one plane at a time */
void orc_range_luma(const uint8_t *src, int sstride, uint8_t *dst, int dstride, int w, int h);
void orc_range_chroma(const uint8_t *src, int sstride, uint8_t *dst, int dstride, int w, int h);

/* Whole frame: returns JPEG size in bytes, or -(needed) if cap is too small, 0 on bad arguments. */
long orc_encode_frame(const uint8_t *y, int ys, const uint8_t *u, int us, const uint8_t *v, int vs,
                      const orc_params *p, uint8_t *out, long cap, orc_debug *dbg);

/* Baseline JPEG entropy decoder (independent of the encoder restatement): quantised levels back out of a
 * finished JPEG, zigzag order, MCU order.  Returns blocks decoded or <0.  info = {w, h, n_ff00, scan_bytes}. */
long orc_jpeg_decode_coefs(const uint8_t *jpeg, long n, int16_t *out, long cap_blocks, long *info);

/* N same-sized frames laid out back to back (frame stride in bytes given), T threads.
 * out: cap_per_frame bytes per frame; sizes[i] receives each size.  Returns 0 on success. */
int orc_encode_batch_mt(const uint8_t *frames, long frame_stride, int n, const orc_params *p,
                        uint8_t *out, long cap_per_frame, long *sizes, int threads);

#ifdef __cplusplus
}
#endif
#endif
