// TEST INFRASTRUCTURE — not product code.
//
// Thin C-ABI harness around the UNMODIFIED reference, compiled by oracle/Makefile from
// the sources where they lie under /root/reference (src/Encoder.cpp, src/Decoder.cpp)
// and linked against the ffmpeg shared libraries the reference vendors
// (lib/ffmpeg/x86_64_shared, FFmpeg git-2021-01-28-6fd0116, libavcodec 58.117.101).
// The output goes to oracle/_ref/libh2j_ref.so (git-ignored).  It exists so that
//   * the C restatement in oracle/mjpeg_oracle.c can be pinned against the real thing,
//   * golden fixtures under tests/golden/ can be generated (tests/golden/make_golden.py),
//   * bench.py --impl reference / cpu_baseline(kind="reference") can time the real
//     reference YUV->JPEG stage (Encoder::yuv2Jpeg, reference src/Encoder.cpp:104).
// Nothing in the product path links or loads this file.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <unistd.h>

#include "Encoder.h"   // reference src/Encoder.h (pulls the vendored ffmpeg headers)
#include "IDecoder.h"  // reference export_inc/IDecoder.h

extern "C" {
#include "libavcodec/avdct.h"
#include "libswscale/swscale.h"
#include "libavutil/imgutils.h"
#include "libavutil/pixdesc.h"
}

namespace {
// The reference logs through printf on every call (src/Decoder.cpp:22 LOG()).  For timing
// runs we silence stdout around the call so the terminal is not the thing being measured.
struct StdoutSilencer {
    int saved = -1;
    explicit StdoutSilencer(bool on) {
        if (!on) return;
        fflush(stdout);
        saved = dup(1);
        FILE *n = fopen("/dev/null", "w");
        if (n) { dup2(fileno(n), 1); fclose(n); }
    }
    ~StdoutSilencer() {
        if (saved < 0) return;
        fflush(stdout);
        dup2(saved, 1);
        close(saved);
    }
};

AVFrame *make_frame(const uint8_t *y, int ys, const uint8_t *u, int us, const uint8_t *v, int vs,
                    int w, int h, int64_t pts, int pix_fmt) {
    AVFrame *f = av_frame_alloc();
    if (!f) return nullptr;
    f->width = w;
    f->height = h;
    f->format = pix_fmt;  // what the decoder would have produced (yuv420p = 0, yuvj420p = 12)
    f->pts = pts;
    if (av_frame_get_buffer(f, 32) < 0) { av_frame_free(&f); return nullptr; }
    const int cw = (w + 1) >> 1, ch = (h + 1) >> 1;
    for (int r = 0; r < h; r++) memcpy(f->data[0] + (size_t)r * f->linesize[0], y + (size_t)r * ys, w);
    for (int r = 0; r < ch; r++) {
        memcpy(f->data[1] + (size_t)r * f->linesize[1], u + (size_t)r * us, cw);
        memcpy(f->data[2] + (size_t)r * f->linesize[2], v + (size_t)r * vs, cw);
    }
    return f;
}
}  // namespace

extern "C" {

// Reference YUV->JPEG stage: Encoder(out).yuv2Jpeg(frame)  (reference src/Encoder.cpp:104-308).
// Planes are tightly described by (ptr, stride); chroma planes are ceil(w/2) x ceil(h/2).
// Returns 1 on success (the reference's bool), 0 on failure.
int ref_yuv2jpeg_file(const uint8_t *y, int ys, const uint8_t *u, int us, const uint8_t *v, int vs,
                      int w, int h, int64_t pts, int pix_fmt, const char *out_path, int quiet) {
    AVFrame *f = make_frame(y, ys, u, us, v, vs, w, h, pts, pix_fmt);
    if (!f) return 0;
    bool ok;
    {
        StdoutSilencer s(quiet != 0);
        ok = Encoder(out_path).yuv2Jpeg(f);
    }
    av_frame_free(&f);
    return ok ? 1 : 0;
}

// Same, but hands the bytes back in memory (goes through a tmpfs file because the reference
// only knows how to write files).  Returns the JPEG size, 0 on failure, -needed if cap is short.
long ref_yuv2jpeg(const uint8_t *y, int ys, const uint8_t *u, int us, const uint8_t *v, int vs,
                  int w, int h, int64_t pts, int pix_fmt, uint8_t *out, long cap) {
    char path[128];
    snprintf(path, sizeof path, "/dev/shm/h2j_ref_%d_%p.jpg", (int)getpid(), (void *)out);
    if (!ref_yuv2jpeg_file(y, ys, u, us, v, vs, w, h, pts, pix_fmt, path, 1)) { unlink(path); return 0; }
    FILE *fp = fopen(path, "rb");
    if (!fp) return 0;
    fseek(fp, 0, SEEK_END);
    long n = ftell(fp);
    fseek(fp, 0, SEEK_SET);
    long ret = n;
    if (n > cap) ret = -n;
    else if (fread(out, 1, n, fp) != (size_t)n) ret = 0;
    fclose(fp);
    unlink(path);
    return ret;
}

// Whole reference path: IDecoder::getInstance()->H265ToJpeg(in, out) (reference src/Decoder.cpp:120).
int ref_h265_to_jpeg(const char *in_path, const char *out_path, int quiet) {
    StdoutSilencer s(quiet != 0);
    auto d = IDecoder::getInstance();
    if (!d) return 0;
    return d->H265ToJpeg(in_path, out_path) ? 1 : 0;
}

// Decode the first frame exactly the way reference src/Decoder.cpp:138-330 does and hand the planes
// back, so tests can feed the very same decoded frame to the reference encoder, the oracle and the GPU.
// planes: caller buffers of at least w*h, cw*ch, cw*ch bytes (tight strides).  info = {w,h,fmt,pts_lo,pts_hi}.
int ref_decode_first_frame(const char *in_path, uint8_t *y, uint8_t *u, uint8_t *v, long cap_y, int64_t *info) {
    AVFormatContext *fmt = nullptr;
    AVCodecContext *cc = nullptr;
    AVFrame *fr = nullptr;
    AVPacket *pkt = nullptr;
    int ok = 0;
    do {
        if (avformat_open_input(&fmt, in_path, nullptr, nullptr) < 0) break;
        if (avformat_find_stream_info(fmt, nullptr) < 0) break;
        int st = av_find_best_stream(fmt, AVMEDIA_TYPE_VIDEO, -1, -1, nullptr, 0);
        if (st < 0) break;
        AVCodecParameters *par = fmt->streams[st]->codecpar;
        AVCodec *codec = avcodec_find_decoder(par->codec_id);
        if (!codec) break;
        cc = avcodec_alloc_context3(codec);
        if (!cc || avcodec_parameters_to_context(cc, par) < 0) break;
        if (avcodec_open2(cc, codec, nullptr) < 0) break;
        fr = av_frame_alloc();
        pkt = av_packet_alloc();
        if (!fr || !pkt) break;
        av_init_packet(pkt);
        pkt->data = nullptr;
        pkt->size = 0;
        if (av_read_frame(fmt, pkt) < 0) break;
        if (pkt->stream_index != st) break;
        int r = avcodec_send_packet(cc, pkt);
        av_packet_unref(pkt);
        if (r < 0) break;
        if (avcodec_receive_frame(cc, fr) != 0) break;
        const int w = fr->width, h = fr->height, cw = (w + 1) >> 1, ch = (h + 1) >> 1;
        info[0] = w; info[1] = h; info[2] = fr->format; info[3] = fr->pts;
        info[4] = fr->linesize[0]; info[5] = fr->linesize[1];
        if ((long)w * h > cap_y) { ok = -1; break; }
        for (int i = 0; i < h; i++) memcpy(y + (size_t)i * w, fr->data[0] + (size_t)i * fr->linesize[0], w);
        for (int i = 0; i < ch; i++) {
            memcpy(u + (size_t)i * cw, fr->data[1] + (size_t)i * fr->linesize[1], cw);
            memcpy(v + (size_t)i * cw, fr->data[2] + (size_t)i * fr->linesize[2], cw);
        }
        ok = 1;
    } while (0);
    if (fmt) avformat_close_input(&fmt);
    if (cc) avcodec_free_context(&cc);
    if (fr) av_frame_free(&fr);
    if (pkt) av_packet_free(&pkt);
    return ok;
}

// The forward DCT libavcodec selects at run time for dct_algo=FF_DCT_AUTO on this CPU
// (public AVDCT API, libavcodec/avdct.h).  In place on a 16-byte aligned int16[64].
int ref_fdct(int16_t *block, int n_blocks) {
    static AVDCT *d = nullptr;
    if (!d) {
        d = avcodec_dct_alloc();
        if (!d || avcodec_dct_init(d) < 0) return 0;
    }
    alignas(16) int16_t tmp[64];
    for (int i = 0; i < n_blocks; i++) {
        memcpy(tmp, block + 64 * i, sizeof tmp);
        d->fdct(tmp);
        memcpy(block + 64 * i, tmp, sizeof tmp);
    }
    return 1;
}

// libswscale yuv420p(limited) -> yuvj420p(full) at the same size, default flags the way a
// `sws_getContext(w,h,YUV420P,w,h,YUVJ420P,SWS_BILINEAR,...)` caller would get it.
int ref_sws_limited_to_full(const uint8_t *y, const uint8_t *u, const uint8_t *v, int w, int h, int flags,
                            uint8_t *oy, uint8_t *ou, uint8_t *ov) {
    const int cw = (w + 1) >> 1;
    SwsContext *c = sws_getContext(w, h, AV_PIX_FMT_YUV420P, w, h, AV_PIX_FMT_YUVJ420P, flags, nullptr, nullptr, nullptr);
    if (!c) return 0;
    const uint8_t *src[4] = {y, u, v, nullptr};
    int sst[4] = {w, cw, cw, 0};
    uint8_t *dst[4] = {oy, ou, ov, nullptr};
    int dstst[4] = {w, cw, cw, 0};
    int r = sws_scale(c, src, sst, 0, h, dst, dstst);
    sws_freeContext(c);
    return r == h ? 1 : 0;
}

// N same-sized tight I420 frames (frame_stride bytes apart), `threads` host threads, each thread running the
// reference's Encoder::yuv2Jpeg on frames tid, tid+threads, ... — one fresh Encoder per frame, exactly as
// reference src/Decoder.cpp:349 does — writing to tmpfs.  sizes[i] receives each JPEG's size (0 on failure).
// Returns the number of frames that encoded.
}  // extern "C"
#include <pthread.h>
#include <sys/stat.h>
#include <time.h>
namespace {
struct MtJob {
    const uint8_t *frames; long stride; int n, w, h, tid, nthreads; long *sizes; int ok;
};
void *mt_worker(void *arg)
{
    MtJob *j = static_cast<MtJob *>(arg);
    const int w = j->w, h = j->h, cw = (w + 1) >> 1, ch = (h + 1) >> 1;
    char path[128];
    snprintf(path, sizeof path, "/dev/shm/h2j_refmt_%d_%d.jpg", (int)getpid(), j->tid);
    for (int i = j->tid; i < j->n; i += j->nthreads) {
        const uint8_t *y = j->frames + (long)i * j->stride;
        const uint8_t *u = y + (long)w * h, *v = u + (long)cw * ch;
        AVFrame *f = make_frame(y, w, u, cw, v, cw, w, h, AV_NOPTS_VALUE, AV_PIX_FMT_YUV420P);
        if (!f) continue;
        const bool ok = Encoder(path).yuv2Jpeg(f);
        av_frame_free(&f);
        long sz = 0;
        if (ok) {
            struct stat st;
            if (stat(path, &st) == 0) sz = (long)st.st_size;
            j->ok++;
        }
        if (j->sizes) j->sizes[i] = sz;
    }
    unlink(path);
    return nullptr;
}
}  // namespace
extern "C" {
int ref_yuv2jpeg_batch_mt(const uint8_t *frames, long frame_stride, int n, int w, int h, long *sizes, int threads)
{
    if (threads < 1) threads = 1;
    if (threads > 512) threads = 512;
    StdoutSilencer s(true);
    pthread_t th[512];
    MtJob jobs[512];
    for (int t = 0; t < threads; t++) {
        jobs[t] = MtJob{frames, frame_stride, n, w, h, t, threads, sizes, 0};
        pthread_create(&th[t], nullptr, mt_worker, &jobs[t]);
    }
    int ok = 0;
    for (int t = 0; t < threads; t++) { pthread_join(th[t], nullptr); ok += jobs[t].ok; }
    return ok;
}

// The reference's own harness loop (main.cpp:37-65), timed: n_calls x IDecoder::getInstance()->H265ToJpeg(in, out) dealt to
// `threads` caller threads; call i writes <out_prefix><i>.jpeg.  Returns the number of successful calls.
int ref_h265_loop(const char *in_path, const char *out_prefix, int n_calls, int threads, int quiet, double *seconds)
{
    if (threads < 1) threads = 1;
    if (threads > 512) threads = 512;
    StdoutSilencer s(quiet != 0);
    struct LoopJob { const char *in, *prefix; int n, tid, nthreads, ok; };
    LoopJob jobs[512];
    pthread_t th[512];
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    for (int t = 0; t < threads; t++) {
        jobs[t] = LoopJob{in_path, out_prefix, n_calls, t, threads, 0};
        pthread_create(&th[t], nullptr, [](void *arg) -> void * {
            LoopJob *j = static_cast<LoopJob *>(arg);
            char out[512];
            for (int i = j->tid; i < j->n; i += j->nthreads) {
                snprintf(out, sizeof out, "%s%d.jpeg", j->prefix, i);
                auto d = IDecoder::getInstance();
                if (d && d->H265ToJpeg(j->in, out)) j->ok++;
            }
            return nullptr;
        }, &jobs[t]);
    }
    int ok = 0;
    for (int t = 0; t < threads; t++) { pthread_join(th[t], nullptr); ok += jobs[t].ok; }
    clock_gettime(CLOCK_MONOTONIC, &t1);
    if (seconds) *seconds = (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
    return ok;
}

// libavcodec's mjpeg encoder opened the way reference src/Encoder.cpp:158-204 opens it -- codec parameters (codec id, type,
// format, size) -> context, time_base 1/25, every other option at its default, ONE frame sent, one packet received -- but for a
// pixel format of the caller's choice.  The reference hard-codes AV_PIX_FMT_YUVJ420P (src/Encoder.cpp:162); this entry pins the
// oracle's 4:2:2 / 4:4:4 modes (the same encoder at other MCU geometries) against the same library.  The
// raw "mjpeg" muxer the reference writes through passes the packet bytes on unchanged, so the packet IS the file.
// planes: y w x h; u, v (w >> hshift) x (h >> vshift) rounded up, tight or strided.  Returns the size, 0 on failure, -needed.
long ref_mjpeg_encode_fmt(const uint8_t *y, int ys, const uint8_t *u, int us, const uint8_t *v, int vs, int w, int h, int pix_fmt,
                          uint8_t *out, long cap)
{
    StdoutSilencer s(true);
    AVCodec *codec = avcodec_find_encoder(AV_CODEC_ID_MJPEG);
    if (!codec) return 0;
    AVCodecParameters *par = avcodec_parameters_alloc();
    AVCodecContext *cc = avcodec_alloc_context3(codec);
    AVFrame *f = av_frame_alloc();
    AVPacket *pkt = av_packet_alloc();
    long ret = 0;
    do {
        if (!par || !cc || !f || !pkt) break;
        par->codec_id = AV_CODEC_ID_MJPEG;
        par->codec_type = AVMEDIA_TYPE_VIDEO;
        par->format = pix_fmt;
        par->width = w;
        par->height = h;
        if (avcodec_parameters_to_context(cc, par) < 0) break;
        cc->time_base = (AVRational){1, 25};
        if (avcodec_open2(cc, codec, nullptr) < 0) break;
        f->width = w;
        f->height = h;
        f->format = pix_fmt;
        f->pts = AV_NOPTS_VALUE;
        if (av_frame_get_buffer(f, 32) < 0) break;
        int hs = 0, vsft = 0;
        av_pix_fmt_get_chroma_sub_sample((AVPixelFormat)pix_fmt, &hs, &vsft);
        const int cw = (w + (1 << hs) - 1) >> hs, ch = (h + (1 << vsft) - 1) >> vsft;
        for (int r = 0; r < h; r++) memcpy(f->data[0] + (size_t)r * f->linesize[0], y + (size_t)r * ys, w);
        for (int r = 0; r < ch; r++) {
            memcpy(f->data[1] + (size_t)r * f->linesize[1], u + (size_t)r * us, cw);
            memcpy(f->data[2] + (size_t)r * f->linesize[2], v + (size_t)r * vs, cw);
        }
        if (avcodec_send_frame(cc, f) < 0) break;
        if (avcodec_receive_packet(cc, pkt) < 0) break;
        ret = pkt->size > cap ? -(long)pkt->size : pkt->size;
        if (ret > 0) memcpy(out, pkt->data, pkt->size);
    } while (0);
    if (pkt) av_packet_free(&pkt);
    if (f) av_frame_free(&f);
    if (cc) avcodec_free_context(&cc);
    if (par) avcodec_parameters_free(&par);
    return ret;
}

const char *ref_version(void) { return LIBAVCODEC_IDENT; }

}  // extern "C"
