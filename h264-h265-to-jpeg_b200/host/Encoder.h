// Host-side mirror of the reference's encoder class (reference src/Encoder.h:33-96) over the h2j_b200 C ABI.
//
// Same name, same constructor, same yuv2Jpeg(AVFrame*) -> bool contract and the same observable behaviour:
// on success the JPEG is written to outputFilePath with fopen(..., "wb+") (reference src/Encoder.cpp:336-361),
// on any failure a line goes to LOG() and false is returned.  The bytes written are identical to the
// reference's.  Dropping this header + Encoder.cpp in place of the reference's src/Encoder.{h,cpp} leaves
// src/Decoder.cpp (the only caller, Decoder.cpp:319), export_inc/IDecoder.h and src/jni untouched.
//
// The libavcodec/libavformat machinery the reference's Encoder drove (AVIOContext, AVFormatContext, mjpeg
// AVCodecContext ...) is gone: the planes go to the GPU through pinned staging and asynchronous copies.
#ifndef H2J_HOST_ENCODER_H
#define H2J_HOST_ENCODER_H

#include <cstdint>
#include <memory>

struct AVFrame;  // libavutil/frame.h — only needed by Encoder.cpp, and only when built with H2J_WITH_LIBAV

/* 堆缓冲大小 / 栈缓冲大小 of the reference (src/Common.h:15,18), kept for source compatibility */
#ifndef HEAP_SIZE
#define HEAP_SIZE (1024 * 1024 * 2)
#endif
#ifndef STACK_SIZE
#define STACK_SIZE (1024)
#endif

// What yuv2Jpeg needs from an AVFrame, for callers (and tests) that have no libavutil headers.
struct H2JFrameView {
    const uint8_t *data[3];
    int linesize[3];
    int width, height;
    int format;  // AVPixelFormat: 0 = AV_PIX_FMT_YUV420P, 12 = AV_PIX_FMT_YUVJ420P
};

class Encoder {
public:
    explicit Encoder(const char *outputFilePath);
    ~Encoder();

    // reference src/Encoder.cpp:89
    bool yuv2Jpeg(AVFrame *pFrame);
    // same, from plain pointers
    bool yuv2Jpeg(const H2JFrameView &frame);

    // Bytes of the last JPEG produced (the reference keeps them in Output::jpeg_data, src/Common.h:61).
    const uint8_t *jpegData() const { return jpeg_.get(); }
    size_t jpegSize() const { return jpegSize_; }

private:
    void release();
    bool saveJpegtoFile(const char *filePath);

    const char *outputFilePath;
    std::unique_ptr<uint8_t[]> jpeg_;
    size_t jpegSize_ = 0;
    size_t jpegCap_ = 0;
};

// Process-wide knobs (optional; defaults reproduce the reference): CUDA device ordinal and whether to convert
// limited-range input to full range first (the reference does not).  Call before the first yuv2Jpeg.
extern "C" void h2j_host_configure(int cuda_device, int range_mode);

// Batch scope for services that convert many pictures (the reference has no such call: a loop over
// IDecoder::H265ToJpeg is what it offers).  Between begin and end, yuv2Jpeg() -- and therefore the unchanged
// IDecoder::H265ToJpeg() that calls it (reference src/Decoder.cpp:319) -- only copies the decoded planes into pinned
// staging and returns true; the GPU encodes the queued pictures `max_frames` at a time (runs of equal size as one
// batch) and the JPEG files are written then.  h2j_host_batch_end() drains the queue and returns the number of files
// written (-1 outside a scope); *failed receives the number of pictures that could not be encoded or saved.
extern "C" int h2j_host_batch_begin(int max_frames);
extern "C" int h2j_host_batch_end(int *failed);

#endif  // H2J_HOST_ENCODER_H
