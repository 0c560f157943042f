// See Encoder.h.  Mirrors reference src/Encoder.cpp's control flow and error behaviour; the encode itself is
// one call into the C ABI (include/h2j_b200.h), which runs the CUDA kernels.  There is no CPU encode here:
// if the GPU library cannot be used, yuv2Jpeg logs and returns false.
#include "Encoder.h"

#include <cerrno>
#include <cstdio>
#include <cstring>
#include <mutex>

#include "../../include/h2j_b200.h"

#ifdef H2J_WITH_LIBAV
extern "C" {
#include "libavutil/frame.h"
}
#endif

extern void LOG(const char *format, ...);  // reference src/Decoder.cpp:22 (host/log_default.cpp when built alone)

namespace {
// One GPU context per process, grown on demand.  The reference builds a fresh libavcodec context per image
// (src/Decoder.cpp:319 constructs an Encoder per call); a CUDA context and its buffers are far too expensive
// for that, so Encoder objects share this one and serialise on it.  Throughput users call the batch API.
struct Shared {
    std::mutex mu;
    h2j_encoder *enc = nullptr;
    int max_w = 0, max_h = 0;
    int device = 0;
    int range_mode = H2J_RANGE_PASSTHROUGH;
    size_t cap = 0;
};
Shared &shared()
{
    static Shared s;
    return s;
}

bool ensure_encoder(Shared &s, int w, int h)
{
    if (s.enc && w <= s.max_w && h <= s.max_h) return true;
    if (s.enc) {
        h2j_destroy(s.enc);
        s.enc = nullptr;
    }
    h2j_settings st;
    h2j_default_settings(&st);
    st.device = s.device;
    st.max_width = w > 1920 ? w : 1920;
    st.max_height = h > 1088 ? h : 1088;
    st.max_batch = 1;
    st.n_slots = 1;
    st.range_mode = s.range_mode;
    // the reference sizes its packet as width*height*3 (src/Encoder.cpp:231) and its copy buffer as HEAP_SIZE
    size_t cap = (size_t)st.max_width * st.max_height * 3;
    if (cap < HEAP_SIZE) cap = HEAP_SIZE;
    st.max_jpeg_bytes = cap;
    const int rc = h2j_create(&st, &s.enc);
    if (rc != H2J_OK) {
        LOG("h2j_create failed, rc=%d (%s), error=%s", rc, h2j_status_string(rc), h2j_last_error(nullptr));
        s.enc = nullptr;
        return false;
    }
    s.max_w = st.max_width;
    s.max_h = st.max_height;
    s.cap = cap;
    return true;
}
}  // namespace

extern "C" void h2j_host_configure(int cuda_device, int range_mode)
{
    Shared &s = shared();
    std::lock_guard<std::mutex> lock(s.mu);
    s.device = cuda_device;
    s.range_mode = range_mode;
    if (s.enc) {
        h2j_destroy(s.enc);
        s.enc = nullptr;
        s.max_w = s.max_h = 0;
    }
}

Encoder::Encoder(const char *const outputFilePath) { this->outputFilePath = outputFilePath; }

Encoder::~Encoder() { release(); }

void Encoder::release()
{
    if (outputFilePath) outputFilePath = nullptr;
}

#ifdef H2J_WITH_LIBAV
bool Encoder::yuv2Jpeg(AVFrame *pFrame)
{
    if (!pFrame) {
        LOG("%s line=%d | null frame", __PRETTY_FUNCTION__, __LINE__);
        release();
        return false;
    }
    H2JFrameView v;
    for (int i = 0; i < 3; i++) {
        v.data[i] = pFrame->data[i];
        v.linesize[i] = pFrame->linesize[i];
    }
    v.width = pFrame->width;
    v.height = pFrame->height;
    v.format = pFrame->format;
    return yuv2Jpeg(v);
}
#else
bool Encoder::yuv2Jpeg(AVFrame *)
{
    LOG("%s | built without libavutil headers (H2J_WITH_LIBAV); use yuv2Jpeg(const H2JFrameView&)", __PRETTY_FUNCTION__);
    release();
    return false;
}
#endif

bool Encoder::yuv2Jpeg(const H2JFrameView &f)
{
    // The reference feeds whatever the decoder produced to an encoder opened as YUVJ420P (src/Encoder.cpp:150);
    // anything but 8-bit 4:2:0 planar would be read as garbage there.  Here it is refused.
    if (f.format != 0 /* AV_PIX_FMT_YUV420P */ && f.format != 12 /* AV_PIX_FMT_YUVJ420P */) {
        LOG("%s line=%d | unsupported pixel format %d (need yuv420p / yuvj420p)", __PRETTY_FUNCTION__, __LINE__, f.format);
        release();
        return false;
    }
    if (!f.data[0] || !f.data[1] || !f.data[2] || f.width < 2 || f.height < 2) {
        LOG("%s line=%d | bad frame: %dx%d", __PRETTY_FUNCTION__, __LINE__, f.width, f.height);
        release();
        return false;
    }
    {
        Shared &s = shared();
        std::lock_guard<std::mutex> lock(s.mu);
        if (!ensure_encoder(s, f.width, f.height)) {
            release();
            return false;
        }
        if (jpegCap_ < s.cap) {
            jpeg_.reset(new (std::nothrow) uint8_t[s.cap]);
            jpegCap_ = jpeg_ ? s.cap : 0;
        }
        if (!jpeg_) {
            LOG("%s line=%d | malloc failed.", __PRETTY_FUNCTION__, __LINE__);
            release();
            return false;
        }
        const uint8_t *planes[3] = {f.data[0], f.data[1], f.data[2]};
        const int strides[3] = {f.linesize[0], f.linesize[1], f.linesize[2]};
        jpegSize_ = 0;
        const int rc = h2j_encode_frame(s.enc, planes, strides, f.width, f.height, jpeg_.get(), jpegCap_, &jpegSize_);
        if (rc != H2J_OK) {
            LOG("h2j_encode_frame failed, rc=%d (%s), error=%s", rc, h2j_status_string(rc), h2j_last_error(s.enc));
            release();
            return false;
        }
    }
    const bool isOk = saveJpegtoFile(this->outputFilePath);
    if (!isOk) {
        LOG("%s line=%d | 保存 Jpeg 文件出错！Jpeg 文件路径：%s", __PRETTY_FUNCTION__, __LINE__, this->outputFilePath);
        return false;
    }
    release();
    return true;
}

bool Encoder::saveJpegtoFile(const char *const filePath)
{
    if (filePath == nullptr || strlen(filePath) == 0) {
        LOG("Jpeg 文件路径为空，请核查！");
        return false;
    }
    FILE *fp_write = fopen(filePath, "wb+");
    if (!fp_write) {
        LOG("%s line=%d | Open file error! filePath=%s, errno=%d", __PRETTY_FUNCTION__, __LINE__, filePath, errno);
        return false;
    }
    const size_t ret = fwrite(jpeg_.get(), 1, jpegSize_, fp_write);
    if (ret == 0) {
        LOG("%s line=%d | fwrite error! Jpeg 文件路径：%s", __PRETTY_FUNCTION__, __LINE__, filePath);
        fclose(fp_write);
        return false;
    }
    LOG("保存 Jpeg 数据到文件: %s", filePath);
    fclose(fp_write);
    return true;
}
