// See Encoder.h.  Mirrors reference src/Encoder.cpp's control flow and error behaviour; the encode itself is
// one call into the C ABI (include/h2j_b200.h), which runs the CUDA kernels.  There is no CPU encode here:
// if the GPU library cannot be used, yuv2Jpeg logs and returns false.
#include "Encoder.h"

#include <cerrno>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

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
struct Pending {
    int w, h;
    std::string path;
};
struct Shared {
    std::mutex mu;
    h2j_encoder *enc = nullptr;
    int max_w = 0, max_h = 0, max_batch = 1;
    int device = 0;
    int range_mode = H2J_RANGE_PASSTHROUGH;
    size_t cap = 0;
    // batch scope (h2j_host_batch_begin .. h2j_host_batch_end): frames wait here, tightly packed I420 in pinned memory
    bool batching = false;
    int batch_frames = 0;
    uint8_t *pool = nullptr;      // pinned, batch_frames slots of slot_bytes
    size_t slot_bytes = 0;
    std::vector<Pending> pending;
    int written = 0, failed = 0;
};
Shared &shared()
{
    static Shared s;
    return s;
}

bool ensure_encoder(Shared &s, int w, int h, int batch = 1)
{
    if (s.enc && w <= s.max_w && h <= s.max_h && batch <= s.max_batch) return true;
    if (s.enc) {  // grow, never shrink
        if (w < s.max_w) w = s.max_w;
        if (h < s.max_h) h = s.max_h;
        if (batch < s.max_batch) batch = s.max_batch;
    }
    if (s.enc) {
        h2j_destroy(s.enc);
        s.enc = nullptr;
    }
    h2j_settings st;
    h2j_default_settings(&st);
    st.device = s.device;
    st.max_width = w > 1920 ? w : 1920;
    st.max_height = h > 1088 ? h : 1088;
    st.max_batch = batch;
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
    s.max_batch = batch;
    s.cap = cap;
    return true;
}

size_t i420_bytes(int w, int h) { return (size_t)w * h + 2 * (size_t)((w + 1) >> 1) * ((h + 1) >> 1); }

bool write_file(const char *filePath, const uint8_t *data, size_t n)
{
    // reference src/Encoder.cpp:336-361 saveJpegtoFile, same open mode, same log lines
    if (filePath == nullptr || strlen(filePath) == 0) {
        LOG("Jpeg 文件路径为空，请核查！");
        return false;
    }
    FILE *fp_write = fopen(filePath, "wb+");
    if (!fp_write) {
        LOG("%s line=%d | Open file error! filePath=%s, errno=%d", __PRETTY_FUNCTION__, __LINE__, filePath, errno);
        return false;
    }
    const size_t ret = fwrite(data, 1, n, fp_write);
    if (ret == 0) {
        LOG("%s line=%d | fwrite error! Jpeg 文件路径：%s", __PRETTY_FUNCTION__, __LINE__, filePath);
        fclose(fp_write);
        return false;
    }
    LOG("保存 Jpeg 数据到文件: %s", filePath);
    fclose(fp_write);
    return true;
}

// Encode everything that is waiting: runs of consecutive same-sized frames go to the GPU as one batch each.
void flush_pending(Shared &s)
{
    size_t i = 0;
    while (i < s.pending.size()) {
        size_t j = i + 1;
        while (j < s.pending.size() && s.pending[j].w == s.pending[i].w && s.pending[j].h == s.pending[i].h) j++;
        const int n = (int)(j - i), w = s.pending[i].w, h = s.pending[i].h;
        bool ok = ensure_encoder(s, w, h, s.batch_frames);
        std::vector<size_t> offs(n + 1, 0);
        std::vector<int> st(n, 0);
        uint8_t *out = nullptr;
        if (ok) {
            out = static_cast<uint8_t *>(h2j_alloc_pinned(s.cap * n));
            if (!out) {
                LOG("%s line=%d | pinned allocation of %zu bytes failed", __PRETTY_FUNCTION__, __LINE__, s.cap * n);
                ok = false;
            }
        }
        if (ok) {
            int rc = h2j_submit_host(s.enc, 0, s.pool + i * s.slot_bytes, s.slot_bytes, n, w, h);
            if (rc == H2J_OK) rc = h2j_collect(s.enc, 0, out, s.cap * n, offs.data(), st.data());
            if (rc != H2J_OK && rc != H2J_ERR_OUTPUT_TOO_SMALL) {  // (too small: per-frame status says which)
                LOG("h2j batch of %d frames failed, rc=%d (%s), error=%s", n, rc, h2j_status_string(rc), h2j_last_error(s.enc));
                ok = false;
            }
        }
        for (int k = 0; k < n; k++) {
            const Pending &p = s.pending[i + k];
            if (ok && st[k] == H2J_OK && write_file(p.path.c_str(), out + offs[k], offs[k + 1] - offs[k])) s.written++;
            else {
                if (ok && st[k] != H2J_OK) LOG("frame for %s failed: %s", p.path.c_str(), h2j_status_string(st[k]));
                s.failed++;
            }
        }
        if (out) h2j_free_pinned(out);
        i = j;
    }
    s.pending.clear();
}

// A frame arrives inside a batch scope: copy its planes into the next pool slot (the only host-side touch of the pixels).
bool enqueue_frame(Shared &s, const H2JFrameView &f, const char *path)
{
    const size_t need = (i420_bytes(f.width, f.height) + 255) / 256 * 256;
    if (s.pool && need > s.slot_bytes) {  // a bigger picture than the pool was cut for: drain, then re-cut
        flush_pending(s);
        h2j_free_pinned(s.pool);
        s.pool = nullptr;
    }
    if (!s.pool) {
        s.slot_bytes = need;
        s.pool = static_cast<uint8_t *>(h2j_alloc_pinned(s.slot_bytes * s.batch_frames));
        if (!s.pool) {
            LOG("%s line=%d | pinned allocation of %zu bytes failed", __PRETTY_FUNCTION__, __LINE__, s.slot_bytes * s.batch_frames);
            return false;
        }
    }
    uint8_t *p = s.pool + s.pending.size() * s.slot_bytes;
    const int cw = (f.width + 1) >> 1, ch = (f.height + 1) >> 1;
    for (int r = 0; r < f.height; r++) memcpy(p + (size_t)r * f.width, f.data[0] + (size_t)r * f.linesize[0], f.width);
    p += (size_t)f.width * f.height;
    for (int pl = 1; pl <= 2; pl++) {
        for (int r = 0; r < ch; r++) memcpy(p + (size_t)r * cw, f.data[pl] + (size_t)r * f.linesize[pl], cw);
        p += (size_t)cw * ch;
    }
    s.pending.push_back(Pending{f.width, f.height, std::string(path)});
    if ((int)s.pending.size() == s.batch_frames) flush_pending(s);
    return true;
}
}  // namespace

extern "C" int h2j_host_batch_begin(int max_frames)
{
    Shared &s = shared();
    std::lock_guard<std::mutex> lock(s.mu);
    if (s.batching || max_frames < 1) return -1;
    s.batching = true;
    s.batch_frames = max_frames;
    s.written = s.failed = 0;
    s.pending.clear();
    return 0;
}

extern "C" int h2j_host_batch_end(int *failed)
{
    Shared &s = shared();
    std::lock_guard<std::mutex> lock(s.mu);
    if (!s.batching) return -1;
    flush_pending(s);
    s.batching = false;
    if (s.pool) {
        h2j_free_pinned(s.pool);
        s.pool = nullptr;
        s.slot_bytes = 0;
    }
    if (failed) *failed = s.failed;
    return s.written;
}

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
        if (s.batching) {
            // batch scope: the picture is queued; it is encoded and its file written when the batch fills up or ends
            if (this->outputFilePath == nullptr || strlen(this->outputFilePath) == 0) {
                LOG("Jpeg 文件路径为空，请核查！");
                return false;
            }
            const bool queued = enqueue_frame(s, f, this->outputFilePath);
            release();
            return queued;
        }
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
