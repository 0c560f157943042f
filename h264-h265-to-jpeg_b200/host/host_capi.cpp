// Plain-C entry into the host mirror so that tests (ctypes) can drive the C++ Encoder class.
#include <cstdint>

#include "Encoder.h"

extern "C" int h2j_host_yuv2jpeg_file(const uint8_t *y, int ys, const uint8_t *u, int us, const uint8_t *v, int vs, int w, int h, int format,
                                      const char *out_path)
{
    H2JFrameView f;
    f.data[0] = y; f.data[1] = u; f.data[2] = v;
    f.linesize[0] = ys; f.linesize[1] = us; f.linesize[2] = vs;
    f.width = w; f.height = h; f.format = format;
    return Encoder(out_path).yuv2Jpeg(f) ? 1 : 0;
}
