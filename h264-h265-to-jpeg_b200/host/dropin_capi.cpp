// C entry into the drop-in build (reference Decoder.cpp + this repo's Encoder.cpp) for the integration test:
// exactly the call the reference's README and JNI bridge make.
#include "IDecoder.h"  // reference export_inc/IDecoder.h

extern "C" int dropin_h265_to_jpeg(const char *in_path, const char *out_path)
{
    auto decoder = IDecoder::getInstance();
    if (!decoder) return 0;
    return decoder->H265ToJpeg(in_path, out_path) ? 1 : 0;
}
