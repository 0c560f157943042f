// C entries into the drop-in build (the reference's Decoder.cpp + this repo's Encoder.cpp) for the integration tests and
// bench.py: exactly the calls the reference's README, main.cpp and JNI bridge make.
#include <chrono>
#include <cstdio>
#include <string>
#include <thread>
#include <vector>
#include <unistd.h>

#include "Encoder.h"
#include "IDecoder.h"  // reference export_inc/IDecoder.h

extern "C" int dropin_h265_to_jpeg(const char *in_path, const char *out_path)
{
    auto decoder = IDecoder::getInstance();
    if (!decoder) return 0;
    return decoder->H265ToJpeg(in_path, out_path) ? 1 : 0;
}

// The reference's own harness (main.cpp:37-65) is a loop around IDecoder::getInstance()->H265ToJpeg(in, out).  This is
// that loop, timed: n_calls conversions of in_path, dealt to `threads` caller threads (call i writes
// <out_prefix><i>.jpeg), inside a batch scope of `batch` pictures when batch > 0.  quiet != 0 sends the per-call LOG()
// lines (printf) to /dev/null for the duration.  Returns the number of files the library reports written.
extern "C" int dropin_loop(const char *in_path, const char *out_prefix, int n_calls, int threads, int batch, int quiet, double *seconds)
{
    if (threads < 1) threads = 1;
    int saved = -1;
    if (quiet) {
        fflush(stdout);
        saved = dup(1);
        FILE *n = fopen("/dev/null", "w");
        if (n) { dup2(fileno(n), 1); fclose(n); }
    }
    std::vector<int> ok(threads, 0);
    const auto t0 = std::chrono::steady_clock::now();
    if (batch > 0 && h2j_host_batch_begin(batch) != 0) batch = -1;
    if (batch >= 0) {
        std::vector<std::thread> pool;
        for (int t = 0; t < threads; t++)
            pool.emplace_back([&, t] {
                for (int i = t; i < n_calls; i += threads) {
                    const std::string out = std::string(out_prefix) + std::to_string(i) + ".jpeg";
                    ok[t] += dropin_h265_to_jpeg(in_path, out.c_str());
                }
            });
        for (auto &th : pool) th.join();
    }
    int done = 0;
    for (int t = 0; t < threads; t++) done += ok[t];
    if (batch > 0) {
        int failed = 0;
        done = h2j_host_batch_end(&failed);
    }
    const auto t1 = std::chrono::steady_clock::now();
    if (saved >= 0) {
        fflush(stdout);
        dup2(saved, 1);
        close(saved);
    }
    if (seconds) *seconds = std::chrono::duration<double>(t1 - t0).count();
    return batch < 0 ? -1 : done;
}
