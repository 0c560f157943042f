// Host-side copy into pinned staging with non-temporal stores.
//
// The destination of h2j_encode_frame's staging copy is read next by the GPU's DMA engine, never by this core: ordinary
// stores would first fetch every destination line into the cache (read for ownership) and evict the decoder's working
// set with pixels nobody reads again.  Streaming stores write the lines straight out: 14.3 instead of 11.7 GB/s for a
// 1080p frame on the hosts of this pool (one thread; the copy is memory-bound, more threads did not help); the source is
// prefetched a block ahead (+9 %).
// Compiled by the host compiler alone (nvcc hands .cpp files through), AVX2 only inside the one function that is
// entered after a CPU check.
#include <cstddef>
#include <cstdint>
#include <cstring>

#if defined(__x86_64__) && defined(__GNUC__)
#include <immintrin.h>

__attribute__((target("avx2"))) static void stream_copy_avx2(uint8_t *d, const uint8_t *s, size_t n)
{
    size_t head = (32 - ((uintptr_t)d & 31)) & 31;  // streaming stores need a 32-byte aligned destination
    if (head > n) head = n;
    memcpy(d, s, head);
    d += head;
    s += head;
    n -= head;
    size_t i = 0;
    // 2 KiB at a time with the NEXT 2 KiB requested into the cache first: the loads of a block then find their lines on the way
    // instead of each waiting for its own miss (a source that is not in the cache: 12.6 -> 13.7 GB/s on the pool's hosts,
    // tools/microbench/host_copy_variants.cpp; a frame the decoder has just written is, and the call measures the same;
    // prefetches do not fault, so running past the end of the source is harmless)
    constexpr size_t kBlock = 2048;
    for (; i + kBlock <= n; i += kBlock) {
        for (size_t k = 0; k < kBlock; k += 64) _mm_prefetch(reinterpret_cast<const char *>(s + i + kBlock + k), _MM_HINT_T0);
        for (size_t k = i; k < i + kBlock; k += 128) {
            const __m256i a = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(s + k));
            const __m256i b = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(s + k + 32));
            const __m256i c = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(s + k + 64));
            const __m256i e = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(s + k + 96));
            _mm256_stream_si256(reinterpret_cast<__m256i *>(d + k), a);
            _mm256_stream_si256(reinterpret_cast<__m256i *>(d + k + 32), b);
            _mm256_stream_si256(reinterpret_cast<__m256i *>(d + k + 64), c);
            _mm256_stream_si256(reinterpret_cast<__m256i *>(d + k + 96), e);
        }
    }
    for (; i + 128 <= n; i += 128) {
        const __m256i a = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(s + i));
        const __m256i b = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(s + i + 32));
        const __m256i c = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(s + i + 64));
        const __m256i e = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(s + i + 96));
        _mm256_stream_si256(reinterpret_cast<__m256i *>(d + i), a);
        _mm256_stream_si256(reinterpret_cast<__m256i *>(d + i + 32), b);
        _mm256_stream_si256(reinterpret_cast<__m256i *>(d + i + 64), c);
        _mm256_stream_si256(reinterpret_cast<__m256i *>(d + i + 96), e);
    }
    _mm_sfence();  // the stores must be globally visible before the cudaMemcpyAsync that follows is submitted
    memcpy(d + i, s + i, n - i);
}

extern "C" void h2j_stream_copy(void *dst, const void *src, size_t n)
{
    static const bool avx2 = __builtin_cpu_supports("avx2");
    if (avx2 && n >= 4096) stream_copy_avx2(static_cast<uint8_t *>(dst), static_cast<const uint8_t *>(src), n);
    else memcpy(dst, src, n);
}
#else
extern "C" void h2j_stream_copy(void *dst, const void *src, size_t n) { memcpy(dst, src, n); }
#endif
