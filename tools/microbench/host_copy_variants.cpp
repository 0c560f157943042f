#include <immintrin.h>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cstdint>
typedef void (*fn)(uint8_t*, const uint8_t*, size_t);
__attribute__((target("avx2"))) static void base(uint8_t *d, const uint8_t *s, size_t n){
    size_t i=0; for(;i+128<=n;i+=128){ __m256i a=_mm256_loadu_si256((const __m256i*)(s+i)),b=_mm256_loadu_si256((const __m256i*)(s+i+32)),c=_mm256_loadu_si256((const __m256i*)(s+i+64)),e=_mm256_loadu_si256((const __m256i*)(s+i+96));
    _mm256_stream_si256((__m256i*)(d+i),a);_mm256_stream_si256((__m256i*)(d+i+32),b);_mm256_stream_si256((__m256i*)(d+i+64),c);_mm256_stream_si256((__m256i*)(d+i+96),e);} _mm_sfence(); }
template<int DIST, int HINT> __attribute__((target("avx2"))) static void pf(uint8_t *d, const uint8_t *s, size_t n){
    size_t i=0; for(;i+128<=n;i+=128){ _mm_prefetch((const char*)(s+i+DIST), (_mm_hint)HINT); _mm_prefetch((const char*)(s+i+DIST+64), (_mm_hint)HINT);
    __m256i a=_mm256_loadu_si256((const __m256i*)(s+i)),b=_mm256_loadu_si256((const __m256i*)(s+i+32)),c=_mm256_loadu_si256((const __m256i*)(s+i+64)),e=_mm256_loadu_si256((const __m256i*)(s+i+96));
    _mm256_stream_si256((__m256i*)(d+i),a);_mm256_stream_si256((__m256i*)(d+i+32),b);_mm256_stream_si256((__m256i*)(d+i+64),c);_mm256_stream_si256((__m256i*)(d+i+96),e);} _mm_sfence(); }
// block-wise: read a 4 KiB block into cache with plain loads first (all loads outstanding), then stream it out
__attribute__((target("avx2"))) static void blocked(uint8_t *d, const uint8_t *s, size_t n){
    size_t i=0; for(;i+4096<=n;i+=4096){ for(int k=0;k<4096;k+=64) _mm_prefetch((const char*)(s+i+4096+k), _MM_HINT_T0);
      for(int k=0;k<4096;k+=128){ __m256i a=_mm256_loadu_si256((const __m256i*)(s+i+k)),b=_mm256_loadu_si256((const __m256i*)(s+i+k+32)),c=_mm256_loadu_si256((const __m256i*)(s+i+k+64)),e=_mm256_loadu_si256((const __m256i*)(s+i+k+96));
      _mm256_stream_si256((__m256i*)(d+i+k),a);_mm256_stream_si256((__m256i*)(d+i+k+32),b);_mm256_stream_si256((__m256i*)(d+i+k+64),c);_mm256_stream_si256((__m256i*)(d+i+k+96),e);} } _mm_sfence(); }
template<int BLK,int AHEAD> __attribute__((target("avx2"))) static void blk(uint8_t *d, const uint8_t *s, size_t n){
    size_t i=0; for(;i+BLK<=n;i+=BLK){ for(int k=0;k<BLK;k+=64) _mm_prefetch((const char*)(s+i+AHEAD*BLK+k), _MM_HINT_T0);
      for(int k=0;k<BLK;k+=128){ __m256i a=_mm256_loadu_si256((const __m256i*)(s+i+k)),b=_mm256_loadu_si256((const __m256i*)(s+i+k+32)),c=_mm256_loadu_si256((const __m256i*)(s+i+k+64)),e=_mm256_loadu_si256((const __m256i*)(s+i+k+96));
      _mm256_stream_si256((__m256i*)(d+i+k),a);_mm256_stream_si256((__m256i*)(d+i+k+32),b);_mm256_stream_si256((__m256i*)(d+i+k+64),c);_mm256_stream_si256((__m256i*)(d+i+k+96),e);} } _mm_sfence(); }
static void mc(uint8_t *d, const uint8_t *s, size_t n){ memcpy(d,s,n); }
int main(){ const size_t N=3110400, NB=64; uint8_t *src=(uint8_t*)aligned_alloc(4096,N*NB+8192), *dst=(uint8_t*)aligned_alloc(4096,N*NB+8192); memset(src,1,N*NB); memset(dst,2,N*NB);
 struct {const char*n; fn f;} v[]={{"memcpy",mc},{"base",base},{"pf512_t0",pf<512,_MM_HINT_T0>},{"pf1024_t0",pf<1024,_MM_HINT_T0>},{"pf2048_t0",pf<2048,_MM_HINT_T0>},{"pf1024_nta",pf<1024,_MM_HINT_NTA>},{"pf4096_nta",pf<4096,_MM_HINT_NTA>},{"blocked4k",blocked},{"blk2k_1",blk<2048,1>},{"blk4k_2",blk<4096,2>},{"blk8k_1",blk<8192,1>},{"blk16k_1",blk<16384,1>},{"blk1k_2",blk<1024,2>},{"blk1k_4",blk<1024,4>}};
 for(int rep=0;rep<2;rep++) for(auto &x: v){ auto t0=std::chrono::steady_clock::now(); for(size_t b=0;b<NB;b++) x.f(dst+b*N, src+((b*7)%NB)*N, N); double dt=std::chrono::duration<double>(std::chrono::steady_clock::now()-t0).count(); printf("%-12s %.2f GB/s  %.3f ms per frame\n", x.n, N*NB/dt/1e9, dt/NB*1e3);} }
