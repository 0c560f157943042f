// Raw host->device link probe: plain pinned cudaMemcpyAsync on several GPUs AT ONCE, one host thread per GPU, nothing
// of the encoder involved.  Answers what the box gives when 1 / 2 / 4 / 8 GPUs pull frames out of host memory together,
// for several GPU subsets (neighbours vs spread over the PCIe switches) and several kinds of host memory:
//   pinned   cudaHostAlloc(default)              what bench.py's e2e pass uses
//   wc       cudaHostAlloc(write-combined)       no CPU cache snooping on the read
//   huge     mmap + MADV_HUGEPAGE + cudaHostRegister   2 MiB pages behind the IOMMU
// Usage: pcie_probe [mb_per_copy=199] [copies=40] ; prints one JSON line per (subset, kind).
//        pcie_probe all [mb_per_copy=64] [copies=12] ; ONE line: every visible GPU copying at once (what bench.py reads to
//        decide which GPUs a job with fewer ranks than GPUs should use).
#include <cuda_runtime.h>
#include <sys/mman.h>

#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#define CK(x)                                                                                  \
    do {                                                                                       \
        cudaError_t e_ = (x);                                                                  \
        if (e_ != cudaSuccess) {                                                               \
            fprintf(stderr, "%s: %s (%s:%d)\n", #x, cudaGetErrorString(e_), __FILE__, __LINE__); \
            exit(1);                                                                           \
        }                                                                                      \
    } while (0)

enum Kind { PINNED, WC, HUGE };
static const char *kind_name[] = {"pinned", "wc", "huge"};

struct HostBuf {
    void *p = nullptr;
    size_t bytes = 0;
    Kind kind = PINNED;
    bool ok = false;
};

static HostBuf host_alloc(size_t bytes, Kind k)
{
    HostBuf b;
    b.bytes = bytes;
    b.kind = k;
    if (k == PINNED) b.ok = cudaHostAlloc(&b.p, bytes, cudaHostAllocDefault) == cudaSuccess;
    else if (k == WC) b.ok = cudaHostAlloc(&b.p, bytes, cudaHostAllocWriteCombined) == cudaSuccess;
    else {
        const size_t al = (bytes + (2u << 20) - 1) & ~((size_t)(2u << 20) - 1);
        void *m = mmap(nullptr, al + (2u << 20), PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
        if (m == MAP_FAILED) return b;
        void *a = (void *)(((uintptr_t)m + (2u << 20) - 1) & ~((uintptr_t)(2u << 20) - 1));
        madvise(a, al, MADV_HUGEPAGE);
        memset(a, 1, al);
        b.ok = cudaHostRegister(a, al, cudaHostRegisterDefault) == cudaSuccess;
        b.p = a;
    }
    if (!b.ok) cudaGetLastError();
    else if (k != HUGE) memset(b.p, 1, bytes);
    return b;
}

struct Result { double gbs; };

static void run_set(const std::vector<int> &devs, Kind kind, size_t bytes, int copies, bool with_d2h)
{
    const int n = (int)devs.size();
    std::vector<double> secs(n, 0.0);
    std::vector<int> fail(n, 0);
    std::atomic<int> ready{0}, go{0};
    std::vector<std::thread> th;
    for (int i = 0; i < n; i++)
        th.emplace_back([&, i] {
            CK(cudaSetDevice(devs[i]));
            void *d = nullptr, *d2 = nullptr;
            CK(cudaMalloc(&d, bytes));
            HostBuf h = host_alloc(bytes, kind);
            HostBuf h2;
            cudaStream_t s, s2;
            CK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
            CK(cudaStreamCreateWithFlags(&s2, cudaStreamNonBlocking));
            if (with_d2h) {
                CK(cudaMalloc(&d2, bytes / 12));
                h2 = host_alloc(bytes / 12, PINNED);
            }
            if (!h.ok) fail[i] = 1;
            else {
                CK(cudaMemcpyAsync(d, h.p, bytes, cudaMemcpyHostToDevice, s));  // warm
                CK(cudaStreamSynchronize(s));
            }
            ready++;
            while (!go.load()) std::this_thread::yield();
            const auto t0 = std::chrono::steady_clock::now();
            if (h.ok) {
                for (int c = 0; c < copies; c++) {
                    CK(cudaMemcpyAsync(d, h.p, bytes, cudaMemcpyHostToDevice, s));
                    if (with_d2h && h2.ok) CK(cudaMemcpyAsync(h2.p, d2, bytes / 12, cudaMemcpyDeviceToHost, s2));
                }
                CK(cudaStreamSynchronize(s));
                CK(cudaStreamSynchronize(s2));
            }
            secs[i] = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            cudaFree(d);
            if (d2) cudaFree(d2);
            if (h.ok) {
                if (kind == HUGE) cudaHostUnregister(h.p);
                else cudaFreeHost(h.p);
            }
            if (h2.ok) cudaFreeHost(h2.p);
            cudaStreamDestroy(s);
            cudaStreamDestroy(s2);
        });
    while (ready.load() < n) std::this_thread::yield();
    go = 1;
    for (auto &t : th) t.join();
    double worst = 0, sum = 0;
    std::string per = "[", set = "[";
    bool failed = false;
    for (int i = 0; i < n; i++) {
        failed |= fail[i] != 0;
        const double g = secs[i] > 0 ? (double)bytes * copies / secs[i] / 1e9 : 0;
        worst = secs[i] > worst ? secs[i] : worst;
        sum += g;
        char b[64];
        snprintf(b, sizeof b, "%s%.1f", i ? ", " : "", g);
        per += b;
        snprintf(b, sizeof b, "%s%d", i ? ", " : "", devs[i]);
        set += b;
    }
    per += "]";
    set += "]";
    printf("{\"gpus\": %s, \"n\": %d, \"host_memory\": \"%s\", \"d2h_alongside\": %s, \"mb_per_copy\": %.0f, \"copies\": %d, \"failed\": %s, "
           "\"h2d_gbs_per_gpu\": %s, \"h2d_gbs_total\": %.1f, \"frames_1080p_per_s\": %.0f}\n",
           set.c_str(), n, kind_name[kind], with_d2h ? "true" : "false", bytes / 1e6, copies, failed ? "true" : "false", per.c_str(),
           worst > 0 ? (double)bytes * copies * n / worst / 1e9 : 0.0, worst > 0 ? (double)bytes * copies * n / worst / 3110400.0 : 0.0);
    fflush(stdout);
}

int main(int argc, char **argv)
{
    if (argc > 1 && !strcmp(argv[1], "all")) {
        int nd = 0;
        CK(cudaGetDeviceCount(&nd));
        std::vector<int> all;
        for (int i = 0; i < nd; i++) all.push_back(i);
        run_set(all, PINNED, (size_t)(argc > 2 ? atoi(argv[2]) : 64) * 1000 * 1000, argc > 3 ? atoi(argv[3]) : 12, false);
        return 0;
    }
    const size_t mb = argc > 1 ? (size_t)atoi(argv[1]) : 199;
    const int copies = argc > 2 ? atoi(argv[2]) : 40;
    int ndev = 0;
    CK(cudaGetDeviceCount(&ndev));
    const size_t bytes = mb * 1000 * 1000;
    std::vector<std::vector<int>> sets;
    sets.push_back({0});
    if (ndev >= 2) sets.push_back({0, 1});
    if (ndev >= 8) { sets.push_back({0, 4}); sets.push_back({0, 2}); }
    if (ndev >= 4) sets.push_back({0, 1, 2, 3});
    if (ndev >= 8) {
        sets.push_back({0, 2, 4, 6});
        sets.push_back({4, 5, 6, 7});
        sets.push_back({0, 1, 2, 3, 4, 5, 6, 7});
    }
    for (auto &s : sets) run_set(s, PINNED, bytes, copies, false);
    for (auto &s : sets)
        if (s.size() == 1 || (int)s.size() == ndev) {
            run_set(s, WC, bytes, copies, false);
            run_set(s, HUGE, bytes, copies, false);
            run_set(s, PINNED, bytes, copies, true);
        }
    return 0;
}
