// Pinned host -> device bandwidth of one box, alone and under contention: every single GPU, every pair, the two halves,
// an interleaved half and all GPUs at once, for plain pinned, write-combined and per-thread-first-touched buffers.
// Answers VERDICT r01 item 5 ("find and name the N >= 4 e2e limiter": switch uplink vs host DRAM vs the links themselves).
//   nvcc -O2 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o scripts/h2d_matrix scripts/h2d_matrix.cu
//   scripts/h2d_matrix > profiles/r02_h2d_matrix.json          (on the multi-GPU box)
#include <cuda_runtime.h>

#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

static const size_t BYTES = 1ull << 30;
static const int REPS = 4;

struct Barrier {
    std::atomic<int> count{0}, gen{0};
    int n;
    explicit Barrier(int n_) : n(n_) {}
    void wait() {
        const int g = gen.load();
        if (count.fetch_add(1) + 1 == n) { count.store(0); gen.fetch_add(1); }
        else while (gen.load() == g) std::this_thread::yield();
    }
};

// GB/s per device of `devs` copying concurrently; mode 0 plain pinned, 1 write-combined, 2 plain pinned allocated and first
// touched by the copying thread itself (the others are allocated by the main thread)
static std::vector<double> run(const std::vector<int>& devs, int mode, std::vector<void*>& host_main, std::vector<void*>& host_wc,
                               std::vector<void*>& dbuf) {
    const int n = (int)devs.size();
    std::vector<double> gbs((size_t)n, 0.0);
    Barrier bar(n);
    std::vector<std::thread> th;
    for (int k = 0; k < n; ++k) {
        th.emplace_back([&, k]() {
            const int d = devs[(size_t)k];
            cudaSetDevice(d);
            void* h = mode == 1 ? host_wc[(size_t)d] : host_main[(size_t)d];
            void* own = nullptr;
            if (mode == 2) {
                cudaHostAlloc(&own, BYTES, cudaHostAllocDefault);
                memset(own, 1, BYTES);
                h = own;
            }
            cudaStream_t st;
            cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
            cudaMemcpyAsync(dbuf[(size_t)d], h, BYTES, cudaMemcpyHostToDevice, st);  // warm-up
            cudaStreamSynchronize(st);
            bar.wait();
            const auto t0 = std::chrono::steady_clock::now();
            for (int r = 0; r < REPS; ++r) cudaMemcpyAsync(dbuf[(size_t)d], h, BYTES, cudaMemcpyHostToDevice, st);
            cudaStreamSynchronize(st);
            const double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            gbs[(size_t)k] = REPS * (double)BYTES / s / 1e9;
            cudaStreamDestroy(st);
            if (own) cudaFreeHost(own);
        });
    }
    for (auto& t : th) t.join();
    return gbs;
}

static void print_case(const char* name, const std::vector<int>& devs, const std::vector<double>& g, bool last = false) {
    double sum = 0;
    printf("  {\"case\": \"%s\", \"gpus\": [", name);
    for (size_t i = 0; i < devs.size(); ++i) printf("%s%d", i ? ", " : "", devs[i]);
    printf("], \"gbs_per_gpu\": [");
    for (size_t i = 0; i < g.size(); ++i) { printf("%s%.1f", i ? ", " : "", g[i]); sum += g[i]; }
    printf("], \"aggregate_gbs\": %.1f}%s\n", sum, last ? "" : ",");
}

int main() {
    int n = 0;
    cudaGetDeviceCount(&n);
    std::vector<void*> host_main((size_t)n), host_wc((size_t)n), dbuf((size_t)n);
    for (int d = 0; d < n; ++d) {
        cudaSetDevice(d);
        cudaHostAlloc(&host_main[(size_t)d], BYTES, cudaHostAllocDefault);
        cudaHostAlloc(&host_wc[(size_t)d], BYTES, cudaHostAllocWriteCombined);
        memset(host_main[(size_t)d], 1, BYTES);
        memset(host_wc[(size_t)d], 1, BYTES);
        cudaMalloc(&dbuf[(size_t)d], BYTES);
    }
    printf("{\"what\": \"pinned host -> device bandwidth, %d x 1 GiB copies per GPU, GPUs of a case copy concurrently\", \"n_gpus\": %d,\n"
           " \"cases\": [\n", REPS, n);
    for (int d = 0; d < n; ++d) print_case("single", {d}, run({d}, 0, host_main, host_wc, dbuf));
    for (int a = 0; a < n; ++a)
        for (int b = a + 1; b < n; ++b) print_case("pair", {a, b}, run({a, b}, 0, host_main, host_wc, dbuf));
    std::vector<int> all, lo, hi, mix;
    for (int d = 0; d < n; ++d) { all.push_back(d); (d < n / 2 ? lo : hi).push_back(d); if ((d % 4) < 2) mix.push_back(d); }
    if (n >= 4) {
        print_case("lower half", lo, run(lo, 0, host_main, host_wc, dbuf));
        print_case("upper half", hi, run(hi, 0, host_main, host_wc, dbuf));
        print_case("interleaved half", mix, run(mix, 0, host_main, host_wc, dbuf));
        print_case("lower half, write-combined", lo, run(lo, 1, host_main, host_wc, dbuf));
        print_case("lower half, buffers allocated and first touched by the copying threads", lo, run(lo, 2, host_main, host_wc, dbuf));
    }
    print_case("all, write-combined", all, run(all, 1, host_main, host_wc, dbuf));
    print_case("all, buffers allocated and first touched by the copying threads", all, run(all, 2, host_main, host_wc, dbuf));
    print_case("all", all, run(all, 0, host_main, host_wc, dbuf), true);
    printf(" ]}\n");
    return 0;
}
