// What a process pays before its first kernel: cuInit (cudaGetDeviceCount), primary context (cudaSetDevice + cudaFree(0)),
// a first kernel launch (module load) and 1 GiB of pinned host memory.  Run several times in a row, alone and while
// another process keeps a context open on the GPU; names the start-up share of the drop-in programs' wall time
// (DESIGN.md "end to end").
//   nvcc -O2 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o scripts/cuda_startup scripts/cuda_startup.cu
//   scripts/cuda_startup [hold_seconds]      (hold_seconds > 0: create a context and sleep -- the "other process")
#include <cuda_runtime.h>
#include <unistd.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>

__global__ void touch(int* p) { *p = 1; }

static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

int main(int argc, char** argv) {
    const double t0 = now();
    int n = 0;
    cudaGetDeviceCount(&n);
    const double t1 = now();
    cudaSetDevice(0);
    cudaFree(0);
    const double t2 = now();
    int* d = nullptr;
    cudaMalloc(&d, 4);
    touch<<<1, 1>>>(d);
    cudaDeviceSynchronize();
    const double t3 = now();
    void* h = nullptr;
    cudaHostAlloc(&h, 1ull << 30, cudaHostAllocDefault);
    const double t4 = now();
    if (argc > 1 && atof(argv[1]) > 0) {
        printf("holding a context on device 0 for %s s\n", argv[1]);
        fflush(stdout);
        usleep((useconds_t)(atof(argv[1]) * 1e6));
        return 0;
    }
    printf("{\"devices\": %d, \"cuInit_s\": %.3f, \"primary_context_s\": %.3f, \"first_kernel_s\": %.3f, \"pinned_1GiB_s\": %.3f, \"total_s\": %.3f}\n",
           n, t1 - t0, t2 - t1, t3 - t2, t4 - t3, t4 - t0);
    fflush(stdout);
    _exit(0);
}
