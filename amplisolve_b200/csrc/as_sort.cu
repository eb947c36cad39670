// Call list housekeeping of the _host caller pipeline, on the device: tile-local slot ids -> panel slot ids, and the
// reference's row order (sample, then slot, then alt: the order in which callVariants walks files, rows and
// substitutions, VC:672, VC:723, VC:869-3288).  A (sample, slot, alt) triple is unique, so the order is total.
// The sort itself is CUB's radix sort of 64-bit keys (library code, off the hot path: ~1e-3 of the records are calls).
#include <cub/device/device_radix_sort.cuh>

#include "as_kernels.h"

__global__ void call_slot_offset_kernel(as_call* __restrict__ calls, const unsigned long long* __restrict__ n_prev,
                                        const unsigned long long* __restrict__ n_now, unsigned long long cap, int32_t off) {
    const unsigned long long lo = min(*n_prev, cap), hi = min(*n_now, cap);
    for (unsigned long long i = lo + (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < hi;
         i += (unsigned long long)gridDim.x * blockDim.x)
        calls[i].slot += off;
}

__global__ void call_key_kernel(const as_call* __restrict__ calls, int64_t n, uint64_t* __restrict__ keys, uint32_t* __restrict__ idx) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const as_call c = calls[i];
    keys[i] = ((uint64_t)(uint32_t)c.sample << 33) | ((uint64_t)(uint32_t)c.slot << 2) | (uint64_t)(c.alt & 3);
    idx[i] = (uint32_t)i;
}

__global__ void call_gather_kernel(const uint4* __restrict__ calls, const uint32_t* __restrict__ idx, int64_t n, uint4* __restrict__ out) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // three 16-byte pieces per 48-byte call
    if (t >= n * 3) return;
    const int64_t i = t / 3, part = t % 3;
    out[t] = calls[(int64_t)idx[i] * 3 + part];
}

cudaError_t as_launch_call_slot_offset(as_call* d_calls, const unsigned long long* d_n_prev, const unsigned long long* d_n_now,
                                       int64_t cap, int32_t off, cudaStream_t st) {
    if (off == 0 || cap <= 0) return cudaSuccess;
    call_slot_offset_kernel<<<148, 256, 0, st>>>(d_calls, d_n_prev, d_n_now, (unsigned long long)cap, off);
    return cudaGetLastError();
}

// scratch layout: keys_in | keys_out | idx_in | idx_out | cub temp
static size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

size_t as_sort_calls_scratch_bytes(int64_t n) {
    if (n <= 0) return 0;
    size_t cub_bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, cub_bytes, (const uint64_t*)nullptr, (uint64_t*)nullptr, (const uint32_t*)nullptr,
                                    (uint32_t*)nullptr, (int)n, 0, 64);
    return 2 * align256((size_t)n * 8) + 2 * align256((size_t)n * 4) + align256(cub_bytes);
}

cudaError_t as_launch_sort_calls(const as_call* d_calls, int64_t n, as_call* d_sorted, void* d_scratch, size_t scratch_bytes,
                                 cudaStream_t st) {
    if (n <= 0) return cudaSuccess;
    if (n > 0x7fffffffll) return cudaErrorInvalidValue;
    char* p = (char*)d_scratch;
    uint64_t* k_in = (uint64_t*)p;  p += align256((size_t)n * 8);
    uint64_t* k_out = (uint64_t*)p; p += align256((size_t)n * 8);
    uint32_t* i_in = (uint32_t*)p;  p += align256((size_t)n * 4);
    uint32_t* i_out = (uint32_t*)p; p += align256((size_t)n * 4);
    size_t cub_bytes = scratch_bytes - (size_t)(p - (char*)d_scratch);
    const unsigned blocks = (unsigned)((n + 255) / 256);
    call_key_kernel<<<blocks, 256, 0, st>>>(d_calls, n, k_in, i_in);
    cudaError_t e = cub::DeviceRadixSort::SortPairs(p, cub_bytes, k_in, k_out, i_in, i_out, (int)n, 0, 64, st);
    if (e != cudaSuccess) return e;
    call_gather_kernel<<<(unsigned)((n * 3 + 255) / 256), 256, 0, st>>>(reinterpret_cast<const uint4*>(d_calls), i_out, n,
                                                                      reinterpret_cast<uint4*>(d_sorted));
    return cudaGetLastError();
}
