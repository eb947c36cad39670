// Fisher strand-bias tests of the called variants on the device (VC:3797-3814 evaluated per call at
// VC:902).  One warp per 2x2 table; the lanes split the support [lo, hi] of the hypergeometric distribution into 32
// contiguous runs, each lane adds the pdf(k) <= cutoff of its run in ascending k, and the 32 partial sums are added in
// lane order -- the reference's ascending-k sum up to the association of the additions.
//
// pdf(k) = exp(lc(r,k) + lc(N-r,n-k) - lc(N,n)), lc(n,k) = lg[n] - lg[k] - lg[n-k], lg[i] = lgamma(i + 1) taken from a
// table the HOST filled with its own lgamma: the log-domain values are bit-identical to as_fisher_test's; only exp is
// the device's (<= 1 ulp from glibc's).  Boost.Math, which the reference links, differs from both by ~1e-10 relative
// (tests/golden/fisher_boost.npz), far above that.
#include "as_kernels.h"

namespace {

// canonical operand order (the smaller of k, n - k first), as log_choose of as_host.cpp: tied terms stay tied
__device__ __forceinline__ double lc(const double* __restrict__ lg, unsigned n, unsigned k) {
    const unsigned lo = min(k, n - k), hi = max(k, n - k);
    return __dsub_rn(__dsub_rn(lg[n], lg[lo]), lg[hi]);
}

__global__ void __launch_bounds__(256)
fisher_kernel(const int4* __restrict__ tables, int64_t n_tables, const double* __restrict__ lg, double* __restrict__ p_out) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t t = warp0; t < n_tables; t += n_warps) {
        const int4 T = tables[t];
        const unsigned a = (unsigned)T.x, b = (unsigned)T.y, c = (unsigned)T.z, d = (unsigned)T.w;
        const unsigned N = a + b + c + d, r = a + c, nn = c + d;          // VC:3800-3802
        const unsigned hi = min(r, nn);                                    // VC:3803
        const int lo_i = (int)(r + nn - N);
        const unsigned lo = lo_i > 0 ? (unsigned)lo_i : 0u;                // VC:3804
        const double lcNn = lc(lg, N, nn);
        const double cutoff = exp(__dsub_rn(__dadd_rn(lc(lg, r, c), lc(lg, N - r, nn - c)), lcNn));  // VC:3806
        const unsigned L = hi >= lo ? hi - lo + 1u : 0u;
        const unsigned run = (L + 31u) / 32u;
        const unsigned k0 = lo + (unsigned)lane * run;
        const unsigned k1 = min(k0 + run, hi + 1u);
        double acc = 0.0;
        for (unsigned k = k0; k < k1; ++k) {                               // VC:3808-3812
            const double pk = exp(__dsub_rn(__dadd_rn(lc(lg, r, k), lc(lg, N - r, nn - k)), lcNn));
            if (pk <= cutoff) acc = __dadd_rn(acc, pk);
        }
        double total = 0.0;
#pragma unroll 1
        for (int l = 0; l < 32; ++l) total = __dadd_rn(total, __shfl_sync(0xffffffffu, acc, l));
        if (lane == 0) p_out[t] = total;
    }
}

}  // namespace

cudaError_t as_launch_fisher(const int32_t* d_tables, int64_t n, const double* d_lg, double* d_p, cudaStream_t st) {
    if (n <= 0) return cudaSuccess;
    const int64_t warps = n;
    const unsigned blocks = (unsigned)std::min<int64_t>((warps + 7) / 8, 148 * 8 * 4);  // 8 warps per block, a few waves at most
    fisher_kernel<<<blocks, 256, 0, st>>>(reinterpret_cast<const int4*>(d_tables), n, d_lg, d_p);
    return cudaGetLastError();
}
