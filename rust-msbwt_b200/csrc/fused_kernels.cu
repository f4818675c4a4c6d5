// fused_kernels.cu -- the fused path of the fixed-k device entry point: symbol bytes in, counts out, one kernel
// (count_kmers_oct_kernel<true>, oct_kernel.cuh) + the few k-mers it sets aside.
//
// Replaces, for k <= 32 on an index with an oct image, the pack / seed kernel followed by the search kernel:
// BWT::count_kmer (src/msbwt_core.rs:125-161) on a batch of n * k symbol bytes.
#include "oct_kernel.cuh"

namespace msbwt {

// The k-mers the fused kernel set aside (a symbol outside ACGT): BWT::count_kmer step by step over the one-step
// blocks, straight from the caller's symbol bytes (src/msbwt_core.rs:125-161: symbols >= 6 are refused --
// here: flagged in `status` and counted as if they were '$').
__global__ void __launch_bounds__(kCountThreads, 4)
count_exceptions_kernel(IndexView ix, const uint8_t *__restrict__ syms, uint32_t k, const uint32_t *__restrict__ exc,
                        uint64_t *__restrict__ out, uint32_t *__restrict__ status) {
    __shared__ uint64_t cb_smem[4];
    const CBase<false> cb = stage_cbase<false>(ix, cb_smem);
    const uint32_t n = exc[0];
    for (uint32_t i = blockIdx.x * kCountThreads + threadIdx.x; i < n; i += gridDim.x * kCountThreads) {
        const uint32_t q = exc[1u + i];
        const uint8_t *src = syms + (uint64_t)q * k;
        uint32_t l = 0, h = (uint32_t)ix.total;
        for (uint32_t t = k; t > 0 && l != h; t--) {
            uint32_t sym = src[t - 1];
            if (sym >= (uint32_t)kAlphabet) { atomicOr(status, 1u); sym = 0; }
            rank_step<false, 1>(ix, cb, sym, l, h);
        }
        out[q] = (uint64_t)(h - l);
    }
}

bool fused_path_applies(const IndexView &ix, const uint8_t *d_syms, uint32_t k) {
    return ix.oct && !index_is_wide(ix) && k >= 1 && k <= 32 && (reinterpret_cast<uintptr_t>(d_syms) & 15u) == 0;
}
uint64_t fused_scratch_bytes(uint64_t n) { return (n + 4) * sizeof(uint32_t); }

// n <= 2^30 k-mers of k <= 32 symbol bytes (16-byte aligned buffer) straight to their counts.  `d_scratch`:
// fused_scratch_bytes(n) bytes of device memory -- [0] chunk dispenser, [1] number of exceptions, [2..] their
// query indices.
cudaError_t launch_count_fused(int device, const IndexView &ix, const uint8_t *d_syms, uint32_t k, uint64_t n,
                               uint64_t *d_out, uint32_t *d_status, uint32_t *d_scratch, cudaStream_t st, int *launches) {
    if (!n) return cudaSuccess;
    if (n > kMaxPerLaunch) return cudaErrorInvalidValue;
    static bool prepared[64] = {};  // per device: 4 CTAs x 53 KB of staging per SM need the large shared-memory configuration
    if (device < 0 || device >= 64 || !prepared[device]) {
        if (cudaError_t e = cudaFuncSetAttribute((const void *)count_kmers_oct_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kOctSmemRaw); e != cudaSuccess) return e;
        cudaFuncSetAttribute((const void *)count_kmers_oct_kernel<true>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        if (device >= 0 && device < 64) prepared[device] = true;
    }
    if (cudaError_t e = cudaMemsetAsync(d_scratch, 0, 2 * sizeof(uint32_t), st); e != cudaSuccess) return e;
    const PackedLayout lay{n, 1u, 1u};
    const unsigned grid = oct_grid(device, (const void *)count_kmers_oct_kernel<true>, kOctSmemRaw, n);
    count_kmers_oct_kernel<true><<<grid, kCountThreads, kOctSmemRaw, st>>>(ix, nullptr, lay, k, d_out, d_scratch, d_syms, (uint32_t)n, d_scratch + 1);
    if (cudaError_t e = cudaGetLastError(); e != cudaSuccess) return e;
    count_exceptions_kernel<<<(unsigned)sm_count(device), kCountThreads, 0, st>>>(ix, d_syms, k, d_scratch + 1, d_out, d_status);
    if (launches) (*launches) += 2;
    return cudaGetLastError();
}

}  // namespace msbwt
