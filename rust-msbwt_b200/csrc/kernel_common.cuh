// kernel_common.cuh -- helpers shared by the translation units that hold the search kernels
// (kernels.cu, quad_kernels.cu; split because ptxas 12.9 crashes on the module that holds them all).
#pragma once
#include <cuda_runtime.h>

#include "engine.h"

namespace msbwt {

// Symbols per step of the kernel that walks live list A: 4 with a quad image, 2 with a pair image, else 1.
__host__ __device__ __forceinline__ uint32_t list_a_stride(const IndexView &ix) {
    return ix.quad ? 4u : (ix.pair ? 2u : 1u);
}

// Suffix-table depth for an all-ACGT k-mer.  With a multi-step image (stride 2 or 4) the depth is
// picked from the `stride` deepest levels {ts, ts-1, ..} so that the number of symbols left is a
// multiple of the stride (k below those levels: no table).
__host__ __device__ __forceinline__ uint32_t acgt_table_depth(uint32_t k, uint32_t ts, uint32_t stride) {
    if (!ts) return 0;
    if (k >= ts) {
        const uint32_t back = (stride - (k - ts) % stride) % stride;
        return back < ts ? ts - back : 0;
    }
    return k + stride > ts ? k : 0;  // level k itself is one of the kept ones: the table answers everything
}

inline int sm_count(int device) {
    static int cached[64];
    if (device < 0 || device >= 64) return 148;
    if (!cached[device]) {
        int v = 0;
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, device) != cudaSuccess || v <= 0) v = 148;
        cached[device] = v;
    }
    return cached[device];
}

// one full wave of CTAs (a multiple of the SM count), fewer if there is less work
inline unsigned persistent_grid(int device, const void *kernel, int threads, uint64_t work_groups,
                                int groups_per_cta) {
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, 0) != cudaSuccess || per_sm < 1)
        per_sm = 1;
    uint64_t full = (uint64_t)sm_count(device) * (uint64_t)per_sm;
    uint64_t need = (work_groups + groups_per_cta - 1) / groups_per_cta;
    if (need < 1) need = 1;
    return (unsigned)(need < full ? need : full);
}

}  // namespace msbwt
