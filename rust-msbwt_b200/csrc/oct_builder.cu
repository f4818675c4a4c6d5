// oct_builder.cu -- device-side construction of the OCT image (layout.h) from the quad image and the
// per-position quad codes that are already resident on the device.
//
// The reference has no counterpart (its index is the sampled table of construct_fmindex,
// src/rle_bwt.rs:387-467); what the oct image must reproduce is the composition of eight
// RleBWT::constrain_range calls (src/rle_bwt.rs:202-287); layout.h states the identity.
//
//   1. scatter: one thread per BWT position j with a valid quad code a: LF^4(j) = rank4(a, j) through the
//               quad image (one sector), b = code4(LF^4(j)) (one random read); when b is valid too the
//               position's 20-bit offset is appended to line (a*256+b, j >> 20): the slot comes from an
//               atomicAdd on the line's occurrence counter, which keeps counting past the line's capacity.
//   2. stamp  : one warp per code: exclusive prefix sum of the occurrence counters over the code's
//               buckets, plus C8[code]; lines holding more than kOctCapacity occurrences are counted
//               (the kernel answers those through the quad image).
//   C8[c] = eight constrain_range calls of our own kernel applied to position 0.
#include <algorithm>

#include "../../include/msbwt_gpu.h"
#include "device_rank.cuh"
#include "engine.h"

namespace msbwt {

namespace {

__global__ void __launch_bounds__(256) oct_scatter_kernel(IndexView ix, const uint16_t *__restrict__ codes4,
                                                          uint64_t nbuck8, uint32_t *__restrict__ lines) {
    const C4Base<false> c4{};
    const uint64_t step = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; j < ix.total; j += step) {
        const uint32_t a = codes4[j];
        if (!(a & 0x100u)) continue;
        uint32_t l = (uint32_t)j, h = (uint32_t)j;
        quad_step<false>(ix, c4, a & 255u, l, h);  // l = LF^4(j)
        const uint32_t b = codes4[l];
        if (!(b & 0x100u)) continue;
        const uint32_t code = ((a & 255u) << 8) | (b & 255u);
        uint32_t *line = lines + ((uint64_t)code * nbuck8 + (j >> kOctBucketShift)) * kOctLineWords;
        const uint32_t slot = atomicAdd(line + 1, 1u);
        if (slot < (uint32_t)kOctCapacity) {
            const uint32_t off = (uint32_t)j & ((1u << kOctBucketShift) - 1u);
            uint8_t *e = reinterpret_cast<uint8_t *>(line) + 8 + 3 * slot;
            e[0] = (uint8_t)off;
            e[1] = (uint8_t)(off >> 8);
            e[2] = (uint8_t)(off >> 16);
        }
    }
}

// one warp per code: checkpoints of its buckets
__global__ void __launch_bounds__(256) oct_stamp_kernel(const uint64_t *__restrict__ c8, uint64_t nbuck8,
                                                        uint32_t *__restrict__ lines, unsigned long long *__restrict__ overflow) {
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t code = blockIdx.x * 8u + (threadIdx.x >> 5);
    if (code >= (uint32_t)kOctCodes) return;
    uint32_t *base = lines + (uint64_t)code * nbuck8 * kOctLineWords;
    uint32_t run = (uint32_t)c8[code];
    uint32_t over = 0;
    for (uint64_t b0 = 0; b0 < nbuck8; b0 += 32) {
        const uint64_t b = b0 + lane;
        const uint32_t cnt = b < nbuck8 ? base[b * kOctLineWords + 1] : 0u;
        over += cnt > (uint32_t)kOctCapacity;
        uint32_t incl = cnt;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= (uint32_t)d) incl += t;
        }
        if (b < nbuck8) base[b * kOctLineWords] = run + incl - cnt;
        run += __shfl_sync(0xffffffffu, incl, 31);
    }
#pragma unroll
    for (int d = 16; d; d >>= 1) over += __shfl_xor_sync(0xffffffffu, over, d);
    if (lane == 0 && over) atomicAdd(overflow, (unsigned long long)over);
}

struct Scratch {
    std::vector<void *> ptrs;
    ~Scratch() { for (void *p : ptrs) cudaFree(p); }
    template <class T> cudaError_t alloc(T **p, size_t count) {
        cudaError_t e = cudaMalloc((void **)p, (count ? count : 1) * sizeof(T));
        if (e == cudaSuccess) ptrs.push_back(*p);
        return e;
    }
};

#define O_TRY(expr)                                                                        \
    do {                                                                                   \
        cudaError_t e_ = (expr);                                                           \
        if (e_ != cudaSuccess) {                                                           \
            why = std::string("oct image: ") + #expr + ": " + cudaGetErrorString(e_);      \
            return e_ == cudaErrorMemoryAllocation ? MSBWT_ENOMEM : MSBWT_ECUDA;           \
        }                                                                                  \
    } while (0)

}  // namespace

uint64_t oct_image_bytes(uint64_t total) {
    return (uint64_t)kOctCodes * ((total >> kOctBucketShift) + 1) * kOctLineBytes;
}

int build_oct_image_on_device(int device, const IndexView &ix, const uint16_t *d_codes4, OctImage &img,
                              std::string &why, int *launches) {
    if (!ix.quad || !d_codes4) { why = "oct image: needs the quad image and its codes"; return MSBWT_EINVAL; }
    if (index_is_wide(ix)) { why = "oct image: only for indexes with 32-bit positions (N < 2^32, one superblock)"; return MSBWT_EINVAL; }
    const uint64_t nbuck8 = (ix.total >> kOctBucketShift) + 1;
    const uint64_t nlines = (uint64_t)kOctCodes * nbuck8;
    img.nbuck8 = nbuck8;
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);

    Scratch tmp;
    uint8_t *d_sym = nullptr;
    uint64_t *d_pos = nullptr;
    unsigned long long *d_over = nullptr;
    O_TRY(cudaMalloc((void **)&img.lines, nlines * kOctLineBytes));
    O_TRY(cudaMemsetAsync(img.lines, 0xFF, nlines * kOctLineBytes));           // empty slots: 0xFFFFFF
    O_TRY(cudaMemset2DAsync(img.lines, kOctLineBytes, 0, 8, nlines));          // checkpoint + counter
    O_TRY(tmp.alloc(&d_sym, kOctCodes));
    O_TRY(tmp.alloc(&d_pos, 3 * (size_t)kOctCodes));
    O_TRY(tmp.alloc(&d_over, 1));
    O_TRY(cudaMemsetAsync(d_over, 0, sizeof(unsigned long long)));

    // C8[c]: the eight steps applied to position 0
    static const uint8_t acgt[4] = {1, 2, 3, 5};
    O_TRY(cudaMemsetAsync(d_pos, 0, 3 * (size_t)kOctCodes * sizeof(uint64_t)));
    uint64_t *cur = d_pos, *nxt = d_pos + kOctCodes, *spare = d_pos + 2 * (size_t)kOctCodes;
    std::vector<uint8_t> h_sym(kOctCodes);
    for (int r = 0; r < 8; r++) {
        for (int c = 0; c < kOctCodes; c++) h_sym[(size_t)c] = acgt[(c >> (2 * (7 - r))) & 3];
        O_TRY(cudaMemcpy(d_sym, h_sym.data(), h_sym.size(), cudaMemcpyHostToDevice));
        O_TRY(launch_constrain_ranges(device, ix, d_sym, cur, cur, kOctCodes, nxt, spare, nullptr, launches));
        std::swap(cur, nxt);
    }

    // 1. scatter
    if (ix.total) {
        const unsigned grid = (unsigned)std::min<uint64_t>((ix.total + 255) / 256, (uint64_t)sms * 32);
        oct_scatter_kernel<<<grid, 256>>>(ix, d_codes4, nbuck8, reinterpret_cast<uint32_t *>(img.lines));
        O_TRY(cudaGetLastError());
        if (launches) (*launches)++;
    }
    // 2. stamp
    oct_stamp_kernel<<<kOctCodes / 8, 256>>>(cur, nbuck8, reinterpret_cast<uint32_t *>(img.lines), d_over);
    O_TRY(cudaGetLastError());
    if (launches) (*launches)++;
    unsigned long long over = 0;
    O_TRY(cudaMemcpy(&over, d_over, sizeof(over), cudaMemcpyDeviceToHost));
    img.overflow_lines = over;
    O_TRY(cudaDeviceSynchronize());
    return MSBWT_OK;
}

void free_oct_image(OctImage &img) {
    if (img.lines) cudaFree(img.lines);
    img.lines = nullptr;
}

}  // namespace msbwt
