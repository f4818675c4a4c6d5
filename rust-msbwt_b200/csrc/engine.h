// engine.h -- internal interface between the C ABI (capi.cu), the host-side layout
// builder (loader.cu) and the kernels (kernels.cu).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <string>
#include <vector>

#include "layout.h"

namespace msbwt {

constexpr int kCountThreads = 256;
constexpr uint64_t kMaxPerLaunch = 1ull << 30;  // queries per pack/count launch (u32 indices inside the kernels)
// resident CTAs per SM the kernels are compiled for (register budget = 65536 / (256 * min_ctas)):
// one thread per query keeps two 64-byte blocks (32 registers) in flight, a lane pair half of that
constexpr int min_ctas(bool wide, int lanes) { return lanes == 2 ? (wide ? 4 : 5) : (wide ? 3 : 4); }

inline uint32_t words_for_k(uint32_t k) { return k ? (k + kSymsPerWord - 1) / kSymsPerWord : 1; }

// ---- loader.cu: RLE byte stream / .npy -> host block image ----
struct HostImage {
    std::vector<uint32_t> blocks;  // nblocks * 16 words
    std::vector<uint32_t> aux;     // nblocks * 2 ($, N checkpoints)
    std::vector<uint64_t> cbase;   // n_super * 8
    uint64_t counts[kAlphabet] = {0, 0, 0, 0, 0, 0};
    uint64_t start[kAlphabet] = {0, 0, 0, 0, 0, 0};
    uint64_t total = 0;
    uint64_t nblocks = 0;
    uint32_t n_super = 0;
    uint32_t sb_shift = kDefaultSuperShift;
};

// ---- builder.cu: RLE byte stream -> block image built ON the current device ----
struct DeviceImage {
    uint4 *blocks = nullptr;
    uint32_t *aux = nullptr;
    uint64_t *cbase = nullptr;
    uint64_t counts[kAlphabet] = {0, 0, 0, 0, 0, 0};
    uint64_t start[kAlphabet] = {0, 0, 0, 0, 0, 0};
    uint64_t total = 0;
    uint64_t nblocks = 0;
    uint32_t n_super = 0;
    uint32_t sb_shift = kDefaultSuperShift;
};
int build_image_on_device(const uint8_t *h_rle, uint64_t len, uint32_t sb_shift, DeviceImage &img, std::string &why);
void free_device_image(DeviceImage &img);

// return an msbwt_status; on failure `why` explains
int validate_rle(const uint8_t *rle, uint64_t len, std::string &why);
int build_image_from_rle(const uint8_t *rle, uint64_t len, uint32_t sb_shift, HostImage &img, std::string &why);
int read_npy_payload(const char *path, std::vector<uint8_t> &payload, std::string &why);

// ---- kernels.cu ----
cudaError_t launch_pack_seed(const IndexView &ix, const uint8_t *d_syms, uint32_t k, uint64_t n, uint64_t *d_packed,
                             uint64_t *d_out, uint32_t *d_status, cudaStream_t st);
cudaError_t launch_table_extend(int device, const IndexView &ix, const void *d_parent, void *d_child,
                                uint32_t n_child, cudaStream_t st);
bool index_is_wide(const IndexView &ix);
// Device scratch produced by pack_seed_kernel and consumed by count_kmers_packed_kernel
// (all u64 units, n = queries in the batch, `words` = words_for_k(k)):
//   [0, n)                      W0    first symbol word of the i-th LIVE query      (compacted)
//   [n, n + seedw*n)            SEED  starting range: l | h<<32, or l then h (WIDE)   (compacted)
//   [.., + (n+1)/2)             QIDX  u32 original query index of the i-th live query (compacted)
//   [.., + (words-1)*n)         WX    symbol words 1.. of query q, word-major, by ORIGINAL index
//   [.., + 1)                   LIVE  number of live queries (u64 counter)
// Queries that need no search step (empty seed range, or the suffix table answered every symbol)
// are finished by the pack kernel itself and never reach the search kernel.
struct PackedLayout {
    uint64_t n;
    uint32_t words, seedw;
    __host__ __device__ uint64_t w0() const { return 0; }
    __host__ __device__ uint64_t seed() const { return n; }
    __host__ __device__ uint64_t qidx() const { return n + (uint64_t)seedw * n; }
    __host__ __device__ uint64_t wx() const { return qidx() + (n + 1) / 2; }
    __host__ __device__ uint64_t live() const { return wx() + (uint64_t)(words - 1) * n; }
    __host__ __device__ uint64_t total() const { return live() + 1; }
};
inline PackedLayout packed_layout(const IndexView &ix, uint32_t k, uint64_t n) {
    return PackedLayout{n, words_for_k(k), index_is_wide(ix) ? 2u : 1u};
}
// `launches` (optional) is incremented once per kernel launch issued
cudaError_t launch_count_packed(int device, const IndexView &ix, int lanes, const uint64_t *d_packed, uint32_t k,
                                uint64_t n, uint64_t *d_out, cudaStream_t st, int *launches);
cudaError_t launch_count_bytes(int device, const IndexView &ix, const uint8_t *d_syms,
                               const uint64_t *d_offsets, uint64_t n, uint64_t *d_out, uint32_t *d_status,
                               cudaStream_t st, int *launches);
cudaError_t launch_constrain_ranges(int device, const IndexView &ix, const uint8_t *d_sym, const uint64_t *d_l,
                                    const uint64_t *d_h, uint64_t n, uint64_t *d_out_l, uint64_t *d_out_h,
                                    cudaStream_t st, int *launches);
cudaError_t launch_gather(int device, const void *d_buf, uint64_t buf_bytes, uint32_t granule,
                          uint64_t n_gathers, uint64_t seed, uint64_t *d_sink, cudaStream_t st);

}  // namespace msbwt
