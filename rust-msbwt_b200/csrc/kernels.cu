// kernels.cu -- sm_100a kernels for the batched backward search.
//
// Replaces the reference's hot loop: BWT::count_kmer (src/msbwt_core.rs:125-161)
// calling RleBWT::constrain_range (src/rle_bwt.rs:202-287) once per symbol.
// Integer-only, HBM/L2 random-gather bound; no tensor cores (nothing here is a
// dense contraction).  See layout.h for the block format.
#include "engine.h"

namespace msbwt {

// ---------------------------------------------------------------- device helpers

__device__ __forceinline__ uint4 ldg_block_chunk(const uint4 *p) {
    // read-only path, 128-bit per lane; 8 adjacent lanes cover one 128-byte block
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}

__device__ __forceinline__ uint32_t below_mask(int nbits) {
    // mask of the low `nbits` bits, nbits clamped to [0,32]
    uint32_t m;
    int w = max(nbits, 0);
    asm("bmsk.clamp.b32 %0, %1, %2;" : "=r"(m) : "r"(0), "r"(w));
    return m;
}

// Occurrences, within this lane's 32 symbols, of the symbol selected by the
// (x0,x1,x2) plane-inversion masks, restricted to block offsets < p.
__device__ __forceinline__ uint32_t lane_count(const uint4 &c, uint32_t x0, uint32_t x1, uint32_t x2,
                                               uint32_t p, uint32_t sub) {
    uint32_t m = (c.y ^ x0) & (c.z ^ x1) & (c.w ^ x2);
    return __popc(m & below_mask((int)p - (int)(sub << 5)));
}

// One constrain_range for the 8-lane group this thread belongs to.  All 32 lanes of
// the warp must call it together (the shuffles are warp-wide, segmented by 8).
// `live` gates the loads; dead groups compute garbage that the caller discards.
__device__ __forceinline__ void group_step(const IndexView &ix, const uint64_t *cb, uint32_t sub,
                                           bool live, uint32_t sym, uint64_t &l, uint64_t &h) {
    const uint64_t bl = l >> kBlockShift, bh = h >> kBlockShift;
    uint4 cl = make_uint4(0, 0, 0, 0), ch;
    if (live) cl = ldg_block_chunk(ix.blocks + bl * kLanesPerBlock + sub);
    ch = cl;
    if (live && bh != bl) ch = ldg_block_chunk(ix.blocks + bh * kLanesPerBlock + sub);

    const uint32_t x0 = (sym & 1u) ? 0u : ~0u, x1 = (sym & 2u) ? 0u : ~0u, x2 = (sym & 4u) ? 0u : ~0u;
    uint32_t cnt = lane_count(cl, x0, x1, x2, (uint32_t)l & (kBlockSyms - 1), sub) |
                   (lane_count(ch, x0, x1, x2, (uint32_t)h & (kBlockSyms - 1), sub) << 16);
    cnt += __shfl_xor_sync(0xffffffffu, cnt, 1);
    cnt += __shfl_xor_sync(0xffffffffu, cnt, 2);
    cnt += __shfl_xor_sync(0xffffffffu, cnt, 4);
    const uint32_t hl = __shfl_sync(0xffffffffu, cl.x, sym, kLanesPerBlock);
    const uint32_t hh = __shfl_sync(0xffffffffu, ch.x, sym, kLanesPerBlock);
    const uint64_t base_l = cb[((bl >> ix.sb_shift) << 3) + sym];
    const uint64_t base_h = cb[((bh >> ix.sb_shift) << 3) + sym];
    l = base_l + hl + (cnt & 0xffffu);
    h = base_h + hh + (cnt >> 16);
}

// copy cbase into shared memory when it fits; returns the pointer to use
__device__ __forceinline__ const uint64_t *stage_cbase(const IndexView &ix, uint64_t *smem) {
    if (ix.n_super > (uint32_t)kMaxSuperInSmem) return ix.cbase;
    for (uint32_t i = threadIdx.x; i < ix.n_super * 8u; i += blockDim.x) smem[i] = ix.cbase[i];
    __syncthreads();
    return smem;
}

// ---------------------------------------------------------------- K0: pack + validate

// One thread per (query, word).  Word w of query q holds symbols consumed at steps
// 21w .. 21w+20 of the backward search (step t reads kmer[k-1-t]), first step in the
// top 3 bits below bit 63.  Stored word-major: packed[w*n + q].
__global__ void pack_fixed_kernel(const uint8_t *__restrict__ syms, uint32_t k, uint64_t n,
                                  uint32_t words, uint64_t *__restrict__ packed,
                                  uint32_t *__restrict__ status) {
    const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= n * words) return;
    const uint64_t q = tid % n;
    const uint32_t w = (uint32_t)(tid / n);
    const uint8_t *src = syms + q * k;
    const uint32_t t0 = w * kSymsPerWord;
    const uint32_t cnt = min((uint32_t)kSymsPerWord, k - t0);
    uint64_t word = 0;
    bool bad = false;
    for (uint32_t i = 0; i < cnt; i++) {
        const uint32_t s = src[k - 1 - (t0 + i)];
        bad |= (s >= (uint32_t)kAlphabet);
        word |= (uint64_t)(s & 7u) << (60 - 3 * i);
    }
    packed[(uint64_t)w * n + q] = word;
    if (bad) atomicOr(status, 1u);
}

// ---------------------------------------------------------------- K1: count_kmers

// Persistent kernel: every 8-lane group owns a stream of queries (q, q+G, q+2G, ...)
// and refills itself as soon as its current query is finished, so a warp's four
// groups never wait for each other's k-mers to end.
__global__ void __launch_bounds__(kCountThreads)
count_kmers_packed_kernel(IndexView ix, const uint64_t *__restrict__ packed, uint32_t k, uint64_t n,
                          uint64_t *__restrict__ out) {
    __shared__ uint64_t cb_smem[kMaxSuperInSmem * 8];
    const uint64_t *cb = stage_cbase(ix, cb_smem);

    const uint32_t sub = threadIdx.x & (kLanesPerBlock - 1);
    const uint64_t groups = (uint64_t)gridDim.x * (kCountThreads / kLanesPerBlock);
    uint64_t q = (uint64_t)blockIdx.x * (kCountThreads / kLanesPerBlock) + (threadIdx.x / kLanesPerBlock);

    uint64_t l = 0, h = 0, word = 0, next_word = 0;
    uint32_t rem = 0;   // symbols still to consume
    int shift = 60;     // bit offset of the next symbol in `word`
    uint32_t widx = 0;  // index of `word` within the query
    bool live = q < n;
    if (live) {
        l = 0; h = ix.total; rem = k;
        if (k) word = packed[q];
        if (q + groups < n && k) next_word = packed[q + groups];
    }

    while (__any_sync(0xffffffffu, live)) {
        // retire + refill (msbwt_core.rs:151-153,160: empty range or all symbols consumed)
        while (live && (rem == 0 || l == h)) {
            if (sub == 0) out[q] = h - l;
            q += groups;
            live = q < n;
            if (live) {
                l = 0; h = ix.total; rem = k; shift = 60; widx = 0;
                word = next_word;
                if (q + groups < n && k) next_word = packed[q + groups];
            }
        }
        __syncwarp();
        if (live && shift < 0) {  // next 21 symbols
            widx++;
            word = packed[(uint64_t)widx * n + q];
            shift = 60;
        }
        const uint32_t sym = (uint32_t)(word >> (shift & 63)) & 7u;
        uint64_t nl = l, nh = h;
        group_step(ix, cb, sub, live, sym, nl, nh);
        if (live) { l = nl; h = nh; rem--; shift -= 3; }
    }
}

// Variable-length form, symbols read straight from the caller's byte layout.
__global__ void __launch_bounds__(kCountThreads)
count_kmers_bytes_kernel(IndexView ix, const uint8_t *__restrict__ syms,
                         const uint64_t *__restrict__ offsets, uint64_t n, uint64_t *__restrict__ out,
                         uint32_t *__restrict__ status) {
    __shared__ uint64_t cb_smem[kMaxSuperInSmem * 8];
    const uint64_t *cb = stage_cbase(ix, cb_smem);

    const uint32_t sub = threadIdx.x & (kLanesPerBlock - 1);
    const uint64_t groups = (uint64_t)gridDim.x * (kCountThreads / kLanesPerBlock);
    uint64_t q = (uint64_t)blockIdx.x * (kCountThreads / kLanesPerBlock) + (threadIdx.x / kLanesPerBlock);

    uint64_t l = 0, h = 0, beg = 0, cur = 0;  // cur: one past the next symbol to consume
    bool live = q < n;
    if (live) { l = 0; h = ix.total; beg = offsets[q]; cur = offsets[q + 1]; }

    while (__any_sync(0xffffffffu, live)) {
        while (live && (cur == beg || l == h)) {
            if (sub == 0) out[q] = h - l;
            q += groups;
            live = q < n;
            if (live) { l = 0; h = ix.total; beg = offsets[q]; cur = offsets[q + 1]; }
        }
        __syncwarp();
        uint32_t sym = 0;
        if (live) {
            sym = syms[cur - 1];
            if (sym >= (uint32_t)kAlphabet) { atomicOr(status, 1u); sym = 0; }
        }
        uint64_t nl = l, nh = h;
        group_step(ix, cb, sub, live, sym, nl, nh);
        if (live) { l = nl; h = nh; cur--; }
    }
}

// ---------------------------------------------------------------- K2: constrain_ranges

__global__ void __launch_bounds__(kCountThreads)
constrain_ranges_kernel(IndexView ix, const uint8_t *__restrict__ sym, const uint64_t *__restrict__ l,
                        const uint64_t *__restrict__ h, uint64_t n, uint64_t *__restrict__ out_l,
                        uint64_t *__restrict__ out_h) {
    __shared__ uint64_t cb_smem[kMaxSuperInSmem * 8];
    const uint64_t *cb = stage_cbase(ix, cb_smem);
    const uint32_t sub = threadIdx.x & (kLanesPerBlock - 1);
    const uint64_t groups = (uint64_t)gridDim.x * (kCountThreads / kLanesPerBlock);
    const uint64_t q = (uint64_t)blockIdx.x * (kCountThreads / kLanesPerBlock) + (threadIdx.x / kLanesPerBlock);
    // all four groups of a warp must stay in the loop together (group_step shuffles warp-wide),
    // so the trip count is decided by the warp's first group
    const uint64_t warp_first = q - ((threadIdx.x / kLanesPerBlock) & 3u);
    for (uint64_t it = 0; warp_first + it * groups < n; it++) {
        const uint64_t i = q + it * groups;
        const bool live = i < n;
        uint64_t a = 0, b = 0;
        uint32_t s = 0;
        if (live) { a = l[i]; b = h[i]; s = sym[i]; }
        group_step(ix, cb, sub, live, s, a, b);
        if (live && sub == 0) { out_l[i] = a; out_h[i] = b; }
    }
}

// ---------------------------------------------------------------- K4: gather roofline

__device__ __forceinline__ uint64_t mix64(uint64_t z) {  // splitmix64 finaliser
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

// LANES lanes x 16 B = one granule.  Every group issues independent random granule
// reads, 4 in flight per lane, and xors what it read into a sink so nothing is elided.
template <int LANES>
__global__ void __launch_bounds__(256) gather_kernel(const uint4 *__restrict__ buf, uint64_t n_granules,
                                                     uint64_t n_gathers, uint64_t seed,
                                                     uint64_t *__restrict__ sink) {
    const uint32_t sub = threadIdx.x % LANES;
    const uint64_t groups = (uint64_t)gridDim.x * (256 / LANES);
    const uint64_t g = (uint64_t)blockIdx.x * (256 / LANES) + threadIdx.x / LANES;
    uint32_t acc = 0;
    for (uint64_t i = g; i < n_gathers; i += 4 * groups) {
        uint4 v[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const uint64_t j = i + (uint64_t)u * groups;
            const uint64_t gi = mix64(j ^ seed) % n_granules;
            v[u] = make_uint4(0, 0, 0, 0);
            if (j < n_gathers) v[u] = ldg_block_chunk(buf + gi * LANES + sub);
        }
#pragma unroll
        for (int u = 0; u < 4; u++) acc ^= v[u].x ^ v[u].y ^ v[u].z ^ v[u].w;
    }
    if (acc == 0x9E3779B9u) atomicAdd((unsigned long long *)sink, 1ull);  // practically never; defeats DCE
}

// ---------------------------------------------------------------- launch wrappers

static int g_sm_count[64];

static int sm_count(int device) {
    if (device < 0 || device >= 64) return 148;
    if (!g_sm_count[device]) {
        int v = 0;
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, device) != cudaSuccess || v <= 0) v = 148;
        g_sm_count[device] = v;
    }
    return g_sm_count[device];
}

static unsigned persistent_grid(int device, const void *kernel, int threads, uint64_t work_groups,
                                int groups_per_cta) {
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, 0) != cudaSuccess || per_sm < 1)
        per_sm = 1;
    uint64_t full = (uint64_t)sm_count(device) * (uint64_t)per_sm;  // one wave: a multiple of the SM count
    uint64_t need = (work_groups + groups_per_cta - 1) / groups_per_cta;
    if (need < 1) need = 1;
    return (unsigned)(need < full ? need : full);
}

cudaError_t launch_pack_fixed(const uint8_t *d_syms, uint32_t k, uint64_t n, uint64_t *d_packed,
                              uint32_t *d_status, cudaStream_t st) {
    const uint32_t words = words_for_k(k);
    const uint64_t items = n * words;
    if (!items) return cudaSuccess;
    const unsigned blocks = (unsigned)((items + 255) / 256);
    pack_fixed_kernel<<<blocks, 256, 0, st>>>(d_syms, k, n, words, d_packed, d_status);
    return cudaGetLastError();
}

cudaError_t launch_count_packed(int device, const IndexView &ix, const uint64_t *d_packed, uint32_t k,
                                uint64_t n, uint64_t *d_out, cudaStream_t st) {
    if (!n) return cudaSuccess;
    const unsigned grid = persistent_grid(device, (const void *)count_kmers_packed_kernel, kCountThreads, n,
                                          kCountThreads / kLanesPerBlock);
    count_kmers_packed_kernel<<<grid, kCountThreads, 0, st>>>(ix, d_packed, k, n, d_out);
    return cudaGetLastError();
}

cudaError_t launch_count_bytes(int device, const IndexView &ix, const uint8_t *d_syms,
                               const uint64_t *d_offsets, uint64_t n, uint64_t *d_out, uint32_t *d_status,
                               cudaStream_t st) {
    if (!n) return cudaSuccess;
    const unsigned grid = persistent_grid(device, (const void *)count_kmers_bytes_kernel, kCountThreads, n,
                                          kCountThreads / kLanesPerBlock);
    count_kmers_bytes_kernel<<<grid, kCountThreads, 0, st>>>(ix, d_syms, d_offsets, n, d_out, d_status);
    return cudaGetLastError();
}

cudaError_t launch_constrain_ranges(int device, const IndexView &ix, const uint8_t *d_sym, const uint64_t *d_l,
                                    const uint64_t *d_h, uint64_t n, uint64_t *d_out_l, uint64_t *d_out_h,
                                    cudaStream_t st) {
    if (!n) return cudaSuccess;
    const unsigned grid = persistent_grid(device, (const void *)constrain_ranges_kernel, kCountThreads, n,
                                          kCountThreads / kLanesPerBlock);
    constrain_ranges_kernel<<<grid, kCountThreads, 0, st>>>(ix, d_sym, d_l, d_h, n, d_out_l, d_out_h);
    return cudaGetLastError();
}

cudaError_t launch_gather(int device, const void *d_buf, uint64_t buf_bytes, uint32_t granule,
                          uint64_t n_gathers, uint64_t seed, uint64_t *d_sink, cudaStream_t st) {
    const uint64_t n_granules = buf_bytes / granule;
    if (!n_granules || !n_gathers) return cudaErrorInvalidValue;
    const unsigned grid = (unsigned)sm_count(device) * 8u;
    const uint4 *buf = (const uint4 *)d_buf;
    switch (granule) {
        case 32: gather_kernel<2><<<grid, 256, 0, st>>>(buf, n_granules, n_gathers, seed, d_sink); break;
        case 64: gather_kernel<4><<<grid, 256, 0, st>>>(buf, n_granules, n_gathers, seed, d_sink); break;
        case 128: gather_kernel<8><<<grid, 256, 0, st>>>(buf, n_granules, n_gathers, seed, d_sink); break;
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

}  // namespace msbwt
