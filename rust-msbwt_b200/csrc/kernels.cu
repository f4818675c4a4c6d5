// kernels.cu -- sm_100a kernels for the batched backward search.
//
// Replaces the reference's hot loop: BWT::count_kmer (src/msbwt_core.rs:125-161)
// calling RleBWT::constrain_range (src/rle_bwt.rs:202-287) once per symbol.
// Integer-only, random-gather (HBM sector / L2) bound; no tensor cores (nothing here
// is a dense contraction).  See layout.h for the block format.
//
// Work mapping: 4 adjacent lanes form a group that owns one query at a time; the group
// fetches a 64-byte block with one 128-bit ld.global.nc per lane, each lane matches its
// 32 symbols against the query symbol (3 LOP3 + POPC), and two shuffle-xor steps sum
// the lanes (the l and h boundaries share the reduction, 16 bits each).  A warp
// therefore advances 8 queries per instruction.
#include "engine.h"

namespace msbwt {

// ---------------------------------------------------------------- device helpers

__device__ __forceinline__ uint64_t policy_evict_last() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t policy_evict_first() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}

// index block chunk: read-only path, no L1 allocation, keep in L2
__device__ __forceinline__ uint4 ldg_index(const uint4 *p, uint64_t pol) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p), "l"(pol));
    return r;
}
__device__ __forceinline__ uint4 ldg_plain(const uint4 *p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}
// streaming data (packed queries, results): do not let it displace the index in L2
__device__ __forceinline__ uint64_t ldg_stream(const uint64_t *p, uint64_t pol) {
    uint64_t r;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.u64 %0, [%1], %2;" : "=l"(r) : "l"(p), "l"(pol));
    return r;
}
__device__ __forceinline__ void stg_stream(uint64_t *p, uint64_t v, uint64_t pol) {
    asm volatile("st.global.L1::no_allocate.L2::cache_hint.u64 [%0], %1, %2;" ::"l"(p), "l"(v), "l"(pol) : "memory");
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

template <bool WIDE> struct Pos { using type = uint32_t; };
template <> struct Pos<true> { using type = uint64_t; };

// Per-CTA constants: C array (+ superblock bases).  NARROW (N < 2^32, one superblock):
// 8 x u32 in shared memory.  WIDE: u64 rows per superblock, shared memory when they fit.
template <bool WIDE> struct CBase;
template <> struct CBase<false> {
    const uint32_t *c;
    __device__ __forceinline__ uint32_t at(uint32_t, uint32_t sym) const { return c[sym]; }
};
template <> struct CBase<true> {
    const uint64_t *c;
    uint32_t sb_shift;
    __device__ __forceinline__ uint64_t at(uint64_t blk, uint32_t sym) const { return c[((blk >> sb_shift) << 3) + sym]; }
};

template <bool WIDE>
__device__ __forceinline__ CBase<WIDE> stage_cbase(const IndexView &ix, uint64_t *smem) {
    if constexpr (WIDE) {
        if (ix.n_super > (uint32_t)kMaxSuperInSmem) return CBase<true>{ix.cbase, ix.sb_shift};
        for (uint32_t i = threadIdx.x; i < ix.n_super * 8u; i += blockDim.x) smem[i] = ix.cbase[i];
        __syncthreads();
        return CBase<true>{smem, ix.sb_shift};
    } else {
        uint32_t *s32 = reinterpret_cast<uint32_t *>(smem);
        if (threadIdx.x < 8) s32[threadIdx.x] = (uint32_t)ix.cbase[threadIdx.x];
        __syncthreads();
        return CBase<false>{s32};
    }
}

// One constrain_range for the 4-lane group this thread belongs to.  All 32 lanes of
// the warp must call it together (the shuffles are warp-wide, segmented by 4).
// `live` gates the loads; dead groups compute garbage that the caller discards.
template <bool WIDE>
__device__ __forceinline__ void group_step(const IndexView &ix, const CBase<WIDE> &cb, uint64_t keep, uint32_t sub,
                                           bool live, uint32_t sym, typename Pos<WIDE>::type &l,
                                           typename Pos<WIDE>::type &h) {
    using P = typename Pos<WIDE>::type;
    const P bl = l >> kBlockShift, bh = h >> kBlockShift;
    uint4 cl = make_uint4(0, 0, 0, 0), ch;
    if (live) cl = ldg_index(ix.blocks + (size_t)bl * kLanesPerBlock + sub, keep);
    ch = cl;
    if (live && bh != bl) ch = ldg_index(ix.blocks + (size_t)bh * kLanesPerBlock + sub, keep);

    const uint32_t x0 = (sym & 1u) ? 0u : ~0u, x1 = (sym & 2u) ? 0u : ~0u, x2 = (sym & 4u) ? 0u : ~0u;
    uint32_t cnt = lane_count(cl, x0, x1, x2, (uint32_t)l & (kBlockSyms - 1), sub) |
                   (lane_count(ch, x0, x1, x2, (uint32_t)h & (kBlockSyms - 1), sub) << 16);
    cnt += __shfl_xor_sync(0xffffffffu, cnt, 1);
    cnt += __shfl_xor_sync(0xffffffffu, cnt, 2);
    const uint32_t src = (sym - 1u - (sym >> 2)) & 3u;  // ckpt_lane(sym) for A,C,G,T
    uint32_t hl = __shfl_sync(0xffffffffu, cl.x, src, kLanesPerBlock);
    uint32_t hh = __shfl_sync(0xffffffffu, ch.x, src, kLanesPerBlock);
    if ((0x11u >> sym) & 1u) {  // $ or N: checkpoints live in the side array
        if (live) {
            hl = __ldg(ix.aux + (size_t)bl * 2 + (sym >> 2));
            hh = __ldg(ix.aux + (size_t)bh * 2 + (sym >> 2));
        }
    }
    l = cb.at(bl, sym) + hl + (cnt & 0xffffu);
    h = cb.at(bh, sym) + hh + (cnt >> 16);
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

constexpr int kGroupsPerCta = kCountThreads / kLanesPerBlock;

// Persistent kernel: every 4-lane group owns a stream of queries (q, q+G, q+2G, ...)
// and refills itself as soon as its current query is finished, so a warp's eight
// groups never wait for each other's k-mers to end.  `packed`/`out` are already offset
// to this launch's first query; `stride` is the word-major stride of `packed`.
template <bool WIDE>
__global__ void __launch_bounds__(kCountThreads, kCountMinCtas)
count_kmers_packed_kernel(IndexView ix, const uint64_t *__restrict__ packed, uint64_t stride, uint32_t k,
                          uint32_t n, uint64_t *__restrict__ out) {
    using P = typename Pos<WIDE>::type;
    __shared__ uint64_t cb_smem[WIDE ? kMaxSuperInSmem * 8 : 4];
    const CBase<WIDE> cb = stage_cbase<WIDE>(ix, cb_smem);
    const uint64_t keep = policy_evict_last(), stream = policy_evict_first();

    const uint32_t sub = threadIdx.x & (kLanesPerBlock - 1);
    const uint32_t groups = gridDim.x * kGroupsPerCta;
    uint32_t q = blockIdx.x * kGroupsPerCta + (threadIdx.x / kLanesPerBlock);

    P l = 0, h = 0;
    uint64_t word = 0, next_word = 0;
    uint32_t rem = 0;   // symbols still to consume
    int shift = 60;     // bit offset of the next symbol in `word`
    uint32_t widx = 0;  // index of `word` within the query
    bool live = q < n;
    if (live) {
        h = (P)ix.total; rem = k;
        if (k) word = ldg_stream(packed + q, stream);
        if (k && q + groups < n) next_word = ldg_stream(packed + q + groups, stream);
    }

    while (__any_sync(0xffffffffu, live)) {
        // retire + refill (msbwt_core.rs:151-153,160: empty range or all symbols consumed)
        while (live && (rem == 0 || l == h)) {
            if (sub == 0) stg_stream(out + q, (uint64_t)(h - l), stream);
            q += groups;
            live = q < n;
            if (live) {
                l = 0; h = (P)ix.total; rem = k; shift = 60; widx = 0;
                word = next_word;
                if (k && q + groups < n) next_word = ldg_stream(packed + q + groups, stream);
            }
        }
        __syncwarp();
        if (live && shift < 0) {  // next 21 symbols
            widx++;
            word = ldg_stream(packed + (uint64_t)widx * stride + q, stream);
            shift = 60;
        }
        const uint32_t sym = (uint32_t)(word >> (shift & 63)) & 7u;
        P nl = l, nh = h;
        group_step<WIDE>(ix, cb, keep, sub, live, sym, nl, nh);
        if (live) { l = nl; h = nh; rem--; shift -= 3; }
    }
}

// Variable-length form, symbols read straight from the caller's byte layout.
template <bool WIDE>
__global__ void __launch_bounds__(kCountThreads, kCountMinCtas)
count_kmers_bytes_kernel(IndexView ix, const uint8_t *__restrict__ syms, const uint64_t *__restrict__ offsets,
                         uint32_t n, uint64_t *__restrict__ out, uint32_t *__restrict__ status) {
    using P = typename Pos<WIDE>::type;
    __shared__ uint64_t cb_smem[WIDE ? kMaxSuperInSmem * 8 : 4];
    const CBase<WIDE> cb = stage_cbase<WIDE>(ix, cb_smem);
    const uint64_t keep = policy_evict_last();

    const uint32_t sub = threadIdx.x & (kLanesPerBlock - 1);
    const uint32_t groups = gridDim.x * kGroupsPerCta;
    uint32_t q = blockIdx.x * kGroupsPerCta + (threadIdx.x / kLanesPerBlock);

    P l = 0, h = 0;
    uint64_t beg = 0, cur = 0;  // cur: one past the next symbol to consume
    bool live = q < n;
    if (live) { h = (P)ix.total; beg = offsets[q]; cur = offsets[q + 1]; }

    while (__any_sync(0xffffffffu, live)) {
        while (live && (cur == beg || l == h)) {
            if (sub == 0) out[q] = (uint64_t)(h - l);
            q += groups;
            live = q < n;
            if (live) { l = 0; h = (P)ix.total; beg = offsets[q]; cur = offsets[q + 1]; }
        }
        __syncwarp();
        uint32_t sym = 0;
        if (live) {
            sym = syms[cur - 1];
            if (sym >= (uint32_t)kAlphabet) { atomicOr(status, 1u); sym = 0; }
        }
        P nl = l, nh = h;
        group_step<WIDE>(ix, cb, keep, sub, live, sym, nl, nh);
        if (live) { l = nl; h = nh; cur--; }
    }
}

// ---------------------------------------------------------------- K2: constrain_ranges

template <bool WIDE>
__global__ void __launch_bounds__(kCountThreads, kCountMinCtas)
constrain_ranges_kernel(IndexView ix, const uint8_t *__restrict__ sym, const uint64_t *__restrict__ l,
                        const uint64_t *__restrict__ h, uint32_t n, uint64_t *__restrict__ out_l,
                        uint64_t *__restrict__ out_h) {
    using P = typename Pos<WIDE>::type;
    __shared__ uint64_t cb_smem[WIDE ? kMaxSuperInSmem * 8 : 4];
    const CBase<WIDE> cb = stage_cbase<WIDE>(ix, cb_smem);
    const uint64_t keep = policy_evict_last();
    const uint32_t sub = threadIdx.x & (kLanesPerBlock - 1);
    const uint32_t groups = gridDim.x * kGroupsPerCta;
    const uint32_t q = blockIdx.x * kGroupsPerCta + (threadIdx.x / kLanesPerBlock);
    // all eight groups of a warp must stay in the loop together (group_step shuffles warp-wide),
    // so the trip count is decided by the warp's first group
    const uint32_t warp_first = q - ((threadIdx.x / kLanesPerBlock) & 7u);
    for (uint64_t it = 0; warp_first + it * groups < n; it++) {
        const uint64_t i = q + it * groups;
        const bool live = i < n;
        P a = 0, b = 0;
        uint32_t s = 0;
        if (live) { a = (P)l[i]; b = (P)h[i]; s = sym[i]; }
        group_step<WIDE>(ix, cb, keep, sub, live, s, a, b);
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
            if (j < n_gathers) v[u] = ldg_plain(buf + gi * LANES + sub);
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

// one full wave of CTAs (a multiple of the SM count), fewer if there is less work
static unsigned persistent_grid(int device, const void *kernel, int threads, uint64_t work_groups,
                                int groups_per_cta) {
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, 0) != cudaSuccess || per_sm < 1)
        per_sm = 1;
    uint64_t full = (uint64_t)sm_count(device) * (uint64_t)per_sm;
    uint64_t need = (work_groups + groups_per_cta - 1) / groups_per_cta;
    if (need < 1) need = 1;
    return (unsigned)(need < full ? need : full);
}

static bool is_wide(const IndexView &ix) { return ix.n_super > 1 || (ix.total >> 32) != 0; }

constexpr uint64_t kMaxPerLaunch = 1ull << 30;  // keeps q + groups inside u32

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
                                uint64_t n, uint64_t *d_out, cudaStream_t st, int *launches) {
    for (uint64_t q0 = 0; q0 < n; q0 += kMaxPerLaunch) {
        const uint32_t m = (uint32_t)((n - q0) < kMaxPerLaunch ? (n - q0) : kMaxPerLaunch);
        if (is_wide(ix)) {
            const unsigned grid = persistent_grid(device, (const void *)count_kmers_packed_kernel<true>, kCountThreads, m, kGroupsPerCta);
            count_kmers_packed_kernel<true><<<grid, kCountThreads, 0, st>>>(ix, d_packed + q0, n, k, m, d_out + q0);
        } else {
            const unsigned grid = persistent_grid(device, (const void *)count_kmers_packed_kernel<false>, kCountThreads, m, kGroupsPerCta);
            count_kmers_packed_kernel<false><<<grid, kCountThreads, 0, st>>>(ix, d_packed + q0, n, k, m, d_out + q0);
        }
        if (launches) (*launches)++;
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

cudaError_t launch_count_bytes(int device, const IndexView &ix, const uint8_t *d_syms,
                               const uint64_t *d_offsets, uint64_t n, uint64_t *d_out, uint32_t *d_status,
                               cudaStream_t st, int *launches) {
    for (uint64_t q0 = 0; q0 < n; q0 += kMaxPerLaunch) {
        const uint32_t m = (uint32_t)((n - q0) < kMaxPerLaunch ? (n - q0) : kMaxPerLaunch);
        if (is_wide(ix)) {
            const unsigned grid = persistent_grid(device, (const void *)count_kmers_bytes_kernel<true>, kCountThreads, m, kGroupsPerCta);
            count_kmers_bytes_kernel<true><<<grid, kCountThreads, 0, st>>>(ix, d_syms, d_offsets + q0, m, d_out + q0, d_status);
        } else {
            const unsigned grid = persistent_grid(device, (const void *)count_kmers_bytes_kernel<false>, kCountThreads, m, kGroupsPerCta);
            count_kmers_bytes_kernel<false><<<grid, kCountThreads, 0, st>>>(ix, d_syms, d_offsets + q0, m, d_out + q0, d_status);
        }
        if (launches) (*launches)++;
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

cudaError_t launch_constrain_ranges(int device, const IndexView &ix, const uint8_t *d_sym, const uint64_t *d_l,
                                    const uint64_t *d_h, uint64_t n, uint64_t *d_out_l, uint64_t *d_out_h,
                                    cudaStream_t st, int *launches) {
    for (uint64_t q0 = 0; q0 < n; q0 += kMaxPerLaunch) {
        const uint32_t m = (uint32_t)((n - q0) < kMaxPerLaunch ? (n - q0) : kMaxPerLaunch);
        if (is_wide(ix)) {
            const unsigned grid = persistent_grid(device, (const void *)constrain_ranges_kernel<true>, kCountThreads, m, kGroupsPerCta);
            constrain_ranges_kernel<true><<<grid, kCountThreads, 0, st>>>(ix, d_sym + q0, d_l + q0, d_h + q0, m, d_out_l + q0, d_out_h + q0);
        } else {
            const unsigned grid = persistent_grid(device, (const void *)constrain_ranges_kernel<false>, kCountThreads, m, kGroupsPerCta);
            constrain_ranges_kernel<false><<<grid, kCountThreads, 0, st>>>(ix, d_sym + q0, d_l + q0, d_h + q0, m, d_out_l + q0, d_out_h + q0);
        }
        if (launches) (*launches)++;
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
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
