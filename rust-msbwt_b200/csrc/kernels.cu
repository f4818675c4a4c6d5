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
#include <type_traits>

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

// ---------------------------------------------------------------- K0: pack + validate + seed

// Packed query layout (word-major, stride n): `words` symbol words, then the seed.
// Word w of query q holds the symbols consumed at steps 21w .. 21w+20 of the backward search
// (step t reads kmer[k-1-t]), first step in bits 62..60; bit 63 of word 0 says "the seed came
// from the suffix table, the first table_s steps are already done".  The seed is the range
// the search starts from: NARROW one word (l | h << 32), WIDE two words (l, h).
// One thread per query: pack, validate (symbol >= 6 sets *status) and, when the last
// table_s symbols are all ACGT, gather the seed range from the suffix table.
template <bool WIDE>
__global__ void pack_seed_kernel(IndexView ix, const uint8_t *__restrict__ syms, uint32_t k, uint64_t n,
                                 uint32_t words, uint64_t *__restrict__ packed, uint32_t *__restrict__ status) {
    const uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n) return;
    const uint8_t *src = syms + q * k;
    const uint32_t ts = ix.table_s;
    bool bad = false, acgt = (ts != 0 && k >= ts);
    uint64_t tidx = 0, word0 = 0;
    for (uint32_t w = 0; w < words; w++) {
        const uint32_t t0 = w * kSymsPerWord;
        const uint32_t cnt = k > t0 ? min((uint32_t)kSymsPerWord, k - t0) : 0u;
        uint64_t word = 0;
        for (uint32_t i = 0; i < cnt; i++) {
            const uint32_t sy = src[k - 1 - (t0 + i)];
            bad |= (sy >= (uint32_t)kAlphabet);
            word |= (uint64_t)(sy & 7u) << (60 - 3 * i);
            if (t0 + i < ts) {
                acgt &= ((0x2Eu >> (sy & 7u)) & 1u) != 0;                   // {1,2,3,5}
                tidx = (tidx << 2) | ((sy - 1u - (sy >> 2)) & 3u);
            }
        }
        if (w == 0) word0 = word; else packed[(uint64_t)w * n + q] = word;
    }
    uint64_t lo = 0, hi = ix.total;
    if (acgt && !bad) {
        if constexpr (WIDE) {
            const ulonglong2 e = __ldg(reinterpret_cast<const ulonglong2 *>(ix.table) + tidx);
            lo = e.x; hi = e.y;
        } else {
            const uint2 e = __ldg(reinterpret_cast<const uint2 *>(ix.table) + tidx);
            lo = e.x; hi = e.y;
        }
        word0 |= 1ull << 63;
    }
    packed[q] = word0;
    if constexpr (WIDE) {
        packed[(uint64_t)words * n + q] = lo;
        packed[(uint64_t)(words + 1) * n + q] = hi;
    } else {
        packed[(uint64_t)words * n + q] = lo | (hi << 32);
    }
    if (bad) atomicOr(status, 1u);
}

// ---------------------------------------------------------------- K1: count_kmers

constexpr int kGroupsPerCta = kCountThreads / kLanesPerBlock;

// Persistent kernel: every 4-lane group owns a stream of queries (q, q+G, q+2G, ...)
// and refills itself as soon as its current query is finished, so a warp's eight
// groups never wait for each other's k-mers to end.  The next query's first word and
// seed are loaded one query ahead so a refill never exposes memory latency.
// `packed`/`out` are already offset to this launch's first query; `stride` is the
// word-major stride of `packed`, `words` the number of symbol words per query.
template <bool WIDE>
__global__ void __launch_bounds__(kCountThreads, WIDE ? kCountMinCtasWide : kCountMinCtas)
count_kmers_packed_kernel(IndexView ix, const uint64_t *__restrict__ packed, uint64_t stride, uint32_t words,
                          uint32_t k, uint32_t n, uint64_t *__restrict__ out) {
    using P = typename Pos<WIDE>::type;
    __shared__ uint64_t cb_smem[WIDE ? kMaxSuperInSmem * 8 : 4];
    const CBase<WIDE> cb = stage_cbase<WIDE>(ix, cb_smem);
    const uint64_t keep = policy_evict_last(), stream = policy_evict_first();

    const uint32_t sub = threadIdx.x & (kLanesPerBlock - 1);
    const uint32_t groups = gridDim.x * kGroupsPerCta;
    uint32_t q = blockIdx.x * kGroupsPerCta + (threadIdx.x / kLanesPerBlock);
    const uint64_t *seeds = packed + (uint64_t)words * stride;
    const uint32_t ts = ix.table_s;

    P l = 0, h = 0;
    uint64_t word = 0, next_word = 0, next_lo = 0;
    [[maybe_unused]] uint64_t next_hi = 0;
    uint32_t rem = 0;   // symbols still to consume
    int shift = 60;     // bit offset of the next symbol in `word`
    uint32_t widx = 0;  // index of `word` within the query

    auto prefetch = [&](uint32_t qq) {
        next_word = ldg_stream(packed + qq, stream);
        next_lo = ldg_stream(seeds + qq, stream);
        if constexpr (WIDE) next_hi = ldg_stream(seeds + stride + qq, stream);
    };
    auto begin = [&]() {  // start the prefetched query
        word = next_word;
        if constexpr (WIDE) { l = next_lo; h = next_hi; } else { l = (uint32_t)next_lo; h = (uint32_t)(next_lo >> 32); }
        const uint32_t done = (word >> 63) ? ts : 0u;  // steps already answered by the suffix table
        rem = k - done;
        shift = 60 - 3 * (int)done;
        widx = 0;
    };

    bool live = q < n;
    if (live) {
        prefetch(q);
        begin();
        if (q + groups < n) prefetch(q + groups);
    }

    while (__any_sync(0xffffffffu, live)) {
        // retire + refill (msbwt_core.rs:151-153,160: empty range or all symbols consumed)
        while (live && (rem == 0 || l == h)) {
            if (sub == 0) stg_stream(out + q, (uint64_t)(h - l), stream);
            q += groups;
            live = q < n;
            if (live) {
                begin();
                if (q + groups < n) prefetch(q + groups);
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

// ---------------------------------------------------------------- K3: suffix table levels

// child[idx] = constrain_range(ACGT[idx & 3], parent[idx >> 2]); an empty parent stays empty
// (count_kmer returns 0 as soon as the range is empty, msbwt_core.rs:151-153).
template <bool WIDE>
__global__ void __launch_bounds__(kCountThreads, kCountMinCtas)
table_extend_kernel(IndexView ix, const void *__restrict__ parent_v, void *__restrict__ child_v, uint32_t n_child) {
    using P = typename Pos<WIDE>::type;
    using E = typename std::conditional<WIDE, ulonglong2, uint2>::type;
    const E *parent = reinterpret_cast<const E *>(parent_v);
    E *child = reinterpret_cast<E *>(child_v);
    __shared__ uint64_t cb_smem[WIDE ? kMaxSuperInSmem * 8 : 4];
    const CBase<WIDE> cb = stage_cbase<WIDE>(ix, cb_smem);
    const uint64_t keep = policy_evict_last();
    const uint32_t sub = threadIdx.x & (kLanesPerBlock - 1);
    const uint32_t groups = gridDim.x * kGroupsPerCta;
    const uint32_t g = blockIdx.x * kGroupsPerCta + (threadIdx.x / kLanesPerBlock);
    const uint32_t warp_first = g - ((threadIdx.x / kLanesPerBlock) & 7u);
    for (uint64_t it = 0; warp_first + it * groups < n_child; it++) {
        const uint64_t i = g + it * groups;
        bool live = i < n_child;
        P a = 0, b = 0;
        if (live) {
            const E e = parent[i >> 2];
            a = (P)e.x; b = (P)e.y;
        }
        const bool empty = (a == b);
        const uint32_t sym = (0x5321u >> ((i & 3u) * 4u)) & 7u;  // A,C,G,T = 1,2,3,5
        group_step<WIDE>(ix, cb, keep, sub, live && !empty, sym, a, b);
        if (live && sub == 0) {
            E e;
            e.x = empty ? 0 : a; e.y = empty ? 0 : b;
            child[i] = e;
        }
    }
}

// ---------------------------------------------------------------- K4: gather roofline

__device__ __forceinline__ uint32_t mix32(uint32_t x) {  // murmur3 finaliser
    x ^= x >> 16; x *= 0x85EBCA6Bu;
    x ^= x >> 13; x *= 0xC2B2AE35u;
    return x ^ (x >> 16);
}

// LANES lanes x 16 B = one granule.  Every group issues independent random granule
// reads, 8 in flight per lane, and xors what it read into a sink so nothing is elided.
// The index stream is a 32-bit hash masked to a power-of-two granule count so that the
// kernel stays far from issue-bound (a 64-bit modulo here would dominate).
template <int LANES>
__global__ void __launch_bounds__(256) gather_kernel(const uint4 *__restrict__ buf, uint32_t granule_mask,
                                                     uint32_t n_gathers, uint32_t seed,
                                                     uint64_t *__restrict__ sink) {
    constexpr int kInFlight = 8;
    const uint32_t sub = threadIdx.x % LANES;
    const uint32_t groups = gridDim.x * (256 / LANES);
    const uint32_t g = blockIdx.x * (256 / LANES) + threadIdx.x / LANES;
    uint32_t acc = 0;
    for (uint32_t i = g; i < n_gathers; i += kInFlight * groups) {
        uint4 v[kInFlight];
#pragma unroll
        for (int u = 0; u < kInFlight; u++) {
            const uint32_t j = i + (uint32_t)u * groups;
            const uint32_t gi = mix32(j ^ seed) & granule_mask;
            v[u] = make_uint4(0, 0, 0, 0);
            if (j < n_gathers) v[u] = ldg_plain(buf + (size_t)gi * LANES + sub);
        }
#pragma unroll
        for (int u = 0; u < kInFlight; u++) acc ^= v[u].x ^ v[u].y ^ v[u].z ^ v[u].w;
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

cudaError_t launch_pack_seed(const IndexView &ix, const uint8_t *d_syms, uint32_t k, uint64_t n, uint64_t *d_packed,
                             uint32_t *d_status, cudaStream_t st) {
    if (!n) return cudaSuccess;
    const uint32_t words = words_for_k(k);
    const unsigned blocks = (unsigned)((n + 255) / 256);
    if (is_wide(ix)) pack_seed_kernel<true><<<blocks, 256, 0, st>>>(ix, d_syms, k, n, words, d_packed, d_status);
    else pack_seed_kernel<false><<<blocks, 256, 0, st>>>(ix, d_syms, k, n, words, d_packed, d_status);
    return cudaGetLastError();
}

cudaError_t launch_count_packed(int device, const IndexView &ix, const uint64_t *d_packed, uint32_t k,
                                uint64_t n, uint64_t *d_out, cudaStream_t st, int *launches) {
    const uint32_t words = words_for_k(k);
    for (uint64_t q0 = 0; q0 < n; q0 += kMaxPerLaunch) {
        const uint32_t m = (uint32_t)((n - q0) < kMaxPerLaunch ? (n - q0) : kMaxPerLaunch);
        if (is_wide(ix)) {
            const unsigned grid = persistent_grid(device, (const void *)count_kmers_packed_kernel<true>, kCountThreads, m, kGroupsPerCta);
            count_kmers_packed_kernel<true><<<grid, kCountThreads, 0, st>>>(ix, d_packed + q0, n, words, k, m, d_out + q0);
        } else {
            const unsigned grid = persistent_grid(device, (const void *)count_kmers_packed_kernel<false>, kCountThreads, m, kGroupsPerCta);
            count_kmers_packed_kernel<false><<<grid, kCountThreads, 0, st>>>(ix, d_packed + q0, n, words, k, m, d_out + q0);
        }
        if (launches) (*launches)++;
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

cudaError_t launch_table_extend(int device, const IndexView &ix, const void *d_parent, void *d_child,
                                uint32_t n_child, cudaStream_t st) {
    if (is_wide(ix)) {
        const unsigned grid = persistent_grid(device, (const void *)table_extend_kernel<true>, kCountThreads, n_child, kGroupsPerCta);
        table_extend_kernel<true><<<grid, kCountThreads, 0, st>>>(ix, d_parent, d_child, n_child);
    } else {
        const unsigned grid = persistent_grid(device, (const void *)table_extend_kernel<false>, kCountThreads, n_child, kGroupsPerCta);
        table_extend_kernel<false><<<grid, kCountThreads, 0, st>>>(ix, d_parent, d_child, n_child);
    }
    return cudaGetLastError();
}

bool index_is_wide(const IndexView &ix) { return is_wide(ix); }

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
    uint64_t n_granules = buf_bytes / granule;
    if (!n_granules || !n_gathers || n_gathers >= (1ull << 31)) return cudaErrorInvalidValue;
    while (n_granules & (n_granules - 1)) n_granules &= n_granules - 1;  // round down to a power of two
    if (n_granules > (1ull << 31)) n_granules = 1ull << 31;
    const uint32_t mask = (uint32_t)(n_granules - 1);
    const unsigned grid = (unsigned)sm_count(device) * 8u;
    const uint4 *buf = (const uint4 *)d_buf;
    switch (granule) {
        case 32: gather_kernel<2><<<grid, 256, 0, st>>>(buf, mask, (uint32_t)n_gathers, (uint32_t)seed, d_sink); break;
        case 64: gather_kernel<4><<<grid, 256, 0, st>>>(buf, mask, (uint32_t)n_gathers, (uint32_t)seed, d_sink); break;
        case 128: gather_kernel<8><<<grid, 256, 0, st>>>(buf, mask, (uint32_t)n_gathers, (uint32_t)seed, d_sink); break;
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

}  // namespace msbwt
