// kernels.cu -- sm_100a kernels for the batched backward search.
//
// Replaces the reference's hot loop: BWT::count_kmer (src/msbwt_core.rs:125-161)
// calling RleBWT::constrain_range (src/rle_bwt.rs:202-287) once per symbol.
// Integer-only, random-gather (HBM access rate / L2) bound; no tensor cores (nothing
// here is a dense contraction).  See layout.h for the block format.
//
// Two work mappings over the same block image (template parameter LANES):
//   LANES = 1  one thread per query: the thread reads both 32-byte halves of a block with
//              two 256-bit ld.global.nc, matches 4 x 32 symbols (3 LOP3 each), masks and
//              popcounts.  No shuffles.  ~4 warp instructions per query-step: the mapping
//              for an L2-resident index, where issue slots are the limit.
//   LANES = 2  a lane pair per query: each lane reads ONE half (one coalesced 64-byte
//              request per block -- HBM serves ~39 G random requests/s whatever their
//              size, so one request per rank matters more than instruction count), counts
//              its 64 symbols, and three shuffles combine the pair.
// In both, a lane (pair) whose k-mer ends refills itself from its own query stream, so
// every lane of a warp stays busy.  (v1-v3 used 8- then 4-lane groups per query; ncu
// showed them issue-bound at ~12 warp instructions per query-step -- profiles/.)
#include <type_traits>

#include "engine.h"

namespace msbwt {

// ---------------------------------------------------------------- device helpers

__device__ __forceinline__ uint64_t policy_evict_first() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}

struct Half { uint32_t w[8]; };  // 32 bytes = one sector of a block

// half an index block: read-only path, no L1 allocation, evict-last in L2 (256-bit load)
__device__ __forceinline__ Half ldg_index256(const void *p) {
    Half r;
    asm volatile("ld.global.nc.L1::no_allocate.L2::evict_last.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r.w[0]), "=r"(r.w[1]), "=r"(r.w[2]), "=r"(r.w[3]), "=r"(r.w[4]), "=r"(r.w[5]), "=r"(r.w[6]),
                   "=r"(r.w[7])
                 : "l"(p));
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

// the 64 match bits of one half for the symbol selected by the plane-inversion masks
__device__ __forceinline__ void match_half(const Half &v, uint32_t x0, uint32_t x1, uint32_t x2, uint32_t &m0,
                                           uint32_t &m1) {
    m0 = (v.w[2] ^ x0) & (v.w[4] ^ x1) & (v.w[6] ^ x2);
    m1 = (v.w[3] ^ x0) & (v.w[5] ^ x1) & (v.w[7] ^ x2);
}

// occurrences among 64 match bits at half offsets < p (p may be <= 0 or >= 64)
__device__ __forceinline__ uint32_t count_below64(uint32_t m0, uint32_t m1, int p) {
    return __popc(m0 & below_mask(p)) + __popc(m1 & below_mask(p - 32));
}

// One constrain_range: [l,h) -> [C[sym]+rank(sym,l), C[sym]+rank(sym,h)).
// LANES == 1: the calling thread does all of it.  LANES == 2: the two lanes of a pair call
// it together with identical (sym, l, h); `half` = lane & 1.  Every non-exited lane of the warp
// reaches the shuffles in the same iteration of the caller's loop (lanes leave only by returning),
// so they use the full mask -- a per-pair mask would make ptxas emit MATCH/REDUX guards.
template <bool WIDE, int LANES>
__device__ __forceinline__ void rank_step(const IndexView &ix, const CBase<WIDE> &cb, uint32_t sym,
                                          typename Pos<WIDE>::type &l, typename Pos<WIDE>::type &h,
                                          uint32_t half = 0) {
    using P = typename Pos<WIDE>::type;
    const P bl = l >> kBlockShift, bh = h >> kBlockShift;
    const bool two = bh != bl;
    const char *base = reinterpret_cast<const char *>(ix.blocks);
    const uint32_t x0 = (sym & 1u) - 1u, x1 = ((sym >> 1) & 1u) - 1u, x2 = ((sym >> 2) & 1u) - 1u;  // 0 or ~0
    const uint32_t slot = (sym - 1u - (sym >> 2)) & 3u;  // ckpt_slot(sym) for A,C,G,T
    const int pl = (int)((uint32_t)l & (kBlockSyms - 1)), ph = (int)((uint32_t)h & (kBlockSyms - 1));
    uint32_t ckl, ckh, cl, ch;
    if constexpr (LANES == 1) {
        // issue every load before the first use: up to four 32-byte sectors in flight per thread
        const Half l0 = ldg_index256(base + (size_t)bl * kBlockBytes);
        const Half l1 = ldg_index256(base + (size_t)bl * kBlockBytes + 32);
        Half h0, h1;
        if (two) {
            h0 = ldg_index256(base + (size_t)bh * kBlockBytes);
            h1 = ldg_index256(base + (size_t)bh * kBlockBytes + 32);
        }
        uint32_t ml[4], mh[4];
        match_half(l0, x0, x1, x2, ml[0], ml[1]);
        match_half(l1, x0, x1, x2, ml[2], ml[3]);
        const uint32_t lo = (slot & 1u) ? l0.w[1] : l0.w[0], hi = (slot & 1u) ? l1.w[1] : l1.w[0];
        ckl = (slot & 2u) ? hi : lo;
        ckh = ckl;
#pragma unroll
        for (int j = 0; j < 4; j++) mh[j] = ml[j];
        if (two) {
            match_half(h0, x0, x1, x2, mh[0], mh[1]);
            match_half(h1, x0, x1, x2, mh[2], mh[3]);
            const uint32_t lo2 = (slot & 1u) ? h0.w[1] : h0.w[0], hi2 = (slot & 1u) ? h1.w[1] : h1.w[0];
            ckh = (slot & 2u) ? hi2 : lo2;
        }
        cl = count_below64(ml[0], ml[1], pl) + count_below64(ml[2], ml[3], pl - 64);
        ch = count_below64(mh[0], mh[1], ph) + count_below64(mh[2], mh[3], ph - 64);
    } else {
        const Half a = ldg_index256(base + (size_t)bl * kBlockBytes + half * 32);
        Half b;
        if (two) b = ldg_index256(base + (size_t)bh * kBlockBytes + half * 32);
        uint32_t ml0, ml1, mh0, mh1;
        match_half(a, x0, x1, x2, ml0, ml1);
        uint32_t cand_l = (slot & 1u) ? a.w[1] : a.w[0], cand_h = cand_l;
        mh0 = ml0; mh1 = ml1;
        if (two) {
            match_half(b, x0, x1, x2, mh0, mh1);
            cand_h = (slot & 1u) ? b.w[1] : b.w[0];
        }
        const int off = (int)half * 64;
        uint32_t cnt = count_below64(ml0, ml1, pl - off) | (count_below64(mh0, mh1, ph - off) << 16);
        cnt += __shfl_xor_sync(0xffffffffu, cnt, 1);
        ckl = __shfl_sync(0xffffffffu, cand_l, slot >> 1, 2);  // the half that owns this symbol's checkpoint
        ckh = __shfl_sync(0xffffffffu, cand_h, slot >> 1, 2);
        cl = cnt & 0xffffu;
        ch = cnt >> 16;
    }
    if ((0x11u >> sym) & 1u) {  // $ or N: checkpoints live in the side array
        ckl = __ldg(ix.aux + (size_t)bl * 2 + (sym >> 2));
        ckh = __ldg(ix.aux + (size_t)bh * 2 + (sym >> 2));
    }
    l = cb.at(bl, sym) + ckl + cl;
    h = cb.at(bh, sym) + ckh + ch;
}

// ---------------------------------------------------------------- K0: pack + validate + seed

// One thread per query: validate (symbol >= 6 sets *status), look the last table_s symbols up in
// the suffix table when they are all ACGT (that many steps are then already done), pack the
// REMAINING symbols 21 per u64 word (step t of the remaining search in bits 62-3(t%21) .. of word
// t/21; step order = from the k-mer's last symbol to its first), and
//   * finish the query right here when nothing is left to search (empty range -> count 0,
//     msbwt_core.rs:151-153; or no symbols left -> h-l), or
//   * append it to the compacted live list (PackedLayout) for the search kernel.
template <bool WIDE>
__global__ void pack_seed_kernel(IndexView ix, const uint8_t *__restrict__ syms, uint32_t k, PackedLayout lay,
                                 uint64_t *__restrict__ packed, uint64_t *__restrict__ out,
                                 uint32_t *__restrict__ status) {
    const uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = q < lay.n;
    const uint8_t *src = syms + (valid ? q : 0) * k;
    const uint32_t ts = ix.table_s;
    bool bad = false, acgt = valid && (ts != 0 && k >= ts);
    uint64_t tidx = 0;
    if (acgt) {  // the last ts symbols: table index + validation in one pass
        for (uint32_t t = 0; t < ts; t++) {
            const uint32_t sy = src[k - 1 - t];
            bad |= sy >= (uint32_t)kAlphabet;
            acgt &= sy < 8u && ((0x2Eu >> sy) & 1u) != 0;  // {1,2,3,5}
            tidx = (tidx << 2) | ((sy - 1u - (sy >> 2)) & 3u);
        }
    }
    uint64_t lo = 0, hi = ix.total;
    uint32_t done = 0;
    if (acgt) {
        if constexpr (WIDE) {
            const ulonglong2 e = __ldg(reinterpret_cast<const ulonglong2 *>(ix.table) + tidx);
            lo = e.x; hi = e.y;
        } else {
            const uint2 e = __ldg(reinterpret_cast<const uint2 *>(ix.table) + tidx);
            lo = e.x; hi = e.y;
        }
        done = ts;
    }
    // pack (and validate) the symbols the table did not consume
    uint64_t word0 = 0;
    if (valid) {
        const uint32_t rest = k - done;
        for (uint32_t w = 0; w * kSymsPerWord < rest; w++) {
            const uint32_t t0 = w * kSymsPerWord;
            const uint32_t cnt = min((uint32_t)kSymsPerWord, rest - t0);
            uint64_t word = 0;
            for (uint32_t i = 0; i < cnt; i++) {
                const uint32_t sy = src[k - 1 - (done + t0 + i)];
                bad |= (sy >= (uint32_t)kAlphabet);
                word |= (uint64_t)(sy & 7u) << (60 - 3 * i);
            }
            if (w == 0) word0 = word; else packed[lay.wx() + (uint64_t)(w - 1) * lay.n + q] = word;
        }
    }
    const bool finished = valid && (lo == hi || done == k);
    if (finished) out[q] = hi - lo;
    const bool live = valid && !finished;
    // warp-aggregated append to the live list
    const uint32_t mask = __ballot_sync(0xffffffffu, live);
    if (mask) {
        const uint32_t lane = threadIdx.x & 31u, leader = __ffs(mask) - 1;
        unsigned long long base = 0;
        if (lane == leader) base = atomicAdd(reinterpret_cast<unsigned long long *>(packed + lay.live()), (unsigned long long)__popc(mask));
        base = __shfl_sync(0xffffffffu, base, leader);
        if (live) {
            const uint64_t pos = base + __popc(mask & ((1u << lane) - 1u));
            packed[lay.w0() + pos] = word0 | ((uint64_t)(done != 0) << 63);
            if constexpr (WIDE) {
                packed[lay.seed() + pos] = lo;
                packed[lay.seed() + lay.n + pos] = hi;
            } else {
                packed[lay.seed() + pos] = lo | (hi << 32);
            }
            reinterpret_cast<uint32_t *>(packed + lay.qidx())[pos] = (uint32_t)q;
        }
    }
    if (bad) atomicOr(status, 1u);
}

// ---------------------------------------------------------------- K1: count_kmers

// Persistent kernel: every owner -- a thread (LANES = 1) or a lane pair (LANES = 2) -- walks its
// own stream of live queries (i, i+T, i+2T, ... of the compacted list) and refills itself as soon
// as its current k-mer is finished.  The next query's first word, seed and index are loaded one
// query ahead and the next symbol word 21 steps ahead, so neither exposes memory latency.
template <bool WIDE, int LANES>
__global__ void __launch_bounds__(kCountThreads, min_ctas(WIDE, LANES))
count_kmers_packed_kernel(IndexView ix, const uint64_t *__restrict__ packed, PackedLayout lay, uint32_t k,
                          uint64_t *__restrict__ out) {
    using P = typename Pos<WIDE>::type;
    __shared__ uint64_t cb_smem[WIDE ? kMaxSuperInSmem * 8 : 4];
    const CBase<WIDE> cb = stage_cbase<WIDE>(ix, cb_smem);
    const uint64_t stream = policy_evict_first();

    const uint32_t n = (uint32_t)packed[lay.live()];  // live queries (written by the pack kernel)
    const uint32_t tid = blockIdx.x * kCountThreads + threadIdx.x;
    const uint32_t owners = gridDim.x * kCountThreads / LANES;  // concurrent query streams
    const uint32_t half = tid & (LANES - 1);
    uint32_t i = tid / LANES;
    if (i >= n) return;
    const uint64_t *w0 = packed + lay.w0(), *seeds = packed + lay.seed(), *wx = packed + lay.wx();
    const uint32_t *qidx = reinterpret_cast<const uint32_t *>(packed + lay.qidx());
    const uint32_t ts = ix.table_s;

    P l = 0, h = 0;
    uint64_t word = 0, pend = 0, next_word = 0, next_lo = 0;
    [[maybe_unused]] uint64_t next_hi = 0;
    uint32_t q = 0, next_q = 0;
    uint32_t rem = 0;   // symbols still to consume
    int shift = 60;     // bit offset of the next symbol in `word`
    uint32_t widx = 0;  // index of `word` within the query's remaining symbols

    auto prefetch = [&](uint32_t ii) {
        next_word = ldg_stream(w0 + ii, stream);
        next_lo = ldg_stream(seeds + ii, stream);
        if constexpr (WIDE) next_hi = ldg_stream(seeds + lay.n + ii, stream);
        next_q = __ldg(qidx + ii);
    };
    auto begin = [&]() {  // start the prefetched query
        word = next_word;
        q = next_q;
        if constexpr (WIDE) { l = next_lo; h = next_hi; } else { l = (uint32_t)next_lo; h = (uint32_t)(next_lo >> 32); }
        rem = k - ((word >> 63) ? ts : 0u);  // the suffix table already answered ts steps
        shift = 60;
        widx = 0;
        if (rem > (uint32_t)kSymsPerWord) pend = ldg_stream(wx + q, stream);  // symbol word 1, needed 21 steps from now
    };

    prefetch(i);
    begin();
    if (i + owners < n) prefetch(i + owners);

    for (;;) {
        // retire + refill (msbwt_core.rs:151-153,160: empty range or all symbols consumed)
        while (rem == 0 || l == h) {
            if (half == 0) stg_stream(out + q, (uint64_t)(h - l), stream);
            i += owners;
            if (i >= n) return;
            begin();
            if (i + owners < n) prefetch(i + owners);
        }
        if (shift < 0) {  // next 21 symbols: already in flight since the previous word began
            word = pend;
            widx++;
            shift = 60;
            if (rem > (uint32_t)kSymsPerWord) pend = ldg_stream(wx + (uint64_t)widx * lay.n + q, stream);
        }
        const uint32_t sym = (uint32_t)(word >> shift) & 7u;
        rank_step<WIDE, LANES>(ix, cb, sym, l, h, half);
        rem--;
        shift -= 3;
    }
}

// Variable-length form, symbols read straight from the caller's byte layout.
template <bool WIDE>
__global__ void __launch_bounds__(kCountThreads, min_ctas(WIDE, 1))
count_kmers_bytes_kernel(IndexView ix, const uint8_t *__restrict__ syms, const uint64_t *__restrict__ offsets,
                         uint32_t n, uint64_t *__restrict__ out, uint32_t *__restrict__ status) {
    using P = typename Pos<WIDE>::type;
    __shared__ uint64_t cb_smem[WIDE ? kMaxSuperInSmem * 8 : 4];
    const CBase<WIDE> cb = stage_cbase<WIDE>(ix, cb_smem);
    const uint32_t threads = gridDim.x * kCountThreads;
    uint32_t q = blockIdx.x * kCountThreads + threadIdx.x;
    if (q >= n) return;
    P l = 0, h = (P)ix.total;
    uint64_t beg = offsets[q], cur = offsets[q + 1];  // cur: one past the next symbol to consume
    for (;;) {
        while (cur == beg || l == h) {
            out[q] = (uint64_t)(h - l);
            q += threads;
            if (q >= n) return;
            l = 0; h = (P)ix.total; beg = offsets[q]; cur = offsets[q + 1];
        }
        uint32_t sym = syms[cur - 1];
        if (sym >= (uint32_t)kAlphabet) { atomicOr(status, 1u); sym = 0; }
        rank_step<WIDE, 1>(ix, cb, sym, l, h);
        cur--;
    }
}

// ---------------------------------------------------------------- K2: constrain_ranges

template <bool WIDE>
__global__ void __launch_bounds__(kCountThreads, min_ctas(WIDE, 1))
constrain_ranges_kernel(IndexView ix, const uint8_t *__restrict__ sym, const uint64_t *__restrict__ l,
                        const uint64_t *__restrict__ h, uint32_t n, uint64_t *__restrict__ out_l,
                        uint64_t *__restrict__ out_h) {
    using P = typename Pos<WIDE>::type;
    __shared__ uint64_t cb_smem[WIDE ? kMaxSuperInSmem * 8 : 4];
    const CBase<WIDE> cb = stage_cbase<WIDE>(ix, cb_smem);
    const uint32_t threads = gridDim.x * kCountThreads;
    for (uint64_t i = blockIdx.x * kCountThreads + threadIdx.x; i < n; i += threads) {
        P a = (P)l[i], b = (P)h[i];
        rank_step<WIDE, 1>(ix, cb, sym[i], a, b);
        out_l[i] = a;
        out_h[i] = b;
    }
}

// ---------------------------------------------------------------- K3: suffix table levels

// child[idx] = constrain_range(ACGT[idx & 3], parent[idx >> 2]); an empty parent stays empty
// (count_kmer returns 0 as soon as the range is empty, msbwt_core.rs:151-153).
template <bool WIDE>
__global__ void __launch_bounds__(kCountThreads, min_ctas(WIDE, 1))
table_extend_kernel(IndexView ix, const void *__restrict__ parent_v, void *__restrict__ child_v, uint32_t n_child) {
    using P = typename Pos<WIDE>::type;
    using E = typename std::conditional<WIDE, ulonglong2, uint2>::type;
    const E *parent = reinterpret_cast<const E *>(parent_v);
    E *child = reinterpret_cast<E *>(child_v);
    __shared__ uint64_t cb_smem[WIDE ? kMaxSuperInSmem * 8 : 4];
    const CBase<WIDE> cb = stage_cbase<WIDE>(ix, cb_smem);
    const uint32_t threads = gridDim.x * kCountThreads;
    for (uint64_t i = blockIdx.x * kCountThreads + threadIdx.x; i < n_child; i += threads) {
        const E e = parent[i >> 2];
        P a = (P)e.x, b = (P)e.y;
        E o;
        o.x = 0; o.y = 0;
        if (a != b) {
            const uint32_t sym = (0x5321u >> ((i & 3u) * 4u)) & 7u;  // A,C,G,T = 1,2,3,5
            rank_step<WIDE, 1>(ix, cb, sym, a, b);
            o.x = a; o.y = b;
        }
        child[i] = o;
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


cudaError_t launch_pack_seed(const IndexView &ix, const uint8_t *d_syms, uint32_t k, uint64_t n, uint64_t *d_packed,
                             uint64_t *d_out, uint32_t *d_status, cudaStream_t st) {
    if (!n) return cudaSuccess;
    const PackedLayout lay = packed_layout(ix, k, n);
    cudaError_t e = cudaMemsetAsync(d_packed + lay.live(), 0, sizeof(uint64_t), st);
    if (e != cudaSuccess) return e;
    const unsigned blocks = (unsigned)((n + 255) / 256);
    if (is_wide(ix)) pack_seed_kernel<true><<<blocks, 256, 0, st>>>(ix, d_syms, k, lay, d_packed, d_out, d_status);
    else pack_seed_kernel<false><<<blocks, 256, 0, st>>>(ix, d_syms, k, lay, d_packed, d_out, d_status);
    return cudaGetLastError();
}

template <bool WIDE, int LANES>
static cudaError_t launch_count_packed_t(int device, const IndexView &ix, const uint64_t *d_packed,
                                         const PackedLayout &lay, uint32_t k, uint64_t *d_out, cudaStream_t st) {
    const unsigned grid = persistent_grid(device, (const void *)count_kmers_packed_kernel<WIDE, LANES>, kCountThreads,
                                          lay.n, kCountThreads / LANES);
    count_kmers_packed_kernel<WIDE, LANES><<<grid, kCountThreads, 0, st>>>(ix, d_packed, lay, k, d_out);
    return cudaGetLastError();
}

// n <= kMaxPerLaunch (the callers chunk): query indices are u32 inside the kernels
cudaError_t launch_count_packed(int device, const IndexView &ix, int lanes, const uint64_t *d_packed, uint32_t k,
                                uint64_t n, uint64_t *d_out, cudaStream_t st, int *launches) {
    if (!n) return cudaSuccess;
    if (n > kMaxPerLaunch) return cudaErrorInvalidValue;
    const PackedLayout lay = packed_layout(ix, k, n);
    cudaError_t e;
    if (is_wide(ix))
        e = lanes == 2 ? launch_count_packed_t<true, 2>(device, ix, d_packed, lay, k, d_out, st)
                       : launch_count_packed_t<true, 1>(device, ix, d_packed, lay, k, d_out, st);
    else
        e = lanes == 2 ? launch_count_packed_t<false, 2>(device, ix, d_packed, lay, k, d_out, st)
                       : launch_count_packed_t<false, 1>(device, ix, d_packed, lay, k, d_out, st);
    if (launches) (*launches)++;
    return e;
}

cudaError_t launch_table_extend(int device, const IndexView &ix, const void *d_parent, void *d_child,
                                uint32_t n_child, cudaStream_t st) {
    if (is_wide(ix)) {
        const unsigned grid = persistent_grid(device, (const void *)table_extend_kernel<true>, kCountThreads, n_child, kCountThreads);
        table_extend_kernel<true><<<grid, kCountThreads, 0, st>>>(ix, d_parent, d_child, n_child);
    } else {
        const unsigned grid = persistent_grid(device, (const void *)table_extend_kernel<false>, kCountThreads, n_child, kCountThreads);
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
            const unsigned grid = persistent_grid(device, (const void *)count_kmers_bytes_kernel<true>, kCountThreads, m, kCountThreads);
            count_kmers_bytes_kernel<true><<<grid, kCountThreads, 0, st>>>(ix, d_syms, d_offsets + q0, m, d_out + q0, d_status);
        } else {
            const unsigned grid = persistent_grid(device, (const void *)count_kmers_bytes_kernel<false>, kCountThreads, m, kCountThreads);
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
            const unsigned grid = persistent_grid(device, (const void *)constrain_ranges_kernel<true>, kCountThreads, m, kCountThreads);
            constrain_ranges_kernel<true><<<grid, kCountThreads, 0, st>>>(ix, d_sym + q0, d_l + q0, d_h + q0, m, d_out_l + q0, d_out_h + q0);
        } else {
            const unsigned grid = persistent_grid(device, (const void *)constrain_ranges_kernel<false>, kCountThreads, m, kCountThreads);
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
