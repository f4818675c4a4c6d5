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
// A third kernel walks the PAIR image (layout.h): a quad of lanes per query, one coalesced
// 128-byte line per TWO backward-search steps -- the mapping for an index that lives in HBM,
// where every L2 miss costs a whole 128-byte line fill whatever the request size.
// In all of them a lane (group) whose k-mer ends refills itself from its own query stream, so
// every lane of a warp stays busy.  (v1-v3 used 8- then 4-lane groups per query; ncu
// showed them issue-bound at ~12 warp instructions per query-step -- profiles/.)
#include <type_traits>

#include "device_rank.cuh"
#include "engine.h"

namespace msbwt {

// ---------------------------------------------------------------- K0: pack + validate + seed

// One thread per query.  The CTA first stages its 256 * k query bytes in shared memory with
// coalesced 16-byte loads (the per-thread layout is k-byte rows: read straight from global memory
// it costs one sector per byte load), then every thread
//   1. validates its k symbols (symbol >= 6 sets *status) and takes the base-4 value of its last
//      min(k, table_s) symbols,
//   2. picks the path: PAIR (list A) when the index has a pair image and the whole k-mer is ACGT --
//      the suffix-table depth is then chosen from {table_s, table_s - 1} so that an EVEN number of
//      symbols is left -- otherwise ONE-STEP (list B) with depth table_s when the last table_s
//      symbols are ACGT,
//   3. looks the starting range up, packs the REMAINING symbols (step t of the remaining search in
//      the top bits of word t/32 (A, 2 bits each) or t/21 (B, 3 bits each); step order = from the
//      k-mer's last symbol to its first), and
//   4. finishes the query right here when nothing is left to search (empty range -> count 0,
//      msbwt_core.rs:151-153; or no symbols left -> h-l), or appends it to its live list.
constexpr uint32_t kPackSmemMaxK = 160;  // 256 * k + 16 bytes of shared memory; longer k-mers read global memory

template <bool WIDE>
__global__ void __launch_bounds__(256)
pack_seed_kernel(IndexView ix, const uint8_t *__restrict__ syms, uint32_t k, PackedLayout lay,
                 uint64_t *__restrict__ packed, uint64_t *__restrict__ out, uint32_t *__restrict__ status) {
    extern __shared__ uint4 pack_smem_v[];
    uint8_t *smem = reinterpret_cast<uint8_t *>(pack_smem_v);
    const uint64_t q0 = (uint64_t)blockIdx.x * blockDim.x;
    const uint64_t q = q0 + threadIdx.x;
    const bool valid = q < lay.n;
    const uint8_t *src = syms + (valid ? q : 0) * k;
    if (k <= kPackSmemMaxK) {
        const uint8_t *g = syms + q0 * k;
        const uint32_t rows = (uint32_t)min((uint64_t)blockDim.x, lay.n - q0);
        const uint32_t bytes = rows * k;
        const uint32_t mis = (uint32_t)(reinterpret_cast<uintptr_t>(g) & 15u);  // smem mirrors the global alignment
        const uint32_t head = min(bytes, (16u - mis) & 15u);
        for (uint32_t i = threadIdx.x; i < head; i += blockDim.x) smem[mis + i] = g[i];
        const uint32_t vecs = (bytes - head) >> 4;
        const uint4 *gv = reinterpret_cast<const uint4 *>(g + head);
        uint4 *sv = reinterpret_cast<uint4 *>(smem + mis + head);
        for (uint32_t i = threadIdx.x; i < vecs; i += blockDim.x) sv[i] = ldg_plain(gv + i);
        for (uint32_t i = head + (vecs << 4) + threadIdx.x; i < bytes; i += blockDim.x) smem[mis + i] = g[i];
        __syncthreads();
        src = smem + mis + threadIdx.x * k;
    }
    const uint32_t ts = ix.table_s;
    const bool have_pair = ix.pair != nullptr;

    // 1. validate; is the whole k-mer ACGT; trailing ACGT run (at most ts symbols) as a base-4 number
    bool bad = false, all_acgt = true;
    uint32_t na = 0;
    uint64_t tidx = 0;
    if (valid) {
        for (uint32_t t = 0; t < k; t++) {
            const uint32_t sy = src[k - 1 - t];
            const bool ok = sy < 8u && ((0x2Eu >> sy) & 1u) != 0;  // {1,2,3,5}
            bad |= sy >= (uint32_t)kAlphabet;
            all_acgt &= ok;
            if (t < ts && na == t && ok) {
                tidx = (tidx << 2) | ((sy - 1u - (sy >> 2)) & 3u);
                na++;
            }
        }
    }
    // 2. path and table depth
    uint32_t done = 0;
    bool list_a = false;
    if (valid) {
        if (have_pair && all_acgt) {
            if (ts && k >= ts) done = ((k - ts) & 1u) ? ts - 1u : ts;
            else if (ts && k + 1u == ts) done = k;
            list_a = ((k - done) & 1u) == 0;
        } else if (ts && na >= ts) {
            done = ts;
        }
    }
    // 3. starting range
    uint64_t lo = 0, hi = ix.total;
    uint32_t flag = 0;
    if (done) {
        const bool full = done == ts;
        flag = full ? 1u : 2u;
        const uint64_t e = tidx >> (2u * (na - done));
        const void *tab = full ? ix.table : ix.table2;
        if constexpr (WIDE) {
            const ulonglong2 v = __ldg(reinterpret_cast<const ulonglong2 *>(tab) + e);
            lo = v.x; hi = v.y;
        } else {
            const uint2 v = __ldg(reinterpret_cast<const uint2 *>(tab) + e);
            lo = v.x; hi = v.y;
        }
    }
    const bool finished = valid && (lo == hi || done == k);
    if (finished) out[q] = hi - lo;
    const bool live = valid && !finished;
    // pack the symbols the table did not consume
    uint64_t word0 = 0;
    if (live) {
        const uint32_t rest = k - done;
        if (list_a) {
            for (uint32_t w = 0; w * kPairSymsPerWord < rest; w++) {
                const uint32_t t0 = w * kPairSymsPerWord;
                const uint32_t cnt = min((uint32_t)kPairSymsPerWord, rest - t0);
                uint64_t word = 0;
                for (uint32_t i = 0; i < cnt; i++) {
                    const uint32_t sy = src[k - 1 - (done + t0 + i)];
                    word |= (uint64_t)((sy - 1u - (sy >> 2)) & 3u) << (62 - 2 * i);
                }
                if (w == 0) word0 = word; else packed[lay.wx() + (uint64_t)(w - 1) * lay.n + q] = word;
            }
        } else {
            for (uint32_t w = 0; w * kSymsPerWord < rest; w++) {
                const uint32_t t0 = w * kSymsPerWord;
                const uint32_t cnt = min((uint32_t)kSymsPerWord, rest - t0);
                uint64_t word = 0;
                for (uint32_t i = 0; i < cnt; i++) {
                    const uint32_t sy = src[k - 1 - (done + t0 + i)];
                    word |= (uint64_t)(sy & 7u) << (60 - 3 * i);
                }
                if (w == 0) word0 = word; else packed[lay.wx() + (uint64_t)(w - 1) * lay.n + q] = word;
            }
        }
    }
    // warp-aggregated append to the live lists (A from the front, B from the back)
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t mask_a = __ballot_sync(0xffffffffu, live && list_a);
    const uint32_t mask_b = __ballot_sync(0xffffffffu, live && !list_a);
    unsigned long long *counters = reinterpret_cast<unsigned long long *>(packed + lay.live());
    uint64_t pos = 0;
    if (mask_a) {
        const uint32_t leader = __ffs(mask_a) - 1;
        unsigned long long base = 0;
        if (lane == leader) base = atomicAdd(counters, (unsigned long long)__popc(mask_a));
        base = __shfl_sync(0xffffffffu, base, leader);
        if (live && list_a) pos = base + __popc(mask_a & ((1u << lane) - 1u));
    }
    if (mask_b) {
        const uint32_t leader = __ffs(mask_b) - 1;
        unsigned long long base = 0;
        if (lane == leader) base = atomicAdd(counters + 1, (unsigned long long)__popc(mask_b));
        base = __shfl_sync(0xffffffffu, base, leader);
        if (live && !list_a) pos = lay.n - 1 - (base + __popc(mask_b & ((1u << lane) - 1u)));
    }
    if (live) {
        packed[lay.w0() + pos] = word0;
        if constexpr (WIDE) {
            packed[lay.seed() + pos] = lo;
            packed[lay.seed() + lay.n + pos] = hi;
        } else {
            packed[lay.seed() + pos] = lo | (hi << 32);
        }
        reinterpret_cast<uint32_t *>(packed + lay.qidx())[pos] = (uint32_t)q | (flag << 30);
    }
    if (bad) atomicOr(status, 1u);
}

// ---------------------------------------------------------------- K1: count_kmers

__device__ __forceinline__ uint32_t table_depth(uint32_t flag, uint32_t ts) { return flag == 0 ? 0u : (flag == 1 ? ts : ts - 1u); }

// Persistent kernel over live list B: every owner -- a thread (LANES = 1) or a lane pair (LANES = 2) --
// walks its own stream of live queries (i, i+T, i+2T, ... of the compacted list) and refills itself as
// soon as its current k-mer is finished.  The next query's first word, seed and index are loaded one
// query ahead and the next symbol word 21 steps ahead, so neither exposes memory latency.
template <bool WIDE, int LANES>
__global__ void __launch_bounds__(kCountThreads, min_ctas(WIDE, LANES))
count_kmers_packed_kernel(IndexView ix, const uint64_t *__restrict__ packed, PackedLayout lay, uint32_t k,
                          uint64_t *__restrict__ out) {
    using P = typename Pos<WIDE>::type;
    __shared__ uint64_t cb_smem[WIDE ? kMaxSuperInSmem * 8 : 4];
    const CBase<WIDE> cb = stage_cbase<WIDE>(ix, cb_smem);
    const uint64_t stream = policy_evict_first();

    const uint32_t n = (uint32_t)packed[lay.live() + 1];  // live queries of list B (written by the pack kernel)
    const uint32_t tid = blockIdx.x * kCountThreads + threadIdx.x;
    const uint32_t owners = gridDim.x * kCountThreads / LANES;  // concurrent query streams
    const uint32_t half = tid & (LANES - 1);
    uint32_t i = tid / LANES;
    if (i >= n) return;
    const uint32_t last = (uint32_t)lay.n - 1u;  // list B is stored back to front
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
        const uint32_t at = last - ii;
        next_word = ldg_stream(w0 + at, stream);
        next_lo = ldg_stream(seeds + at, stream);
        if constexpr (WIDE) next_hi = ldg_stream(seeds + lay.n + at, stream);
        next_q = __ldg(qidx + at);
    };
    auto begin = [&]() {  // start the prefetched query
        word = next_word;
        q = next_q & kQidxMask;
        if constexpr (WIDE) { l = next_lo; h = next_hi; } else { l = (uint32_t)next_lo; h = (uint32_t)(next_lo >> 32); }
        rem = k - table_depth(next_q >> 30, ts);  // the suffix table already answered that many steps
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

// Persistent kernel over live list A: a quad of lanes per query walks the PAIR image, two symbols per
// step (one 128-byte line per boundary); same refill / prefetch structure as above.
constexpr int pair_min_ctas(bool wide) { return wide ? 4 : 6; }

template <bool WIDE>
__global__ void __launch_bounds__(kCountThreads, pair_min_ctas(WIDE))
count_kmers_pair_kernel(IndexView ix, const uint64_t *__restrict__ packed, PackedLayout lay, uint32_t k,
                        uint64_t *__restrict__ out) {
    using P = typename Pos<WIDE>::type;
    __shared__ uint64_t c2_smem[WIDE ? kMaxSuperInSmem * 16 : 1];
    const C2Base<WIDE> c2 = stage_c2base<WIDE>(ix, c2_smem);
    const uint64_t stream = policy_evict_first();

    const uint32_t n = (uint32_t)packed[lay.live()];  // live queries of list A
    const uint32_t tid = blockIdx.x * kCountThreads + threadIdx.x;
    const uint32_t owners = gridDim.x * kCountThreads / 4;
    const uint32_t quarter = tid & 3u;
    uint32_t i = tid >> 2;
    if (i >= n) return;
    const uint64_t *w0 = packed + lay.w0(), *seeds = packed + lay.seed(), *wx = packed + lay.wx();
    const uint32_t *qidx = reinterpret_cast<const uint32_t *>(packed + lay.qidx());
    const uint32_t ts = ix.table_s;

    P l = 0, h = 0;
    uint64_t word = 0, pend = 0, next_word = 0, next_lo = 0;
    [[maybe_unused]] uint64_t next_hi = 0;
    uint32_t q = 0, next_q = 0;
    uint32_t rem = 0;   // symbols still to consume (even)
    int shift = 60;     // bit offset of the next pair (4 bits) in `word`
    uint32_t widx = 0;

    auto prefetch = [&](uint32_t ii) {
        next_word = ldg_stream(w0 + ii, stream);
        next_lo = ldg_stream(seeds + ii, stream);
        if constexpr (WIDE) next_hi = ldg_stream(seeds + lay.n + ii, stream);
        next_q = __ldg(qidx + ii);
    };
    auto begin = [&]() {
        word = next_word;
        q = next_q & kQidxMask;
        if constexpr (WIDE) { l = next_lo; h = next_hi; } else { l = (uint32_t)next_lo; h = (uint32_t)(next_lo >> 32); }
        rem = k - table_depth(next_q >> 30, ts);
        shift = 60;
        widx = 0;
        if (rem > (uint32_t)kPairSymsPerWord) pend = ldg_stream(wx + q, stream);  // needed 16 pair steps from now
    };

    prefetch(i);
    begin();
    if (i + owners < n) prefetch(i + owners);

    for (;;) {
        while (rem == 0 || l == h) {
            if (quarter == 0) stg_stream(out + q, (uint64_t)(h - l), stream);
            i += owners;
            if (i >= n) return;
            begin();
            if (i + owners < n) prefetch(i + owners);
        }
        if (shift < 0) {
            word = pend;
            widx++;
            shift = 60;
            if (rem > (uint32_t)kPairSymsPerWord) pend = ldg_stream(wx + (uint64_t)widx * lay.n + q, stream);
        }
        const uint32_t code = (uint32_t)(word >> shift) & 15u;
        pair_step<WIDE>(ix, c2, code, l, h, quarter);
        rem -= 2;
        shift -= 4;
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
    cudaError_t e = cudaMemsetAsync(d_packed + lay.live(), 0, 2 * sizeof(uint64_t), st);
    if (e != cudaSuccess) return e;
    const unsigned blocks = (unsigned)((n + 255) / 256);
    const size_t smem = k <= kPackSmemMaxK ? 256u * (size_t)k + 16u : 0u;
    if (is_wide(ix)) pack_seed_kernel<true><<<blocks, 256, smem, st>>>(ix, d_syms, k, lay, d_packed, d_out, d_status);
    else pack_seed_kernel<false><<<blocks, 256, smem, st>>>(ix, d_syms, k, lay, d_packed, d_out, d_status);
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

template <bool WIDE>
static cudaError_t launch_count_pair_t(int device, const IndexView &ix, const uint64_t *d_packed,
                                       const PackedLayout &lay, uint32_t k, uint64_t *d_out, cudaStream_t st) {
    const unsigned grid = persistent_grid(device, (const void *)count_kmers_pair_kernel<WIDE>, kCountThreads, lay.n,
                                          kCountThreads / 4);
    count_kmers_pair_kernel<WIDE><<<grid, kCountThreads, 0, st>>>(ix, d_packed, lay, k, d_out);
    return cudaGetLastError();
}

// n <= kMaxPerLaunch (the callers chunk): query indices are u32 inside the kernels
cudaError_t launch_count_packed(int device, const IndexView &ix, int lanes, const uint64_t *d_packed, uint32_t k,
                                uint64_t n, uint64_t *d_out, cudaStream_t st, int *launches) {
    if (!n) return cudaSuccess;
    if (n > kMaxPerLaunch) return cudaErrorInvalidValue;
    const PackedLayout lay = packed_layout(ix, k, n);
    cudaError_t e;
    if (ix.pair) {  // list A: the pair image
        e = is_wide(ix) ? launch_count_pair_t<true>(device, ix, d_packed, lay, k, d_out, st)
                        : launch_count_pair_t<false>(device, ix, d_packed, lay, k, d_out, st);
        if (launches) (*launches)++;
        if (e != cudaSuccess) return e;
    }
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
