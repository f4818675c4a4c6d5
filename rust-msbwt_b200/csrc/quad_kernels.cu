// quad_kernels.cu -- the search kernel over live list A for an index with a quad (and oct) image.
// Same structure as the pair / one-step kernels of kernels.cu (persistent grid, every thread walks its
// own query stream and refills itself); kept in its own translation unit because ptxas 12.9 crashes on
// the module that holds every search kernel together.
//
// Replaces the reference's hot loop: BWT::count_kmer (src/msbwt_core.rs:125-161) calling
// RleBWT::constrain_range (src/rle_bwt.rs:202-287) once per symbol -- here four or eight symbols per
// index access, ten with the oct image (layout.h states the identities).
#include "device_rank.cuh"
#include "engine.h"
#include "kernel_common.cuh"

namespace msbwt {

// Persistent kernel over live list A with a QUAD image: one thread per query, four symbols per step
// (one 32-byte sector per boundary).  The depth the suffix table answered is a function of k alone for
// an all-ACGT k-mer (acgt_table_depth), so the remaining count is uniform across the list; a remainder
// that is not a multiple of four (k below the kept table levels) ends with one-step ranks.
constexpr int quad_min_ctas(bool wide) { return wide ? 4 : 6; }

// the rare remainder step of the quad kernel, kept out of line so that its 32 load registers do not
// set the register budget of the quad loop
// (one copy per kernel instantiation: ptxas 12.9 crashes when two entries with different register budgets
// share one out-of-line function)
template <bool WIDE>
__device__ __noinline__ void remainder_step(const IndexView &ix, const CBase<WIDE> &cb, uint32_t sym,
                                            typename Pos<WIDE>::type &l, typename Pos<WIDE>::type &h) {
    rank_step<WIDE, 1>(ix, cb, sym, l, h);
}

template <bool WIDE>
__global__ void __launch_bounds__(kCountThreads, quad_min_ctas(WIDE))
count_kmers_quad_kernel(IndexView ix, const uint64_t *__restrict__ packed, PackedLayout lay, uint32_t k,
                        uint64_t *__restrict__ out) {
    using P = typename Pos<WIDE>::type;
    __shared__ uint64_t c4_smem[WIDE ? kQuadMaxSuperInSmem * kQuadCodes : 1];
    __shared__ uint64_t cb_smem[WIDE ? kMaxSuperInSmem * 8 : 4];
    const C4Base<WIDE> c4 = stage_c4base<WIDE>(ix, c4_smem);
    const CBase<WIDE> cb = stage_cbase<WIDE>(ix, cb_smem);
    const uint64_t stream = policy_evict_first();

    const uint32_t n = (uint32_t)packed[lay.live()];  // live queries of list A
    const uint32_t owners = gridDim.x * kCountThreads;
    uint32_t i = blockIdx.x * kCountThreads + threadIdx.x;
    if (i >= n) return;
    const uint64_t *w0 = packed + lay.w0(), *seeds = packed + lay.seed(), *wx = packed + lay.wx();
    const uint32_t *qidx = reinterpret_cast<const uint32_t *>(packed + lay.qidx());
    const uint32_t rem0 = k - acgt_table_depth(k, ix.table_s, 4u);

    P l = 0, h = 0;
    uint64_t word = 0, pend = 0, next_word = 0, next_lo = 0;
    [[maybe_unused]] uint64_t next_hi = 0;
    uint32_t q = 0, next_q = 0;
    uint32_t rem = 0;   // symbols still to consume
    int shift = 62;     // bit offset of the next symbol (2 bits) in `word`
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
        rem = rem0;
        shift = 62;
        widx = 0;
        if (rem > (uint32_t)kPairSymsPerWord) pend = ldg_stream(wx + q, stream);  // needed 8 quad steps from now
    };

    prefetch(i);
    begin();
    if (i + owners < n) prefetch(i + owners);

    for (;;) {
        while (rem == 0 || l == h) {
            stg_stream(out + q, (uint64_t)(h - l), stream);
            i += owners;
            if (i >= n) return;
            begin();
            if (i + owners < n) prefetch(i + owners);
        }
        if (shift < 0) {  // 32 symbols per word: a quad never straddles two words
            word = pend;
            widx++;
            shift = 62;
            if (rem > (uint32_t)kPairSymsPerWord) pend = ldg_stream(wx + (uint64_t)widx * lay.n + q, stream);
        }
        if (rem >= 4u) {
            const uint32_t code = (uint32_t)(word >> (shift - 6)) & 255u;
            quad_step<WIDE>(ix, c4, code, l, h);
            rem -= 4;
            shift -= 8;
        } else {
            const uint32_t sym = (0x5321u >> (4u * ((uint32_t)(word >> shift) & 3u))) & 7u;  // A,C,G,T = 1,2,3,5
            remainder_step<WIDE>(ix, cb, sym, l, h);
            rem--;
            shift -= 2;
        }
    }
}

// ---------------------------------------------------------------- oct image on top of the quad image
//
// 32-bit positions only.  While kOctSyms (ten) or more symbols are left a step reads one 128-byte line of the
// oct image (layout.h); the quad image (and one-symbol ranks) serve remainders, ranges that straddle two oct
// buckets and the lines that overflowed (two quad steps + two one-symbol steps instead of one oct step).
//
// What HBM random access is bound by is the number of L2 requests that miss -- about 40 G/s whatever their
// size (profiles/r1_gather_*.json) -- and what reaches that bound is the number of them in flight.  So:
//
//  * one thread per query (1024 queries in flight per SM), but the index lines are fetched by the WARP:
//    every lane publishes the address of the line (or the two quad sectors) its query needs next, and in
//    eight rounds the warp copies the 32 lines into shared memory with cp.async, eight lanes x 16 bytes per
//    line, i.e. ONE 128-byte request per line and no load registers.  (A thread that read its own line with
//    four 256-bit loads paid four requests per line and ran at a quarter of the line rate; a quad of lanes
//    per query kept only 384 queries per SM in flight: profiles/r1_o2_*, r1_o6_* summaries.)
//  * every iteration is ISSUE (branch-free, whatever kind of step each lane needs), one wait, CONSUME (each
//    lane ranks in its own staged line; divergent, but without memory accesses): one memory round trip per
//    iteration even when the lanes of a warp need different kinds of step -- one lane in fourteen lands on
//    an overflowed line on 30x reads with 1 % errors.
//  * every warp takes chunks of 512 consecutive queries of the live list from an atomic counter and stages
//    them through shared memory 32 queries at a time, double-buffered (coalesced cp.async one pool ahead);
//    lanes that finished take the next queries of the pool in lane order, so the warp stays full whatever the
//    mix of early exits and no warp is left with a slow slice of a skewed batch.
#ifndef MSBWT_OCT_CTAS
#define MSBWT_OCT_CTAS 4
#endif
constexpr int kOctRowBytes = 144;    // a 128-byte line + 16: rows of consecutive lanes start 4 banks apart (conflict-free LDS.128)
constexpr int kOctPoolBytes = 640;   // 32 x (u64 symbol word, u64 seed range, u32 original index)
constexpr int kOctWarpSmem = 32 * kOctRowBytes + 2 * kOctPoolBytes;  // 5888 bytes per warp, 47104 per CTA
constexpr int kOctChunk = 512;       // queries a warp takes from the live list at a time

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem, bool on) {
    const uint32_t dst = (uint32_t)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(gmem), "r"(on ? 16u : 0u) : "memory");
}
__device__ __forceinline__ void cp_async8(void *smem, const void *gmem, bool on) {
    const uint32_t dst = (uint32_t)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(dst), "l"(gmem), "r"(on ? 8u : 0u) : "memory");
}
__device__ __forceinline__ void cp_async4(void *smem, const void *gmem, bool on) {
    const uint32_t dst = (uint32_t)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst), "l"(gmem), "r"(on ? 4u : 0u) : "memory");
}

// the rare one-symbol remainder step, out of line and by value so that neither its 32 load registers nor a
// stack slot for l / h burden the main loop
__device__ __noinline__ uint2 oct_remainder_step(const IndexView &ix, const uint32_t *cbase, uint32_t sym, uint32_t l, uint32_t h) {
    const CBase<false> cb{cbase};
    rank_step<false, 1>(ix, cb, sym, l, h);
    return make_uint2(l, h);
}

// occurrences below bucket offsets pl / ph contributed by one stored run `(len << b) | off` (0 = empty slot)
__device__ __forceinline__ void oct_add_run(uint32_t e, uint32_t b, uint32_t mask, int pl, int ph, int &sl, int &sh) {
    const int off = (int)(e & mask), len = (int)(e >> b);
    sl += min(max(pl - off, 0), len);
    sh += min(max(ph - off, 0), len);
}
__device__ __forceinline__ void oct_add_runs(const uint4 &v, uint32_t b, uint32_t mask, int pl, int ph, int &sl, int &sh) {
    oct_add_run(v.x, b, mask, pl, ph, sl, sh);
    oct_add_run(v.y, b, mask, pl, ph, sl, sh);
    oct_add_run(v.z, b, mask, pl, ph, sl, sh);
    oct_add_run(v.w, b, mask, pl, ph, sl, sh);
}
// a staged quad sector {checkpoint, 224 occurrence bits}: checkpoint + set bits at offsets < p
__device__ __forceinline__ uint32_t staged_sector_rank(const uint4 &a, const uint4 &b, int p) {
    return a.x + __popc(a.y & below_mask(p)) + __popc(a.z & below_mask(p - 32)) + __popc(a.w & below_mask(p - 64)) +
           __popc(b.x & below_mask(p - 96)) + __popc(b.y & below_mask(p - 128)) + __popc(b.z & below_mask(p - 160)) +
           __popc(b.w & below_mask(p - 192));
}

__global__ void __launch_bounds__(kCountThreads, MSBWT_OCT_CTAS)
count_kmers_oct_kernel(IndexView ix, const uint64_t *__restrict__ packed, PackedLayout lay, uint32_t k,
                       uint64_t *__restrict__ out, uint32_t *__restrict__ work) {
    __shared__ __align__(16) uint8_t smem[(kCountThreads / 32) * kOctWarpSmem];
    __shared__ uint64_t cb_smem[4];
    const CBase<false> cb = stage_cbase<false>(ix, cb_smem);
    const uint64_t stream = policy_evict_first();
    constexpr uint32_t kFull = 0xffffffffu;

    const uint32_t n = (uint32_t)packed[lay.live()];  // live queries of list A
    const uint32_t lane = threadIdx.x & 31u;
    if ((blockIdx.x * (kCountThreads / 32) + (threadIdx.x >> 5)) * 32u >= n) return;  // more warps than pools of work
    const uint64_t *w0 = packed + lay.w0(), *seeds = packed + lay.seed(), *wx = packed + lay.wx();
    const uint32_t *qidx = reinterpret_cast<const uint32_t *>(packed + lay.qidx());
    const uint32_t rem0 = k - list_a_table_depth(ix, k);
    const uint32_t bshift = ix.oct_shift, bmask = (1u << bshift) - 1u;
    const char *const oct_base = reinterpret_cast<const char *>(ix.oct);
    const char *const quad_base = reinterpret_cast<const char *>(ix.quad);
    uint8_t *const rows = smem + (threadIdx.x >> 5) * kOctWarpSmem;  // 32 rows of kOctRowBytes
    uint8_t *const pools = rows + 32 * kOctRowBytes;                  // 2 pools: w0[32] | seed[32] | qidx[32]
    const uint4 *const my_row = reinterpret_cast<const uint4 *>(rows + lane * kOctRowBytes);

    // The live list is handed out in chunks of kOctChunk queries (an atomic counter: a warp whose queries die
    // early simply comes back sooner, whatever the order of the batch) and staged pool by pool: pool A
    // (sequence number `seq`, buffer seq & 1) is being handed to the lanes, pool B (the other buffer) is
    // already staged or on its way.
    uint32_t chunk_next = 0, chunk_end = 0;  // the rest of this warp's current chunk (warp-uniform)
    auto next_pool = [&](uint32_t &base, uint32_t &cnt) {
        if (chunk_next >= chunk_end) {
            uint32_t c = 0;
            if (lane == 0) c = atomicAdd(work, (uint32_t)kOctChunk);
            c = __shfl_sync(kFull, c, 0);
            chunk_next = min(c, n);
            chunk_end = min(c + (uint32_t)kOctChunk, n);
        }
        base = chunk_next;
        cnt = min(32u, chunk_end - chunk_next);
        chunk_next += cnt;
    };
    auto load_pool = [&](uint32_t buf, uint32_t base, uint32_t cnt) {
        uint8_t *p = pools + buf * kOctPoolBytes;
        const uint32_t idx = base + lane;
        const bool on = lane < cnt;
        cp_async8(p + 8u * lane, w0 + idx, on);
        cp_async8(p + 256u + 8u * lane, seeds + idx, on);
        cp_async4(p + 512u + 4u * lane, qidx + idx, on);
    };
    uint32_t seq = 0, a_pos = 0, a_cnt, b_cnt, base;
    next_pool(base, a_cnt);
    load_pool(0u, base, a_cnt);
    next_pool(base, b_cnt);
    load_pool(1u, base, b_cnt);
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncwarp();

    bool active = false;
    uint32_t l = 0, h = 0;
    uint64_t word = 0, pend = 0;
    uint32_t q = 0;
    uint32_t rem = 0;     // symbols still to consume
    int shift = 62;       // bit offset of the next symbol (2 bits) in `word`; negative: already inside `pend`
    uint32_t widx = 0;
    uint32_t forced = 0;  // quad steps to take instead of the next oct step (overflowed line / two buckets)

    // the next `nsym` symbols as one code (first consumed most significant); a step may straddle two words
    auto peek = [&](uint32_t nsym) -> uint32_t {
        const int bits = 2 * (int)nsym, avail = shift + 2;  // avail >= 2 here
        if (avail >= bits) return (uint32_t)(word >> (avail - bits)) & ((1u << bits) - 1u);
        const int need = bits - avail;
        return (uint32_t)(((word & ((1ull << avail) - 1ull)) << need) | (pend >> (64 - need)));
    };

    for (;;) {
        // ---- RETIRE + REFILL (warp-uniform control)
        if (active && (rem == 0 || l == h)) {
            stg_stream(out + q, (uint64_t)(h - l), stream);
            active = false;
        }
        const uint32_t want = __ballot_sync(kFull, !active);
        if (want) {
            const uint32_t avail_a = a_cnt - a_pos, avail = avail_a + b_cnt;
            if (avail) {
                const uint32_t r = __popc(want & ((1u << lane) - 1u));  // idle lanes take queries in lane order
                if (!active && r < avail) {
                    const bool from_a = r < avail_a;
                    const uint8_t *p = pools + ((from_a ? seq : seq + 1u) & 1u) * kOctPoolBytes;
                    const uint32_t slot = from_a ? a_pos + r : r - avail_a;
                    word = *reinterpret_cast<const volatile uint64_t *>(p + 8u * slot);
                    const uint64_t lo = *reinterpret_cast<const volatile uint64_t *>(p + 256u + 8u * slot);
                    q = *reinterpret_cast<const volatile uint32_t *>(p + 512u + 4u * slot) & kQidxMask;
                    l = (uint32_t)lo;
                    h = (uint32_t)(lo >> 32);
                    rem = rem0;
                    shift = 62;
                    widx = 0;
                    forced = 0;
                    active = true;
                    if (rem > (uint32_t)kPairSymsPerWord) pend = ldg_stream(wx + q, stream);
                }
                const uint32_t taken = min((uint32_t)__popc(want), avail);
                if (taken >= avail_a) {  // pool A is used up: B becomes A, the next pool is staged into A's buffer
                    a_cnt = b_cnt;
                    a_pos = taken - avail_a;
                    seq++;
                    next_pool(base, b_cnt);
                    __syncwarp();
                    load_pool((seq + 1u) & 1u, base, b_cnt);  // (committed with this iteration's lines)
                } else {
                    a_pos += taken;
                }
            } else if (want == kFull) {
                return;  // nothing left to hand out and every lane is done
            }
        }
        if (active && shift < 0) {  // 32 symbols per word: on to the next one (a step may have ended inside it)
            word = pend;
            widx++;
            shift += 64;
            if (2u * rem > (uint32_t)(shift + 2)) pend = ldg_stream(wx + (uint64_t)widx * lay.n + q, stream);
        }

        // ---- ISSUE (branch-free): every lane publishes what its query needs, the warp fetches it
        // (a range over two buckets, like an overflowed line, takes its kOctSyms symbols as quad steps and, for
        // the last two of ten, one-symbol steps)
        const bool live = active && rem != 0 && l != h;
        const uint32_t bl = l >> bshift, bh = h >> bshift;
        const bool want_oct = live && rem >= (uint32_t)kOctSyms && forced == 0u;
        const bool is_oct = want_oct && bl == bh;
        if (want_oct && !is_oct) forced = (uint32_t)kOctSyms;  // symbols to take without the oct image
        const bool is_quad = live && !is_oct && rem >= 4u && (forced == 0u || forced >= 4u);
        const uint32_t codem = peek((uint32_t)kOctSyms);
        const uint32_t code8 = peek(4u);
        const uint32_t sl = l / (uint32_t)kQuadSyms, sh = h / (uint32_t)kQuadSyms;
        const char *p0 = is_oct ? oct_base + ((size_t)codem * ix.nbuck8 + bl) * kOctLineBytes
                                : quad_base + ((size_t)code8 * ix.nsec4 + sl) * kQuadSectorBytes;
        // low two bits: kind (1 oct, 2 quad, 0 nothing); the rest: byte distance from the sector of l to the sector of h
        const uint32_t meta = is_oct ? 1u : (is_quad ? (2u | ((sh - sl) * (uint32_t)kQuadSectorBytes)) : 0u);
        const uint32_t p0_lo = (uint32_t)(uintptr_t)p0, p0_hi = (uint32_t)((uintptr_t)p0 >> 32);
        {
            const uint32_t j = lane & 7u;  // this lane's 16 bytes of a line
#pragma unroll
            for (uint32_t c = 0; c < 8u; c++) {
                const uint32_t o = 4u * c + (lane >> 3);  // the lane whose line this is
                const uint32_t m = __shfl_sync(kFull, meta, o);
                const uint64_t a = ((uint64_t)__shfl_sync(kFull, p0_hi, o) << 32) | __shfl_sync(kFull, p0_lo, o);
                // oct: bytes 16j.. of the line; quad: the sector of l into bytes 0..31, the sector of h into 32..63
                const uint64_t src = a + 16u * j + (((m & 3u) == 2u && j >= 2u) ? (uint64_t)(m & ~31u) - 32u : 0u);
                cp_async16(rows + o * kOctRowBytes + 16u * j, reinterpret_cast<const void *>(src),
                           (m & 3u) == 1u || ((m & 3u) == 2u && j < 4u));
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_all;" ::: "memory");
        __syncwarp();

        // ---- CONSUME (each lane ranks in its own staged line)
        if (is_oct) {
            const uint4 a = my_row[0], b = my_row[1];
            if (a.y > (uint32_t)kOctCapacity) {
                forced = (uint32_t)kOctSyms;  // this line cannot hold its runs: the same symbols without the oct image
            } else {
                const int pl = (int)(l & bmask), ph = (int)(h & bmask);
                int cl = 0, ch = 0;
                oct_add_run(a.z, bshift, bmask, pl, ph, cl, ch);
                oct_add_run(a.w, bshift, bmask, pl, ph, cl, ch);
                oct_add_runs(b, bshift, bmask, pl, ph, cl, ch);
                if (a.y > 6u) {
                    oct_add_runs(my_row[2], bshift, bmask, pl, ph, cl, ch);
                    oct_add_runs(my_row[3], bshift, bmask, pl, ph, cl, ch);
                    if (a.y > 14u) {
                        oct_add_runs(my_row[4], bshift, bmask, pl, ph, cl, ch);
                        oct_add_runs(my_row[5], bshift, bmask, pl, ph, cl, ch);
                        if (a.y > 22u) {
                            oct_add_runs(my_row[6], bshift, bmask, pl, ph, cl, ch);
                            oct_add_runs(my_row[7], bshift, bmask, pl, ph, cl, ch);
                        }
                    }
                }
                l = a.x + (uint32_t)cl;
                h = a.x + (uint32_t)ch;
                rem -= (uint32_t)kOctSyms;
                shift -= 2 * kOctSyms;
            }
        } else if (is_quad) {
            const uint32_t nl = staged_sector_rank(my_row[0], my_row[1], (int)(l - sl * (uint32_t)kQuadSyms));
            const uint32_t nh = staged_sector_rank(my_row[2], my_row[3], (int)(h - sh * (uint32_t)kQuadSyms));
            l = nl;
            h = nh;
            rem -= 4;
            shift -= 8;
            forced = forced >= 4u ? forced - 4u : 0u;
        } else if (live) {  // one symbol: the tail of a k-mer, or the last two of ten symbols taken without the oct image
            const uint32_t sym = (0x5321u >> (4u * peek(1u))) & 7u;  // A,C,G,T = 1,2,3,5
            const uint2 r = oct_remainder_step(ix, cb.c, sym, l, h);
            l = r.x;
            h = r.y;
            rem--;
            shift -= 2;
            forced = forced ? forced - 1u : 0u;
        }
        __syncwarp();  // the rows are rewritten by the next ISSUE
    }
}

template <bool WIDE>
static cudaError_t launch_count_quad_t(int device, const IndexView &ix, const uint64_t *d_packed,
                                       const PackedLayout &lay, uint32_t k, uint64_t *d_out, cudaStream_t st) {
    const unsigned grid = persistent_grid(device, (const void *)count_kmers_quad_kernel<WIDE>, kCountThreads, lay.n,
                                          kCountThreads);
    count_kmers_quad_kernel<WIDE><<<grid, kCountThreads, 0, st>>>(ix, d_packed, lay, k, d_out);
    return cudaGetLastError();
}

cudaError_t launch_count_quad(int device, const IndexView &ix, const uint64_t *d_packed, const PackedLayout &lay,
                              uint32_t k, uint64_t *d_out, cudaStream_t st) {
    if (index_is_wide(ix)) return launch_count_quad_t<true>(device, ix, d_packed, lay, k, d_out, st);
    if (ix.oct) {
        static bool carveout_done[64] = {};  // per device: 4 CTAs x 47 KB of staging per SM need the large shared-memory configuration
        bool dummy = false;
        bool &carveout_set = (device >= 0 && device < 64) ? carveout_done[device] : dummy;
        if (!carveout_set) {
            cudaFuncSetAttribute((const void *)count_kmers_oct_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
            carveout_set = true;
        }
        // the chunk dispenser lives in the scratch buffer next to the live counters (engine bookkeeping: the
        // buffer is the engine's own scratch, `const` only towards the caller's data in it)
        uint32_t *work = reinterpret_cast<uint32_t *>(const_cast<uint64_t *>(d_packed) + lay.work());
        if (cudaError_t e = cudaMemsetAsync(work, 0, sizeof(uint32_t), st); e != cudaSuccess) return e;
        const unsigned grid = persistent_grid(device, (const void *)count_kmers_oct_kernel, kCountThreads, lay.n, kCountThreads);
        count_kmers_oct_kernel<<<grid, kCountThreads, 0, st>>>(ix, d_packed, lay, k, d_out, work);
        return cudaGetLastError();
    }
    return launch_count_quad_t<false>(device, ix, d_packed, lay, k, d_out, st);
}

}  // namespace msbwt
