// quad_kernels.cu -- the search kernel over live list A for an index with a quad (and oct) image.
// Same structure as the pair / one-step kernels of kernels.cu (persistent grid, every thread walks its
// own query stream and refills itself); kept in its own translation unit because ptxas 12.9 crashes on
// the module that holds every search kernel together.
//
// Replaces the reference's hot loop: BWT::count_kmer (src/msbwt_core.rs:125-161) calling
// RleBWT::constrain_range (src/rle_bwt.rs:202-287) once per symbol -- here four or eight symbols per
// index access (layout.h states the identities).
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
// 32-bit positions only.  While eight or more symbols are left a step reads one 128-byte line of the oct
// image (layout.h) instead of two quad sectors; the quad image serves remainders of 4..7 symbols, ranges
// that straddle two oct buckets and the lines that overflowed (two quad steps instead of one oct step).
//
// Mapping: a QUAD OF LANES per query, each lane loading one 32-byte sector of the line, so that the line
// is ONE 128-byte request to L2.  What HBM random access is bound by is the number of requests that miss
// (about 40 G/s whatever their size, profiles/r1_gather_*.json): a thread that reads its line with four
// 256-bit loads pays four of them per line and ran at a quarter of the line rate
// (profiles/r1_o2_oct_cfg3_ncu_summary.txt).
//
// Every iteration of the persistent loop is split into ISSUE and CONSUME.  ISSUE is branch-free: whatever
// the query's next step is (oct line, quad sectors, nothing) it is the same predicated 256-bit load, so a
// warp whose eight queries need different kinds of step still has all its index requests in flight at once
// and pays one memory round trip per iteration; CONSUME diverges by kind, without memory accesses.
#ifndef MSBWT_OCT_CTAS
#define MSBWT_OCT_CTAS 6
#endif

// Query staging: every quad of lanes keeps the next kOctGroup queries of its slice (symbol word, seed
// range, original index) in its own 336 bytes of shared memory, filled with cp.async one group ahead: three
// full-line requests per 16 queries.  (A quad that read its next query with three 8-byte loads paid three
// L2 misses per query once the quads of a warp had drifted apart -- more than the two index lines the
// query itself needs: 42.6 GB of DRAM reads per 100 M queries against 25.6 GB of index lines,
// profiles/r1_o5_oct_cfg3_ncu_summary.txt.)
constexpr int kOctGroup = 16;
constexpr int kOctStageWords = 84;  // 16 u64 + 16 u64 + 16 u32 = 80 words, padded: 16-byte multiple, bank-conflict-free

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem, bool on) {
    const uint32_t dst = (uint32_t)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(gmem), "r"(on ? 16u : 0u) : "memory");
}
__device__ __forceinline__ void cp_async8(void *smem, const void *gmem, bool on) {
    const uint32_t dst = (uint32_t)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(dst), "l"(gmem), "r"(on ? 8u : 0u) : "memory");
}

// 32 bytes at p into v when `on` (v is left undefined otherwise); same cache policy as ldg_index256
__device__ __forceinline__ void ldg_index256_if(Half &v, const void *p, uint32_t on) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %9, 0;\n\t"
        "@p ld.global.nc.L1::no_allocate.L2::evict_last.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];\n\t}"
        : "=r"(v.w[0]), "=r"(v.w[1]), "=r"(v.w[2]), "=r"(v.w[3]), "=r"(v.w[4]), "=r"(v.w[5]), "=r"(v.w[6]), "=r"(v.w[7])
        : "l"(p), "r"(on));
}

// the rare one-symbol remainder step, out of line and by value so that neither its 32 load registers nor a
// stack slot for l / h burden the main loop
__device__ __noinline__ uint2 oct_remainder_step(const IndexView &ix, const uint32_t *cbase, uint32_t sym, uint32_t l, uint32_t h) {
    const CBase<false> cb{cbase};
    rank_step<false, 1>(ix, cb, sym, l, h);
    return make_uint2(l, h);
}

// TAIL: the batch ends with 1..3 one-symbol steps (k below the kept table levels); the out-of-line call is
// compiled only into that instantiation, so the common one keeps all its state in registers.
template <bool TAIL>
__global__ void __launch_bounds__(kCountThreads, MSBWT_OCT_CTAS)
count_kmers_oct_kernel(IndexView ix, const uint64_t *__restrict__ packed, PackedLayout lay, uint32_t k,
                       uint64_t *__restrict__ out) {
    __shared__ uint64_t cb_smem[4];
    __shared__ __align__(16) uint32_t stage[(kCountThreads / 4) * kOctStageWords];
    [[maybe_unused]] CBase<false> cb{nullptr};
    if constexpr (TAIL) cb = stage_cbase<false>(ix, cb_smem);
    const uint64_t stream = policy_evict_first();

    const uint32_t n = (uint32_t)packed[lay.live()];  // live queries of list A
    const uint32_t tid = blockIdx.x * kCountThreads + threadIdx.x;
    const uint32_t t = tid & 3u;                         // this lane's sector of the line
    const uint32_t qmask = 0xFu << (threadIdx.x & 28u);  // the lanes of this query
    // every quad of lanes owns one CONTIGUOUS slice of the live list (a multiple of kOctGroup queries)
    const uint32_t nquads = gridDim.x * (kCountThreads / 4);
    const uint32_t per = ((n + nquads - 1) / nquads + (uint32_t)kOctGroup - 1u) & ~((uint32_t)kOctGroup - 1u);
    const uint64_t start64 = (uint64_t)(tid >> 2) * per;
    if (start64 >= n) return;
    uint32_t i = (uint32_t)start64;
    const uint32_t end = (uint32_t)(start64 + per < n ? start64 + per : n);
    const uint64_t *w0 = packed + lay.w0(), *seeds = packed + lay.seed(), *wx = packed + lay.wx();
    const uint32_t *qidx = reinterpret_cast<const uint32_t *>(packed + lay.qidx());
    const uint32_t rem0 = k - acgt_table_depth(k, ix.table_s, 4u);
    const uint32_t bshift = ix.oct_shift, bmask = (1u << bshift) - 1u;
    const char *const oct_base = reinterpret_cast<const char *>(ix.oct) + 32u * t;
    const char *const quad_base = reinterpret_cast<const char *>(ix.quad);
    uint32_t *const my = stage + (threadIdx.x >> 2) * kOctStageWords;  // w0[16] | seed[16] | qidx[16]

    uint32_t l = 0, h = 0;
    uint64_t word = 0, pend = 0;
    uint32_t q = 0;
    uint32_t rem = 0;     // symbols still to consume
    int shift = 62;       // bit offset of the next symbol (2 bits) in `word`
    uint32_t widx = 0;
    uint32_t forced = 0;  // quad steps to take before the next oct step (after an overflowed line)

    // the next kOctGroup queries of the slice, global -> this quad's staging area (i0: a multiple of kOctGroup)
    auto fetch_group = [&](uint32_t i0) {
#pragma unroll
        for (uint32_t c = t; c < 8u; c += 4u) cp_async16(my + 4u * c, w0 + i0 + 2u * c, i0 + 2u * c < end);
#pragma unroll
        for (uint32_t c = t; c < 16u; c += 4u) cp_async8(my + 32u + 2u * c, seeds + i0 + c, i0 + c < end);
        cp_async16(my + 64u + 4u * t, qidx + i0 + 4u * t, i0 + 4u * t < end);
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    auto begin = [&]() {
        const uint32_t j = i & ((uint32_t)kOctGroup - 1u);
        if (j == 0) {
            asm volatile("cp.async.wait_all;" ::: "memory");
            __syncwarp(qmask);
        }
        word = *reinterpret_cast<const volatile uint64_t *>(my + 2u * j);
        const uint64_t lo = *reinterpret_cast<const volatile uint64_t *>(my + 32u + 2u * j);
        q = *reinterpret_cast<const volatile uint32_t *>(my + 64u + j) & kQidxMask;
        l = (uint32_t)lo;
        h = (uint32_t)(lo >> 32);
        rem = rem0;
        shift = 62;
        widx = 0;
        forced = 0;
        if (rem > (uint32_t)kPairSymsPerWord) pend = ldg_stream(wx + q, stream);
        if (j == (uint32_t)kOctGroup - 1u && i + 1u < end) {  // the group is used up: stage the next one under this query's steps
            __syncwarp(qmask);
            fetch_group(i + 1u);
        }
    };

    fetch_group(i);
    begin();

    for (;;) {
        while (rem == 0 || l == h) {
            if (t == 0) stg_stream(out + q, (uint64_t)(h - l), stream);
            if (++i >= end) return;
            begin();
        }
        if (shift < 0) {  // 32 symbols per word; steps of 8 and 4 symbols never straddle two words
            word = pend;
            widx++;
            shift = 62;
            if (rem > (uint32_t)kPairSymsPerWord) pend = ldg_stream(wx + (uint64_t)widx * lay.n + q, stream);
        }

        // ---- ISSUE (branch-free): lane t reads sector t of the oct line; lanes 0 / 1 the quad sectors of l / h
        // (a range over two buckets takes the eight symbols as two quad steps as well: oct steps stay aligned
        // to multiples of eight symbols and never straddle two words)
        const uint32_t bl = l >> bshift, bh = h >> bshift;
        const bool want_oct = rem >= 8u && forced == 0u;
        const bool is_oct = want_oct && bl == bh;
        const bool is_quad = !is_oct && rem >= 4u;
        if (want_oct && !is_oct) forced = 2;
        const uint32_t code16 = (uint32_t)(word >> (shift >= 14 ? shift - 14 : 0)) & 0xFFFFu;
        const uint32_t code8 = (uint32_t)(word >> (shift - 6)) & 255u;
        const uint32_t mine = t == 0 ? l : h;  // the boundary this lane ranks in a quad step
        const uint32_t sec = mine / (uint32_t)kQuadSyms;
        const char *p = is_oct ? oct_base + ((size_t)code16 * ix.nbuck8 + bl) * kOctLineBytes
                               : quad_base + ((size_t)code8 * ix.nsec4 + sec) * kQuadSectorBytes;
        Half v;
        ldg_index256_if(v, p, is_oct || (is_quad && t < 2u));

        // ---- CONSUME
        uint32_t x = 0;  // oct: this sector's occurrences below l (low half) and below h (high half); quad: the new boundary
        if (is_oct) {
            const int pl = (int)(l & bmask), ph = (int)(h & bmask);
            int sl = 0, sh = 0;
#pragma unroll
            for (int w = 0; w < 8; w++) {
                const uint32_t e = (w < 2 && t == 0) ? 0u : v.w[w];  // words 0, 1 of the line are not runs
                const int off = (int)(e & bmask), len = (int)(e >> bshift);
                sl += min(max(pl - off, 0), len);
                sh += min(max(ph - off, 0), len);
            }
            x = (uint32_t)sl | ((uint32_t)sh << 16);  // a line holds <= 30 runs of <= 1024 positions
        } else if (is_quad) {
            x = v.w[0] + sector_count_below(v, (int)(mine - sec * (uint32_t)kQuadSyms));
        }
        const uint32_t x0 = __shfl_sync(qmask, x, 0, 4), x1 = __shfl_sync(qmask, x, 1, 4);
        const uint32_t ckpt = __shfl_sync(qmask, v.w[0], 0, 4), nruns = __shfl_sync(qmask, v.w[1], 0, 4);
        uint32_t sum = x + __shfl_xor_sync(qmask, x, 1, 4);
        sum += __shfl_xor_sync(qmask, sum, 2, 4);
        if (is_oct) {
            if (nruns > (uint32_t)kOctCapacity) {
                forced = 2;  // this line cannot hold its runs: the same eight symbols as two quad steps
            } else {
                l = ckpt + (sum & 0xFFFFu);
                h = ckpt + (sum >> 16);
                rem -= 8;
                shift -= 16;
            }
        } else if (is_quad) {
            l = x0;
            h = x1;
            rem -= 4;
            shift -= 8;
            forced = forced ? forced - 1u : 0u;
        } else if constexpr (TAIL) {
            const uint32_t sym = (0x5321u >> (4u * ((uint32_t)(word >> shift) & 3u))) & 7u;  // A,C,G,T = 1,2,3,5
            const uint2 r = oct_remainder_step(ix, cb.c, sym, l, h);
            l = r.x;
            h = r.y;
            rem--;
            shift -= 2;
        } else {
            rem = 0;  // unreachable: the launcher picks TAIL whenever the remainder is not a multiple of four
        }
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
        if ((k - acgt_table_depth(k, ix.table_s, 4u)) % 4u) {
            const unsigned grid = persistent_grid(device, (const void *)count_kmers_oct_kernel<true>, kCountThreads, lay.n, kCountThreads / 4);
            count_kmers_oct_kernel<true><<<grid, kCountThreads, 0, st>>>(ix, d_packed, lay, k, d_out);
        } else {
            const unsigned grid = persistent_grid(device, (const void *)count_kmers_oct_kernel<false>, kCountThreads, lay.n, kCountThreads / 4);
            count_kmers_oct_kernel<false><<<grid, kCountThreads, 0, st>>>(ix, d_packed, lay, k, d_out);
        }
        return cudaGetLastError();
    }
    return launch_count_quad_t<false>(device, ix, d_packed, lay, k, d_out, st);
}

}  // namespace msbwt
