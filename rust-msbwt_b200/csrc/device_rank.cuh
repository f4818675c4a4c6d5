// device_rank.cuh -- device-side rank primitives over the block images of layout.h, shared by the
// query kernels (kernels.cu) and the pair-image builder (pair_builder.cu).
//
// rank_step  = one RleBWT::constrain_range (src/rle_bwt.rs:202-287) over the 64-byte one-step blocks;
// pair_step  = two of them at once over the 128-byte pair lines.
#pragma once
#include <cstdint>

#include "engine.h"

namespace msbwt {

// ---------------------------------------------------------------- device helpers

__device__ __forceinline__ uint64_t policy_evict_first() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}

__device__ __forceinline__ uint64_t policy_evict_last() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
// a suffix-table entry {l, h}: the table level a batch reads is small (33 MB at depth 11) and every entry is read many
// times per batch, while the index lines around it are read once -- keep it in L2 against that stream
__device__ __forceinline__ uint2 ldg_table_entry(const uint2 *p, uint64_t pol_last) {
    uint2 r;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v2.u32 {%0,%1}, [%2], %3;" : "=r"(r.x), "=r"(r.y) : "l"(p), "l"(pol_last));
    return r;
}

__device__ __forceinline__ ulonglong2 ldg_table_entry(const ulonglong2 *p, uint64_t pol_last) {  // 64-bit positions
    ulonglong2 r;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v2.u64 {%0,%1}, [%2], %3;" : "=l"(r.x), "=l"(r.y) : "l"(p), "l"(pol_last));
    return r;
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

// The symbol stored at BWT position p < N (one 4-byte read per bit-plane of the half that holds it, layout.h), and
// LF(p) = C[B[p]] + rank(B[p], p): what the image builders that walk the one-step blocks use (oct_builder.cu,
// fin_builder.cu: the indexes too large for a quad image to walk LF^4 with).
__device__ __forceinline__ uint32_t symbol_at(const IndexView &ix, uint64_t p) {
    const uint32_t *w = reinterpret_cast<const uint32_t *>(ix.blocks) + (p >> kBlockShift) * (uint64_t)kWordsPerBlock +
                        (((uint32_t)p >> 6) & 1u) * 8u + (((uint32_t)p >> 5) & 1u);
    const uint32_t b = (uint32_t)p & 31u;
    return ((__ldg(w + 2) >> b) & 1u) | (((__ldg(w + 4) >> b) & 1u) << 1) | (((__ldg(w + 6) >> b) & 1u) << 2);
}
template <bool WIDE>
__device__ __forceinline__ typename Pos<WIDE>::type lf_step(const IndexView &ix, const CBase<WIDE> &cb, uint32_t sym,
                                                            typename Pos<WIDE>::type p) {
    typename Pos<WIDE>::type l = p, h = p;
    rank_step<WIDE, 1>(ix, cb, sym, l, h);
    return l;
}

// The four constrain_range calls of a backward-search extension at once: the block(s) holding
// l and h are fetched ONCE and ranked for A, C, G and T.  out_l[j], out_h[j] = constrain_range(ACGT[j], [l,h)).
template <bool WIDE>
__device__ __forceinline__ void rank_fanout4(const IndexView &ix, const CBase<WIDE> &cb, typename Pos<WIDE>::type l,
                                             typename Pos<WIDE>::type h, typename Pos<WIDE>::type (&out_l)[4],
                                             typename Pos<WIDE>::type (&out_h)[4]) {
    using P = typename Pos<WIDE>::type;
    const P bl = l >> kBlockShift, bh = h >> kBlockShift;
    const bool two = bh != bl;
    const char *base = reinterpret_cast<const char *>(ix.blocks);
    const int pl = (int)((uint32_t)l & (kBlockSyms - 1)), ph = (int)((uint32_t)h & (kBlockSyms - 1));
    const Half l0 = ldg_index256(base + (size_t)bl * kBlockBytes);
    const Half l1 = ldg_index256(base + (size_t)bl * kBlockBytes + 32);
    Half h0 = l0, h1 = l1;
    if (two) {
        h0 = ldg_index256(base + (size_t)bh * kBlockBytes);
        h1 = ldg_index256(base + (size_t)bh * kBlockBytes + 32);
    }
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const uint32_t sym = (0x5321u >> (4 * j)) & 7u;  // A,C,G,T = 1,2,3,5; their checkpoint slot is j
        const uint32_t x0 = (sym & 1u) - 1u, x1 = ((sym >> 1) & 1u) - 1u, x2 = ((sym >> 2) & 1u) - 1u;
        uint32_t ml[4], mh[4];
        match_half(l0, x0, x1, x2, ml[0], ml[1]);
        match_half(l1, x0, x1, x2, ml[2], ml[3]);
        match_half(h0, x0, x1, x2, mh[0], mh[1]);
        match_half(h1, x0, x1, x2, mh[2], mh[3]);
        const uint32_t ckl = (j & 2) ? l1.w[j & 1] : l0.w[j & 1], ckh = (j & 2) ? h1.w[j & 1] : h0.w[j & 1];
        out_l[j] = cb.at(bl, sym) + ckl + count_below64(ml[0], ml[1], pl) + count_below64(ml[2], ml[3], pl - 64);
        out_h[j] = cb.at(bh, sym) + ckh + count_below64(mh[0], mh[1], ph) + count_below64(mh[2], mh[3], ph - 64);
    }
}

// ---- pair image (layout.h): two constrain_range steps per 128-byte line

template <bool WIDE> struct C2Base;
template <> struct C2Base<false> {  // N < 2^32: the line's checkpoints are absolute
    __device__ __forceinline__ uint32_t at(uint32_t, uint32_t) const { return 0u; }
};
template <> struct C2Base<true> {
    const uint64_t *c;
    uint32_t sb_shift;
    __device__ __forceinline__ uint64_t at(uint64_t line, uint32_t code) const { return c[((line >> sb_shift) << 4) + code]; }
};

template <bool WIDE>
__device__ __forceinline__ C2Base<WIDE> stage_c2base(const IndexView &ix, uint64_t *smem) {
    if constexpr (WIDE) {
        if (ix.n_super2 > (uint32_t)kMaxSuperInSmem) return C2Base<true>{ix.c2base, ix.sb_shift};
        for (uint32_t i = threadIdx.x; i < ix.n_super2 * 16u; i += blockDim.x) smem[i] = ix.c2base[i];
        __syncthreads();
        return C2Base<true>{smem, ix.sb_shift};
    } else {
        return C2Base<false>{};
    }
}

// the 24 match bits of a quarter for pair `code`; bits 24..31 of the result are garbage and every
// mask applied to it stays below bit 24
__device__ __forceinline__ uint32_t match_quarter(const Half &v, uint32_t x0, uint32_t x1, uint32_t x2, uint32_t x3) {
    const uint32_t t = __byte_perm(v.w[4], v.w[5], 0x0073);  // valid plane: top bytes of words 4,5,6
    const uint32_t valid = __byte_perm(t, v.w[6], 0x0710);
    return valid & (v.w[4] ^ x0) & (v.w[5] ^ x1) & (v.w[6] ^ x2) & (v.w[7] ^ x3);
}

// Two constrain_range steps at once: code = 4*idx(b) + idx(a), b consumed first.
// [l,h) -> [C2[b,a] + rank2(code,l), C2[b,a] + rank2(code,h)).  Called by the four lanes of a quad
// together with identical (code, l, h); `quarter` = lane & 3: one 256-bit load per lane = ONE coalesced
// 128-byte request per line.  (Measured: splitting it into a 128-bit plane load plus a 32-bit load of
// the one checkpoint needed saves registers but doubles the kernel time -- profiles/README.md.)
// Full-mask shuffles: see rank_step.
template <bool WIDE>
__device__ __forceinline__ void pair_step(const IndexView &ix, const C2Base<WIDE> &c2, uint32_t code,
                                          typename Pos<WIDE>::type &l, typename Pos<WIDE>::type &h, uint32_t quarter) {
    using P = typename Pos<WIDE>::type;
    const P bl = l / (P)kPairSyms, bh = h / (P)kPairSyms;
    const int pl = (int)(uint32_t)(l - bl * (P)kPairSyms), ph = (int)(uint32_t)(h - bh * (P)kPairSyms);
    const bool two = bh != bl;
    const char *base = reinterpret_cast<const char *>(ix.pair) + quarter * 32u;
    const Half a = ldg_index256(base + (size_t)bl * kPairBytes);
    Half b;
    if (two) b = ldg_index256(base + (size_t)bh * kPairBytes);
    const uint32_t x0 = (code & 1u) - 1u, x1 = ((code >> 1) & 1u) - 1u, x2 = ((code >> 2) & 1u) - 1u,
                   x3 = ((code >> 3) & 1u) - 1u;  // 0 or ~0
    const uint32_t ml = match_quarter(a, x0, x1, x2, x3);
    const uint32_t lo = (code & 1u) ? a.w[1] : a.w[0], hi = (code & 1u) ? a.w[3] : a.w[2];
    uint32_t cand_l = (code & 2u) ? hi : lo, cand_h = cand_l, mh = ml;
    if (two) {
        mh = match_quarter(b, x0, x1, x2, x3);
        const uint32_t lo2 = (code & 1u) ? b.w[1] : b.w[0], hi2 = (code & 1u) ? b.w[3] : b.w[2];
        cand_h = (code & 2u) ? hi2 : lo2;
    }
    const int off = (int)quarter * kPairQuarterSyms;
    uint32_t cnt = __popc(ml & below_mask(min(pl - off, kPairQuarterSyms))) |
                   (__popc(mh & below_mask(min(ph - off, kPairQuarterSyms))) << 16);
    cnt += __shfl_xor_sync(0xffffffffu, cnt, 1);
    cnt += __shfl_xor_sync(0xffffffffu, cnt, 2);
    const uint32_t ckl = __shfl_sync(0xffffffffu, cand_l, code >> 2, 4);  // the quarter that owns this pair's checkpoint
    const uint32_t ckh = __shfl_sync(0xffffffffu, cand_h, code >> 2, 4);
    l = c2.at(bl, code) + ckl + (cnt & 0xffffu);
    h = c2.at(bh, code) + ckh + (cnt >> 16);
}

// ---- quad image (layout.h): four constrain_range steps per 32-byte sector

template <bool WIDE> struct C4Base;
template <> struct C4Base<false> {  // N < 2^32: the sector's checkpoint is absolute
    __device__ __forceinline__ uint32_t at(uint32_t, uint32_t) const { return 0u; }
};
template <> struct C4Base<true> {
    const uint64_t *c;
    uint32_t sb_shift;
    __device__ __forceinline__ uint64_t at(uint64_t sector, uint32_t code) const {
        return c[((sector >> sb_shift) << 8) + code];
    }
};

template <bool WIDE>
__device__ __forceinline__ C4Base<WIDE> stage_c4base(const IndexView &ix, uint64_t *smem) {
    if constexpr (WIDE) {
        if (ix.n_super4 > (uint32_t)kQuadMaxSuperInSmem) return C4Base<true>{ix.c4base, ix.sb_shift4};
        for (uint32_t i = threadIdx.x; i < ix.n_super4 * (uint32_t)kQuadCodes; i += blockDim.x) smem[i] = ix.c4base[i];
        __syncthreads();
        return C4Base<true>{smem, ix.sb_shift4};
    } else {
        return C4Base<false>{};
    }
}

// set bits of a sector's 224 occurrence bits at offsets < p (0 <= p < 224)
__device__ __forceinline__ uint32_t sector_count_below(const Half &v, int p) {
    uint32_t c = 0;
#pragma unroll
    for (int w = 0; w < 7; w++) c += __popc(v.w[1 + w] & below_mask(p - 32 * w));
    return c;
}

// Four constrain_range steps at once: code = 64*idx(b0) + 16*idx(b1) + 4*idx(b2) + idx(b3), b0 consumed
// first.  [l,h) -> [C4[code] + rank4(code,l), C4[code] + rank4(code,h)).  One thread, one 256-bit load
// per boundary (the second one only when h falls in another sector), both issued before either is used.
template <bool WIDE>
__device__ __forceinline__ void quad_step(const IndexView &ix, const C4Base<WIDE> &c4, uint32_t code,
                                          typename Pos<WIDE>::type &l, typename Pos<WIDE>::type &h) {
    using P = typename Pos<WIDE>::type;
    const P sl = l / (P)kQuadSyms, sh = h / (P)kQuadSyms;
    const int pl = (int)(uint32_t)(l - sl * (P)kQuadSyms), ph = (int)(uint32_t)(h - sh * (P)kQuadSyms);
    const char *base = reinterpret_cast<const char *>(ix.quad) + (size_t)code * ix.nsec4 * kQuadSectorBytes;
    const Half a = ldg_index256(base + (size_t)sl * kQuadSectorBytes);
    Half b = a;
    if (sh != sl) b = ldg_index256(base + (size_t)sh * kQuadSectorBytes);
    l = c4.at(sl, code) + a.w[0] + sector_count_below(a, pl);
    h = c4.at(sh, code) + b.w[0] + sector_count_below(b, ph);
}

}  // namespace msbwt
