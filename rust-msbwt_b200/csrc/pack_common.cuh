// pack_common.cuh -- what the pack / seed kernels share (kernels.cu: pack_seed_kernel, seed_packed_kernel,
// seed_u64_kernel; final_kernels.cu: pack_seed_final_kernel): the seed plan, the suffix-table lookup of an all-ACGT
// k-mer, the general path of k-mers holding `$` / `N`, and the CTA-wide append to the live lists.
#pragma once
#include "device_rank.cuh"
#include "engine.h"
#include "kernel_common.cuh"

namespace msbwt {

// ---------------------------------------------------------------- K0: pack + validate + seed

// CTA-wide append to the two live lists (A grows from slot 0 upwards, B from slot n-1 downwards): one
// atomicAdd per list per CTA -- a per-warp atomic on the same two counters serialises in L2.  Every
// thread of the (256-thread) CTA must call it; returns the caller's slot (meaningless unless live).
__device__ __forceinline__ uint64_t append_live(bool live, bool list_a, unsigned long long *counters, uint64_t n) {
    __shared__ uint32_t warp_cnt[2][8];
    __shared__ unsigned long long warp_base[2][8];
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const uint32_t mask_a = __ballot_sync(0xffffffffu, live && list_a);
    const uint32_t mask_b = __ballot_sync(0xffffffffu, live && !list_a);
    if (lane == 0) { warp_cnt[0][warp] = __popc(mask_a); warp_cnt[1][warp] = __popc(mask_b); }
    __syncthreads();
    if (threadIdx.x < 2) {
        const uint32_t which = threadIdx.x;
        uint32_t total = 0;
        for (uint32_t w = 0; w < 8; w++) total += warp_cnt[which][w];
        unsigned long long base = total ? atomicAdd(counters + which, (unsigned long long)total) : 0ull;
        for (uint32_t w = 0; w < 8; w++) { warp_base[which][w] = base; base += warp_cnt[which][w]; }
    }
    __syncthreads();
    const uint32_t below = (1u << lane) - 1u;
    return list_a ? warp_base[0][warp] + __popc(mask_a & below) : n - 1 - (warp_base[1][warp] + __popc(mask_b & below));
}

// resident CTAs per SM the fixed-k instantiation of the pack kernel is compiled for (engine.h: register budget)
#ifndef MSBWT_PACK_CTAS_FIXED_K
#define MSBWT_PACK_CTAS_FIXED_K 8
#endif
// 16 bytes from global to shared memory without passing through registers (L2 only: the bytes are read once)
__device__ __forceinline__ void stage_cp_async16(void *smem, const void *gmem) {
    const uint32_t dst = (uint32_t)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(gmem) : "memory");
}

constexpr uint32_t kPackThreads = 256;   // CTA size of the pack / seed kernels (append_live counts eight warps)
constexpr uint32_t kPackSmemMaxK = 160;  // 256 * k + 64 bytes of shared memory; longer k-mers read global memory
constexpr uint32_t kPackMaxWords = (kPackSmemMaxK + kPairSymsPerWord - 1) / kPairSymsPerWord;  // 5

// Seeds one all-ACGT k-mer given as 2-bit words (`get(w)`, w < nw: the k-mer's last symbol in the top
// bits of word 0): suffix-table lookup at the depth acgt_table_depth picks, then either the final count
// (written by the caller) or the remaining symbols re-aligned to the top of word 0 and stored for the
// search kernel.  Returns through the reference arguments; stores words 1.. itself.
// What seed_acgt needs to know about the batch, uniform over its queries and computed once on the host by the
// launch wrappers: the suffix-table level an all-ACGT k-mer of this length starts from (list_a_table_depth), that
// level's array, the flag the pair / one-step kernels read their resume depth from, and whether list A takes it.
struct SeedPlan {
    const void *tab;
    uint32_t depth;
    uint32_t flag;
    uint32_t list_a;
};
inline SeedPlan make_seed_plan(const IndexView &ix, uint32_t k) {
    SeedPlan p{nullptr, list_a_table_depth(ix, k), 0u, 0u};
    // the quad kernel finishes a remainder with one-step ranks; the pair kernel cannot
    p.list_a = (list_a_stride(ix) != 2u || ((k - p.depth) & 1u) == 0) ? 1u : 0u;
    if (p.depth) {
        const uint32_t back = ix.table_s - p.depth;  // 0..3
        // read by the pair / one-step kernels only (back <= 1 there); under an oct image the oct kernel derives the
        // depth from k itself and reads bit 30 of a list-A entry as "the final-step line of this query overflowed"
        p.flag = (!ix.oct && back < 2u) ? back + 1u : 0u;
        p.tab = back == 0 ? ix.table : (back == 1 ? ix.table2 : (back == 2 ? ix.table3 : ix.table4));
    }
    return p;
}

template <bool WIDE, class GetWord>
__device__ __forceinline__ void seed_acgt(const IndexView &ix, uint32_t k, const SeedPlan &plan, uint32_t nw, GetWord get, const PackedLayout &lay,
                                          uint64_t q, uint64_t *__restrict__ packed, uint64_t &lo, uint64_t &hi,
                                          uint32_t &flag, bool &list_a, bool &finished, uint64_t &word0) {
    const uint32_t done = plan.depth;
    list_a = plan.list_a != 0;
    lo = 0; hi = ix.total; flag = plan.flag;
    if (done) {
        const uint64_t e = get(0) >> (64u - 2u * done);
        if constexpr (WIDE) {
            const ulonglong2 v = __ldg(reinterpret_cast<const ulonglong2 *>(plan.tab) + e);
            lo = v.x; hi = v.y;
        } else {
            const uint2 v = __ldg(reinterpret_cast<const uint2 *>(plan.tab) + e);
            lo = v.x; hi = v.y;
        }
    }
    finished = lo == hi || done == k;
    if (finished) return;
    const uint32_t rest = k - done;
    if (list_a) {
        const uint32_t nout = (rest + kPairSymsPerWord - 1) / kPairSymsPerWord;
        uint64_t cur = get(0);
#pragma unroll
        for (uint32_t w = 0; w < kPackMaxWords; w++) {
            if (w < nout && w < nw) {  // nout <= nw; nw is a constant in the fixed-k instantiations
                const uint64_t nxt = (w + 1 < nw) ? get(w + 1) : 0;
                const uint64_t word = done ? (cur << (2u * done)) | (nxt >> (64u - 2u * done)) : cur;
                if (w == 0) word0 = word; else packed[lay.wx() + (uint64_t)(w - 1) * lay.n + q] = word;
                cur = nxt;
            }
        }
    } else {  // odd remainder without a usable table depth: re-expand to 3-bit symbols for the one-step kernel
        const uint32_t nout = (rest + kSymsPerWord - 1) / kSymsPerWord;
        for (uint32_t w = 0; w < nout; w++) {
            uint64_t word = 0;
            const uint32_t cnt = min((uint32_t)kSymsPerWord, rest - w * kSymsPerWord);
            for (uint32_t i = 0; i < cnt; i++) {
                const uint32_t t = done + w * kSymsPerWord + i;  // consumption index within the k-mer
                uint64_t src = 0;
#pragma unroll
                for (uint32_t j = 0; j < kPackMaxWords; j++) if (j == (t >> 5)) src = get(j);
                const uint32_t c = (uint32_t)(src >> (62u - 2u * (t & 31u))) & 3u;
                word |= (uint64_t)((0x5321u >> (4u * c)) & 7u) << (60 - 3 * i);
            }
            if (w == 0) word0 = word; else packed[lay.wx() + (uint64_t)(w - 1) * lay.n + q] = word;
        }
    }
}

// General path of the pack kernel, one symbol at a time from `src` (k bytes): k-mers holding a symbol
// outside ACGT (list B, 3 bits per symbol, table depth table_s when the last table_s symbols are ACGT)
// and k-mers longer than kPackSmemMaxK.
template <bool WIDE>
__device__ __forceinline__ void seed_general(const IndexView &ix, const uint8_t *src, uint32_t k, const PackedLayout &lay,
                                          uint64_t q, uint64_t *__restrict__ packed, uint64_t &lo, uint64_t &hi,
                                          uint32_t &flag, bool &finished, uint64_t &word0, bool &bad) {
    const uint32_t ts = ix.table_s;
    uint32_t na = 0;
    uint64_t tidx = 0;
    for (uint32_t t = 0; t < k; t++) {
        const uint32_t sy = src[k - 1 - t];
        const bool ok = sy < 8u && ((0x2Eu >> sy) & 1u) != 0;  // {1,2,3,5}
        bad |= sy >= (uint32_t)kAlphabet;
        if (t < ts && na == t && ok) {
            tidx = (tidx << 2) | ((sy - 1u - (sy >> 2)) & 3u);
            na++;
        }
    }
    const uint32_t done = (ts && na >= ts) ? ts : 0u;
    lo = 0; hi = ix.total; flag = done ? 1u : 0u;
    if (done) {
        if constexpr (WIDE) {
            const ulonglong2 v = __ldg(reinterpret_cast<const ulonglong2 *>(ix.table) + tidx);
            lo = v.x; hi = v.y;
        } else {
            const uint2 v = __ldg(reinterpret_cast<const uint2 *>(ix.table) + tidx);
            lo = v.x; hi = v.y;
        }
    }
    finished = lo == hi || done == k;
    if (finished) return;
    const uint32_t rest = k - done;
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

__device__ __forceinline__ uint64_t reverse_symbol_pairs(uint64_t x) {
    const uint64_t r = __brevll(x);
    return ((r >> 1) & 0x5555555555555555ull) | ((r & 0x5555555555555555ull) << 1);
}

}  // namespace msbwt
