// layout.h -- the device-resident block image of the BWT.
//
// The reference keeps the RLE byte stream plus a sampled struct-of-arrays index
// (`ref_index[]`, `fm_index[6][]`, one sample per 256 symbols: src/rle_bwt.rs:14-24,
// 387-467) and answers rank by a serial byte scan (src/rle_bwt.rs:221-238).
// constrain_range is exactly [C[s]+rank(s,l), C[s]+rank(s,h)) (SURVEY.md facts
// table), so any exact rank structure is bit-exact.  Ours:
//
//   one 64-byte block per 128 BWT symbols, made of two self-contained 32-byte halves
//   (half i covers block offsets 64i .. 64i+63):
//
//       word 0,1   u32 ckpt[2i], ckpt[2i+1]   two of the four checkpoints: occurrences of
//                                             A, C (half 0) / G, T (half 1) before the BLOCK,
//                                             relative to the block's superblock
//       word 2,3   plane0 of symbols 64i..64i+31, 64i+32..64i+63  (bit 0 of each symbol)
//       word 4,5   plane1 ...                                      (bit 1)
//       word 6,7   plane2 ...                                      (bit 2)
//
//   3-bit symbols as bit-planes: 32 symbols are matched against a query symbol with
//   3 logic ops + 1 popc.  A half is exactly one 256-bit ld.global.nc / one 32-byte
//   sector.  Two kernel mappings read it (kernels.cu): one thread per query (both
//   halves, two loads) when the index is L2-resident, or a lane pair per query (one
//   load per lane = one coalesced 64-byte request) when it lives in HBM.  Measured on
//   B200 (profiles/): random reads of 32, 64 and 128 bytes all sustain ~39 G reads/s
//   from HBM and 64-byte reads ~270 G reads/s from L2, so the block is sized by what a
//   rank needs, not by bandwidth.
//   The rare symbols $ and N keep their checkpoints in a side array
//   `aux[blk] = {u32 n$, u32 nN}` that only k-mers containing $ / N touch.
//
//   Positions past the end of the BWT in the last block hold symbol 7 (matches
//   nothing).  There is always a block for position N itself (N>>7), so h == N needs
//   no special case.
//
//   A superblock is 2^sb_shift blocks (default 2^25 blocks = 2^32 symbols) so the
//   per-block counters fit u32 for any N; `cbase[sb][s]` (u64, 8 per superblock) =
//   C[s] + occurrences of s before the superblock.  rank+C for (s,pos) is
//   cbase[blk>>sb_shift][s] + ckpt_s + popc(match & below(pos&127)).
#pragma once
#include <cstdint>

namespace msbwt {

constexpr int kBlockShift = 7;
constexpr int kBlockSyms = 1 << kBlockShift;
constexpr int kBlockBytes = 64;
constexpr int kWordsPerBlock = 16;    // u32 words
constexpr int kAlphabet = 6;          // $ACGNT (src/msbwt_core.rs:4)
constexpr int kDefaultSuperShift = 25;
constexpr int kSymsPerWord = 21;      // packed query word: 21 x 3-bit symbols, first-consumed symbol in the top bits
constexpr int kMaxSuperInSmem = 64;

// index of symbol s's checkpoint within ckpt[4] (A=1,C=2,G=3,T=5); $/N use `aux`
__host__ __device__ constexpr int ckpt_slot(int s) { return s == 5 ? 3 : s - 1; }

// Suffix table (the reference author's planned-but-unimplemented `kmer_cache`,
// src/msbwt_core.rs:133-146, src/rle_bwt.rs:332-346): table[idx] = the BWT range after the
// first `table_s` backward-search steps, for every ACGT-only suffix; idx = the consumed
// symbols as base-4 digits (A,C,G,T = 0..3), first consumed symbol most significant.
// Entries are {u32 l, u32 h} when N < 2^32 (one superblock), {u64 l, u64 h} otherwise.
constexpr int kMaxTableS = 15;

struct IndexView {
    const uint4 *blocks;     // nblocks * 4 uint4 (64 B per block)
    const uint32_t *aux;     // nblocks * 2  ($, N checkpoints)
    const uint64_t *cbase;   // n_super * 8
    const void *table;       // 4^table_s entries, or nullptr
    uint64_t total;          // N
    uint64_t nblocks;        // (N >> 7) + 1
    uint32_t n_super;
    uint32_t sb_shift;
    uint32_t table_s;        // 0 = no table
};

}  // namespace msbwt
