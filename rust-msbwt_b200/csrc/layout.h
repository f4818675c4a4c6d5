// layout.h -- the device-resident block image of the BWT ("layout A").
//
// The reference keeps the RLE byte stream plus a sampled struct-of-arrays index
// (`ref_index[]`, `fm_index[6][]`, one sample per 256 symbols: src/rle_bwt.rs:14-24,
// 387-467) and answers rank by a serial byte scan (src/rle_bwt.rs:221-238).
// constrain_range is exactly [C[s]+rank(s,l), C[s]+rank(s,h)) (SURVEY.md facts
// table), so any exact rank structure is bit-exact.  Ours:
//
//   one 128-byte block per 256 BWT symbols (same stride as the reference's default
//   bin), fetched by 8 lanes x one 16-byte ld.global.nc each.  Lane j's chunk is
//
//       u32 hdr_j | u32 plane0_j | u32 plane1_j | u32 plane2_j
//
//   plane_b_j bit i = bit b of the symbol at block offset 32*j + i  (3-bit symbols,
//   bit-planes so a lane's 32 symbols are matched with 3 logic ops + 1 popc),
//   hdr_0..hdr_5 = number of $,A,C,G,N,T before the block, relative to the block's
//   superblock (u32); hdr_6/hdr_7 are zero.
//
//   Positions past the end of the BWT in the last block hold symbol 7 (matches
//   nothing).  There is always a block for position N itself (N>>8), so h == N needs
//   no special case.
//
//   A superblock is 2^sb_shift blocks (default 2^24 blocks = 2^32 symbols) so the
//   per-block counters fit u32 for any N; `cbase[sb][s]` (u64, 8 per superblock) =
//   C[s] + occurrences of s before the superblock.  rank+C for (s,pos) is
//   cbase[blk>>sb_shift][s] + hdr_s + popc(match & below(pos&255)).
#pragma once
#include <cstdint>

namespace msbwt {

constexpr int kBlockShift = 8;
constexpr int kBlockSyms = 1 << kBlockShift;
constexpr int kBlockBytes = 128;
constexpr int kLanesPerBlock = 8;     // 16 B per lane
constexpr int kWordsPerBlock = 32;    // u32 words
constexpr int kAlphabet = 6;          // $ACGNT (src/msbwt_core.rs:4)
constexpr int kDefaultSuperShift = 24;
constexpr int kSymsPerWord = 21;      // packed query word: 21 x 3-bit symbols, first-consumed symbol in the top bits
constexpr int kMaxSuperInSmem = 64;

struct IndexView {
    const uint4 *blocks;     // nblocks * 8 uint4
    const uint64_t *cbase;   // n_super * 8
    uint64_t total;          // N
    uint64_t nblocks;        // (N >> 8) + 1
    uint32_t n_super;
    uint32_t sb_shift;
};

}  // namespace msbwt
