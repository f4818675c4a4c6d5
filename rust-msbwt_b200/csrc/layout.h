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

// ---- pair ("two-step") blocks -------------------------------------------------------------
//
// Measured on B200 (profiles/r1_gather_dram_bytes.csv): an L2 miss fills a whole 128-byte line
// whatever the size of the request, and HBM serves ~39 G random line fills/s.  A 64-byte block
// therefore wastes half of every line it pulls in.  The pair image spends the whole line on the
// query instead: one line answers TWO backward-search steps at once.
//
// For every BWT position j with B[j] = b in ACGT let LF(j) = C[b] + rank(b, j) and a = B[LF(j)].
// If a is in ACGT too, position j carries the pair code 4*idx(b) + idx(a) (idx: A,C,G,T = 0..3)
// and is "valid"; otherwise it is invalid.  Then for every 0 <= i <= N (pure counting, no BWT
// property needed):
//
//     C[a] + rank(a, C[b] + rank(b, i))  ==  C2[b,a] + #{ j < i : code(j) == (b,a) }
//     with  C2[b,a] = C[a] + rank(a, C[b]),
//
// i.e. two successive constrain_range calls (first b, then a; src/rle_bwt.rs:202-287) equal one
// rank over the pair codes.  Bit-exact by construction; k-mers containing $ or N, and an odd
// leftover step, use the one-step blocks above.
//
//   one 128-byte line per 96 BWT positions = four 32-byte quarters; quarter t (= idx(b)):
//       word 0..3   u32 ckpt[4t + a], a = 0..3: occurrences of pair (t,a) before the block.
//                   N < 2^32: ABSOLUTE, C2[t,a] included (no base lookup at all);
//                   otherwise relative to the pair superblock, base in c2base[sb][16] (u64).
//       word 4..7   positions 24t .. 24t+23 as five bit-planes of 24 bits: word 4+p holds plane p
//                   (p = 0..3 = the four code bits) in bits 0..23; plane 4 (valid) is spread over
//                   the top bytes of words 4,5,6 (bits 0-7, 8-15, 16-23).  Top byte of word 7: 0.
//   A quad of lanes reads the line with one 256-bit load each (one coalesced 128-byte request).
//   Positions >= N are invalid.  There is always a line for position N (N / 96).
constexpr int kPairSyms = 96;
constexpr int kPairBytes = 128;
constexpr int kPairQuarterSyms = 24;
constexpr int kPairWords = 32;        // u32 words per line
constexpr int kPairSymsPerWord = 32;  // packed query word of the pair path: 32 x 2-bit ACGT symbols

// ---- quad ("four-step") sectors -----------------------------------------------------------
//
// The same identity composed once more (induction over the pair identity above): for the quad code
// c = (b0,b1,b2,b3) of position j -- b0 = B[j], b1 = B[LF(j)], b2 = B[LF^2(j)], b3 = B[LF^3(j)], all
// four in ACGT, else j is invalid -- and every 0 <= i <= N,
//
//     four successive constrain_range calls (b0 first) applied to i  ==  C4[c] + #{ j < i : code4(j) == c }
//     with  C4[c] = those four calls applied to 0,
//
// because LF^2 restricted to one pair code is the order-preserving bijection onto
// [C2[code], C2[code] + count(code)).  HBM is large (180 GB) and an L2 miss costs one 128-byte line
// fill whatever it asks for, so this image spends memory to save line fills: ONE occurrence bit-vector
// per quad code (256 of them), cut into self-contained 32-byte sectors:
//
//     sector (c, s), s = position / 224:   word 0      u32 checkpoint: #{ j < 224 s : code4(j) == c };
//                                                      N < 2^32: ABSOLUTE, C4[c] included; otherwise
//                                                      relative to the quad superblock (2^sb_shift4
//                                                      sectors), base in c4base[sb][256] (u64)
//                                          word 1..7   bit t of word 1+w: code4(224 s + 32 w + t) == c
//     address = ((c * nsec4) + s) * 32 bytes, nsec4 = N / 224 + 2 (the sector of position N, plus one
//     trailing all-zero sector per code whose checkpoint is the code's total -- the builder reads it).
//
// Code-major, so the sectors of l and h (a read-set range is a few dozen positions wide) share a
// 128-byte line 97 % of the time.  One thread per query, one 256-bit load per boundary, <= 7 POPC; a
// 31-mer seeded by a depth-15 suffix table needs 4 line fills instead of the pair image's 8.8.
// 256 * N / 7 bytes: 55 GB at N = 1.51 G, 110 GB at N = 3.02 G.  Chosen automatically only when the
// index lives in HBM and the image fits the device comfortably (capi.cu).
constexpr int kQuadSyms = 224;        // positions per sector
constexpr int kQuadWords = 8;         // u32 words per sector
constexpr int kQuadSectorBytes = 32;
constexpr int kQuadCodes = 256;
constexpr int kQuadMaxSuperShift = 24;  // 2^24 sectors * 224 positions < 2^32
constexpr int kQuadMaxSuperInSmem = 8;  // rows of 256 u64 staged in shared memory (16 KB)

// ---- oct lines: kOctSyms steps per line --------------------------------------------------------
//
// One more composition (m = kOctSyms = 10 symbols; the image began with eight, hence the name):
// code_m(j) = the m symbols B[j], B[LF j], .., B[LF^(m-1) j] as base-4 digits, the first most significant
// (valid when all are ACGT), and
//
//     m successive constrain_range calls applied to i  ==  Cm[c] + #{ j < i : code_m(j) == c }.
//
// 4^m codes: a bit-vector per code is out of reach, but each code is rare and on a read set its occurrences
// come in RUNS of consecutive BWT positions: the suffixes that share a long prefix -- the reads covering
// one genome position -- sit next to each other and are preceded by the same m symbols (mean run 11 on
// error-free 30x reads, 3 with 1 % errors).  So the occurrences are stored explicitly as runs.  BWT
// positions are cut into buckets of 2^b (b = the image's bucket shift, chosen when it is built); one
// 128-byte line per (code, bucket), code-major (`address = (c * nbuck8 + (pos >> b)) * 128`):
//
//     word 0      u32 checkpoint: Cm[c] + #{ j < bucket start : code_m(j) == c }   (N < 2^32 only)
//     word 1      u32 number of runs of c in the bucket
//     word 2..31  up to 30 runs `(len << b) | offset within the bucket`, any order; unused words are 0
//
// rank_m(c, p) = word0 + sum over runs of clamp((p & (2^b-1)) - offset, 0, len): one line fill, no order
// needed.  A run never crosses a multiple of 2^cs, cs = oct_chunk_shift(b) <= b, so len <= 2^cs fits its
// 32-b bits and no run crosses a bucket.  When a bucket holds more than 30 runs of one code (word 1 > 30:
// low-complexity text) the kernel falls back to quad / one-step ranks for that query-step, so the result is
// exact on any input.  128 * 4^m * (N / 2^b + 1) bytes; b is the largest shift that keeps the mean number of
// runs per line <= kOctTargetRuns (b = 24, 8 B/symbol, on 30x reads with 1 % errors), never below one bucket
// (128 MB).  N < 2^32: built through the pair and quad images (LF^4 per step of the code builder); the quad image
// may stay beside it and then serves remainders of 4..m-1 symbols.
//
// 64-bit positions (N >= 2^32 -- the reference is u64 throughout, src/msbwt_core.rs:18-24 -- or several
// superblocks): the same lines with a 40-bit checkpoint,
//     word 0      low 32 bits of the checkpoint
//     word 1      min(number of runs, kOctCapacity + 1) | (checkpoint >> 32) << 8
// everything else is bucket-relative and unchanged.  A quad image of such an index has no room (36.6 B per
// position), so the code builder walks LF through the one-step blocks instead and the kernel
// (count_kmers_oct_kernel<.., WIDE = true>) takes remainders and fallbacks as one-symbol steps.  N < 2^40.
//
// Why ten: a 31-mer then is a suffix-table entry at depth 11 -- 4^11 entries of 8 bytes = 33 MB, resident in
// L2 -- plus TWO lines: two HBM requests per query instead of three with eight symbols per line and a
// depth-15 table (8.6 GB, one HBM request per lookup).
constexpr int kOctSyms = 10;
constexpr int kOctCodeBits = 2 * kOctSyms;
constexpr int kOctCodes = 1 << kOctCodeBits;
constexpr int kOctLineBytes = 128;
constexpr int kOctLineWords = 32;
constexpr int kOctCapacity = 30;      // runs per line
constexpr int kOctMinShift = 8, kOctMaxShift = 24;
constexpr int kOctAutoMinShift = 16;  // automatic choice: 16..24
constexpr int kOctTargetRuns = 6;
__host__ __device__ constexpr int oct_chunk_shift(int b) {  // cs = min(31 - b, 10, b)
    int c = 31 - b;
    if (c > 10) c = 10;
    if (c > b) c = b;
    return c;
}

// ---- final-step lines (specification: oracle/final_step.py; holds no positions, so the same for any N) ----
//
// The LAST kFinSyms symbols a count_kmer consumes need no rank, only #{ j in [l, h) : code_20(j) == c }: no
// checkpoints, nothing stored for codes that do not occur.  `1 << lb` lines of 128 bytes per bucket of `1 << b`
// positions (b <= 16); code c lives in line `(bucket << lb) | (fin_mix40(c) & (2^lb - 1))` under the tag
// `fin_mix40(c) >> lb` (fin_mix40 is a bijection of the 40-bit codes, lb >= 12 so that a tag fits kFinTagBits):
//     word 0      words in use after it (0..31), or kFinOverflow: the query takes the oct steps instead
//     then groups `(tag << 4) | nruns` (1..15) followed by nruns words `(len << 16) | offset in the bucket`; a code
//     with more than 15 runs in the bucket has several groups, ADJACENT to one another (the builder writes a line's
//     groups in key order), so a reader that has found a code's first group finds the others right behind it
// A range over two buckets also takes the oct steps.  A 31-mer = an L2-resident depth-11 table entry + ONE line.
constexpr int kFinSyms = 20;
constexpr int kFinCodeBits = 2 * kFinSyms;
constexpr int kFinTagBits = 28;
constexpr int kFinLineBytes = 128;
constexpr int kFinLineWords = 32;
constexpr uint32_t kFinOverflow = 0xFFFFFFFFu;
__host__ __device__ inline uint64_t fin_mix40(uint64_t c) {  // the same constants as oracle/final_step.py
    constexpr uint64_t m = (1ull << kFinCodeBits) - 1ull;
    c &= m;
    c ^= c >> 20;
    c = (c * 0x9E3779B97Full) & m;
    c ^= c >> 20;
    c = (c * 0xC2B2AE3D27ull) & m;
    c ^= c >> 20;
    return c;
}

struct IndexView {
    const uint4 *blocks;     // nblocks * 4 uint4 (64 B per block)
    const uint32_t *aux;     // nblocks * 2  ($, N checkpoints)
    const uint64_t *cbase;   // n_super * 8
    const void *table;       // 4^table_s entries, or nullptr
    const void *table2;      // 4^(table_s-1) entries (kept when a pair or quad image exists), or nullptr
    const void *table3;      // 4^(table_s-2), 4^(table_s-3) entries (kept when the quad image exists)
    const void *table4;
    const uint4 *pair;       // npair * 8 uint4 (128 B per 96 positions), or nullptr
    const uint64_t *c2base;  // n_super2 * 16 (u64), only when positions are 64-bit
    const uint4 *quad;       // 256 * nsec4 sectors of 32 B (2 uint4 each), or nullptr
    const uint64_t *c4base;  // n_super4 * 256 (u64), only when positions are 64-bit
    uint64_t nsec4;          // sectors per quad code: N / 224 + 2
    const uint4 *oct;        // 4^m * nbuck8 lines of 128 B (8 uint4 each), or nullptr
    uint64_t nbuck8;         // buckets per oct code: (N >> oct_shift) + 1
    uint32_t oct_shift;      // b: log2 of the oct bucket size
    uint64_t total;          // N
    uint64_t nblocks;        // (N >> 7) + 1
    uint64_t npair;          // N / 96 + 1
    uint32_t n_super;
    uint32_t sb_shift;
    uint32_t n_super2;       // pair superblocks (2^sb_shift lines each)
    uint32_t n_super4;       // quad superblocks (2^sb_shift4 sectors each)
    uint32_t sb_shift4;
    uint32_t table_s;        // 0 = no table
    const uint4 *fin;        // final-step lines (8 uint4 each), or nullptr
    uint32_t fin_shift;      // b: log2 of the bucket size
    uint32_t fin_lb;         // log2 of the lines per bucket
};

}  // namespace msbwt
