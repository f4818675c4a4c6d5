/*
 * msbwt_oracle.h -- CPU restatement of msbwt2's RleBWT query path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may build, load or call it, and only as the checker
 * or as the timed CPU baseline.  The product (rust-msbwt_b200/) never links
 * or imports it.
 *
 * Parity status: PINNED.  The reference is Rust-only and cannot be compiled in
 * this image (no cargo/rustc), so this restatement is pinned against every
 * known-answer test the reference's own test-suite holds for the path
 * (tests/test_oracle_kat.py lists them with file:line).
 *
 * Each function cites the reference lines (relative to /root/reference/) whose
 * behaviour it restates.
 */
#ifndef MSBWT_ORACLE_H
#define MSBWT_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* src/msbwt_core.rs:4-14 */
#define ORC_VC_LEN 6
#define ORC_LETTER_BITS 3
#define ORC_NUMBER_BITS 5
#define ORC_NUM_POWER 32
#define ORC_MASK 0x07
#define ORC_COUNT_MASK 0x1F

/* status codes: OK, the io::Error returns of load_numpy_file, and the places
 * where the reference panics (reported instead of aborting the test process) */
enum {
    ORC_OK = 0,
    ORC_ERR_IO = 1,          /* open/metadata failure            rle_bwt.rs:84,88 */
    ORC_ERR_SHORT_HEADER = 2,/* read_exact of bytes 10..skip     rle_bwt.rs:101-112 */
    ORC_ERR_SIZE_MISMATCH = 3,/* shape[0] != remaining file size rle_bwt.rs:128-136 */
    ORC_PANIC_SHORT_FILE = 10,/* < 10 bytes                      rle_bwt.rs:91-93 */
    ORC_PANIC_HEADER_PARSE = 11,/* utf8/json/shape unwrap        rle_bwt.rs:115-125 */
    ORC_PANIC_BAD_SYMBOL = 12 /* symbol >= 6: array index panic  rle_bwt.rs:371, msbwt_core.rs:127 */
};

/* src/msbwt_core.rs:18-24 */
typedef struct { uint64_t l, h; } orc_range;

typedef struct orc_rle_bwt orc_rle_bwt;

/* RleBWT::with_bin_power (rle_bwt.rs:309-322); RleBWT::new == bin_power 8 */
orc_rle_bwt *orc_new(unsigned bin_power);
void orc_free(orc_rle_bwt *b);

/* BWT::load_vector (rle_bwt.rs:59-66): copies `len` RLE bytes, then standard_init */
int orc_load_vector(orc_rle_bwt *b, const uint8_t *rle, uint64_t len);
/* BWT::load_numpy_file (rle_bwt.rs:81-155) */
int orc_load_numpy_file(orc_rle_bwt *b, const char *path);

uint64_t orc_get_symbol_count(const orc_rle_bwt *b, uint8_t sym); /* rle_bwt.rs:172-174 */
uint64_t orc_get_total_size(const orc_rle_bwt *b);                /* rle_bwt.rs:191-193 */
uint64_t orc_start_index(const orc_rle_bwt *b, uint8_t sym);
uint64_t orc_end_index(const orc_rle_bwt *b, uint8_t sym);

/* table access for the literal fm_index/ref_index KATs (rle_bwt.rs:537-599) */
uint64_t orc_index_length(const orc_rle_bwt *b);
const uint64_t *orc_ref_index(const orc_rle_bwt *b);
const uint64_t *orc_fm_index(const orc_rle_bwt *b, uint8_t sym);
uint64_t orc_rle_len(const orc_rle_bwt *b);
const uint8_t *orc_rle_bytes(const orc_rle_bwt *b);

/* RleBWT::constrain_range (rle_bwt.rs:202-287); no bounds checks, like `unsafe` */
orc_range orc_constrain_range(const orc_rle_bwt *b, uint8_t sym, orc_range in);

/* BWT::count_kmer (msbwt_core.rs:125-161).  Returns ORC_PANIC_BAD_SYMBOL if any
 * symbol >= 6 (the reference asserts), else ORC_OK with *count set. */
int orc_count_kmer(const orc_rle_bwt *b, const uint8_t *kmer, uint64_t k, uint64_t *count);

/* The reference's single-threaded loop `for q in queries { count_kmer(q) }`
 * over n fixed-length k-mers stored back to back; `threads` > 1 gives the
 * naive static split (the stand-in for a rayon split). */
int orc_count_kmers_fixed(const orc_rle_bwt *b, const uint8_t *syms, uint32_t k,
                          uint64_t n, uint64_t *out, int threads);
/* variable-length form: query i is syms[offsets[i]..offsets[i+1]) */
int orc_count_kmers(const orc_rle_bwt *b, const uint8_t *syms, const uint64_t *offsets,
                    uint64_t n, uint64_t *out, int threads);

/* Accounting pass for bench.py's roofline figure: replays count_kmer for n
 * fixed-k queries and reports how many constrain_range calls the reference
 * executes (steps) and in how many of them l and h fall in different
 * `1<<block_shift`-symbol blocks (two_block_steps).  SURVEY.md section 8(d). */
int orc_count_kmers_stats(const orc_rle_bwt *b, const uint8_t *syms, uint32_t k, uint64_t n,
                          unsigned block_shift, uint64_t *steps, uint64_t *two_block_steps);
int orc_count_kmers_stats_pair(const orc_rle_bwt *b, const uint8_t *syms, uint32_t k, uint64_t n,
                               uint32_t table_s, uint32_t line_syms, unsigned block_shift, uint64_t *out);
int orc_count_kmers_stats_oct(const orc_rle_bwt *b, const uint8_t *syms, uint32_t k, uint64_t n,
                              uint32_t table_s, uint32_t sector_syms, uint32_t line_sectors,
                              unsigned block_shift, unsigned oct_bucket_shift, unsigned oct_syms, uint64_t *out);
int orc_count_kmers_stats_fin(const orc_rle_bwt *b, const uint8_t *syms, uint32_t k, uint64_t n,
                              uint32_t table_s, uint32_t sector_syms, uint32_t line_sectors,
                              unsigned block_shift, unsigned oct_bucket_shift, unsigned oct_syms,
                              unsigned fin_bucket_shift, unsigned fin_syms, uint64_t *out, uint64_t *fin);
int orc_count_kmers_stats_quad(const orc_rle_bwt *b, const uint8_t *syms, uint32_t k, uint64_t n,
                               uint32_t table_s, uint32_t sector_syms, uint32_t line_sectors,
                               unsigned block_shift, uint64_t *out);
/* Same replay, but for an engine that answers the first `skip` steps of every k-mer whose
 * last `skip` symbols are all ACGT from a precomputed suffix table: those steps are not
 * counted, *table_hits counts the k-mers that took the shortcut. */
int orc_count_kmers_stats_skip(const orc_rle_bwt *b, const uint8_t *syms, uint32_t k, uint64_t n,
                               unsigned block_shift, uint32_t skip, uint64_t *steps,
                               uint64_t *two_block_steps, uint64_t *table_hits);

/* bwt_converter.rs:26-80 convert_to_vec: ASCII "$ACGNT"(+'\n') -> RLE bytes.
 * Returns number of bytes written (call with out==NULL to size), or
 * (uint64_t)-1 where the reference panics on an unexpected symbol. */
uint64_t orc_convert_to_vec(const uint8_t *ascii, uint64_t n, uint8_t *out, uint64_t cap);

/* bwt_converter.rs:151-184 run encoding (symbol,count) -> bytes (same digit rule) */
uint64_t orc_encode_runs(const uint8_t *syms, const uint64_t *counts, uint64_t nruns,
                         uint8_t *out, uint64_t cap);

/* bwt_converter.rs:102-130 save_bwt_numpy: 96-byte header + payload */
int orc_save_bwt_numpy(const uint8_t *rle, uint64_t len, const char *path);

/* string_util.rs:3-32, 6-9, 12 */
extern const uint8_t ORC_INT_TO_STRING[6];
extern const uint8_t ORC_COMPLEMENT_INT[6];
uint8_t orc_string_to_int(uint8_t ascii);
void orc_convert_stoi(const uint8_t *ascii, uint64_t n, uint8_t *out);          /* :63-67 */
void orc_convert_itos(const uint8_t *syms, uint64_t n, uint8_t *out);           /* :80-88 */
void orc_reverse_complement_i(const uint8_t *syms, uint64_t n, uint8_t *out);   /* :45-50 */

#ifdef __cplusplus
}
#endif
#endif
