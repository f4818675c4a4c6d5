/*
 * msbwt_oracle.c -- CPU restatement of msbwt2's RleBWT (TEST INFRASTRUCTURE ONLY,
 * see msbwt_oracle.h).  Keeps the reference's struct-of-arrays sampled index
 * (`fm_index[6][]`, `ref_index[]`, `bin_power`) and its byte-at-a-time run scan
 * so that timing it reflects the reference's memory behaviour.
 *
 * Build: see oracle/Makefile (gcc -O3 -march=native -shared -fPIC -pthread).
 */
#include "msbwt_oracle.h"

#include <errno.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>

/* rle_bwt.rs:14-24 */
struct orc_rle_bwt {
    uint8_t *bwt;
    uint64_t bwt_len;
    uint64_t symbol_counts[ORC_VC_LEN];
    uint64_t start_index[ORC_VC_LEN];
    uint64_t end_index[ORC_VC_LEN];
    uint64_t *fm_index[ORC_VC_LEN];
    uint64_t *ref_index;
    uint64_t index_length;
    uint64_t total_size;
    unsigned bin_power;
    uint64_t bin_size;
};

orc_rle_bwt *orc_new(unsigned bin_power) {
    orc_rle_bwt *b = (orc_rle_bwt *)calloc(1, sizeof(*b));
    if (!b) return NULL;
    b->bin_power = bin_power;
    b->bin_size = (uint64_t)1 << bin_power;
    return b;
}

static void drop_tables(orc_rle_bwt *b) {
    for (int y = 0; y < ORC_VC_LEN; y++) { free(b->fm_index[y]); b->fm_index[y] = NULL; }
    free(b->ref_index); b->ref_index = NULL;
    b->index_length = 0;
}

void orc_free(orc_rle_bwt *b) {
    if (!b) return;
    drop_tables(b);
    free(b->bwt);
    free(b);
}

/* rle_bwt.rs:352-384 calculate_totals */
static int calculate_totals(orc_rle_bwt *b) {
    uint8_t prev = 255;
    uint64_t power = 1;
    memset(b->symbol_counts, 0, sizeof(b->symbol_counts));
    for (uint64_t i = 0; i < b->bwt_len; i++) {
        uint8_t v = b->bwt[i];
        uint8_t c = v & ORC_MASK;
        if (c >= ORC_VC_LEN) return ORC_PANIC_BAD_SYMBOL; /* symbol_counts[c] would panic */
        power = (c == prev) ? power * ORC_NUM_POWER : 1;
        prev = c;
        b->symbol_counts[c] += (uint64_t)(v >> ORC_LETTER_BITS) * power;
    }
    uint64_t sum = 0;
    for (int i = 0; i < ORC_VC_LEN; i++) {
        b->start_index[i] = sum;
        sum += b->symbol_counts[i];
        b->end_index[i] = sum;
    }
    b->total_size = b->end_index[ORC_VC_LEN - 1];
    return ORC_OK;
}

/* rle_bwt.rs:387-467 construct_fmindex.  For bin x (position x*bin_size):
 * ref_index[x] = byte offset of the run that contains that position,
 * fm_index[y][x] = count of y strictly before that run. */
static int construct_fmindex(orc_rle_bwt *b) {
    drop_tables(b);
    /* ceil(total/bin) + 1, the reference does this in f64 (:390) */
    uint64_t n = (b->total_size + b->bin_size - 1) / b->bin_size + 1;
    b->index_length = n;
    for (int y = 0; y < ORC_VC_LEN; y++) {
        b->fm_index[y] = (uint64_t *)calloc(n, sizeof(uint64_t));
        if (!b->fm_index[y]) return ORC_ERR_IO;
    }
    b->ref_index = (uint64_t *)calloc(n, sizeof(uint64_t));
    if (!b->ref_index) return ORC_ERR_IO;

    uint64_t running[ORC_VC_LEN] = {0};
    uint64_t run_count = 0, power = 1, bin_end = 0, bin_id = 0, bwt_index = 0, run_start = 0;
    uint8_t prev = 0; /* the reference starts from symbol 0 with an empty run (:406) */

    for (uint64_t x = 0; x < b->bwt_len; x++) {
        uint8_t v = b->bwt[x];
        uint8_t c = v & ORC_MASK;
        if (c == prev) {
            run_count += (uint64_t)(v >> ORC_LETTER_BITS) * power;
            power *= ORC_NUM_POWER;
        } else {
            /* flush every bin edge covered by the run that just ended (:421-428) */
            while (bwt_index + run_count > bin_end) {
                b->ref_index[bin_id] = run_start;
                for (int y = 0; y < ORC_VC_LEN; y++) b->fm_index[y][bin_id] = running[y];
                bin_end += b->bin_size;
                bin_id++;
            }
            running[prev] += run_count;
            bwt_index += run_count;
            prev = c;
            run_start = x;
            run_count = v >> ORC_LETTER_BITS;
            power = ORC_NUM_POWER;
        }
    }
    while (bwt_index + run_count > bin_end) { /* :443-450 */
        b->ref_index[bin_id] = run_start;
        for (int y = 0; y < ORC_VC_LEN; y++) b->fm_index[y][bin_id] = running[y];
        bin_end += b->bin_size;
        bin_id++;
    }
    running[prev] += run_count; /* :453-457 final entry holds the totals */
    b->ref_index[n - 1] = b->bwt_len;
    for (int y = 0; y < ORC_VC_LEN; y++) b->fm_index[y][n - 1] = running[y];
    return ORC_OK;
}

/* rle_bwt.rs:324-348 standard_init */
static int standard_init(orc_rle_bwt *b) {
    int rc = calculate_totals(b);
    if (rc != ORC_OK) return rc;
    return construct_fmindex(b);
}

int orc_load_vector(orc_rle_bwt *b, const uint8_t *rle, uint64_t len) {
    free(b->bwt);
    b->bwt = (uint8_t *)malloc(len ? len : 1);
    if (!b->bwt) return ORC_ERR_IO;
    if (len) memcpy(b->bwt, rle, len);
    b->bwt_len = len;
    return standard_init(b);
}

/* ---- tiny JSON reader: just enough to restate serde_json::from_str + ["shape"][0].as_u64()
 * on the munged header dict (rle_bwt.rs:115-125) ---- */
typedef struct { const char *p, *end; int depth; int found; uint64_t shape0; } jctx;
static void j_ws(jctx *j) { while (j->p < j->end && (*j->p == ' ' || *j->p == '\n' || *j->p == '\t' || *j->p == '\r')) j->p++; }
static int j_value(jctx *j, int want_shape);
static int j_string(jctx *j, const char **s, size_t *n) {
    if (j->p >= j->end || *j->p != '"') return 0;
    j->p++;
    *s = j->p;
    while (j->p < j->end && *j->p != '"') { if (*j->p == '\\') j->p++; j->p++; }
    if (j->p >= j->end) return 0;
    *n = (size_t)(j->p - *s);
    j->p++;
    return 1;
}
static int j_number(jctx *j, int *is_u64, uint64_t *val) {
    const char *s = j->p;
    int neg = 0, frac = 0;
    if (j->p < j->end && *j->p == '-') { neg = 1; j->p++; }
    const char *d0 = j->p;
    uint64_t v = 0;
    while (j->p < j->end && *j->p >= '0' && *j->p <= '9') { v = v * 10 + (uint64_t)(*j->p - '0'); j->p++; }
    if (j->p == d0) { j->p = s; return 0; }
    if (j->p - d0 > 1 && *d0 == '0') return 0; /* JSON forbids leading zeros */
    if (j->p < j->end && *j->p == '.') { frac = 1; j->p++; const char *f0 = j->p; while (j->p < j->end && *j->p >= '0' && *j->p <= '9') j->p++; if (j->p == f0) return 0; }
    if (j->p < j->end && (*j->p == 'e' || *j->p == 'E')) { frac = 1; j->p++; if (j->p < j->end && (*j->p == '+' || *j->p == '-')) j->p++; const char *e0 = j->p; while (j->p < j->end && *j->p >= '0' && *j->p <= '9') j->p++; if (j->p == e0) return 0; }
    *is_u64 = !neg && !frac;
    *val = v;
    return 1;
}
static int j_lit(jctx *j, const char *w) { size_t n = strlen(w); if ((size_t)(j->end - j->p) >= n && !memcmp(j->p, w, n)) { j->p += n; return 1; } return 0; }
static int j_value(jctx *j, int want_shape) {
    j_ws(j);
    if (j->p >= j->end || ++j->depth > 64) return 0;
    int ok = 0;
    char c = *j->p;
    if (c == '{') {
        j->p++; j_ws(j);
        if (j->p < j->end && *j->p == '}') { j->p++; ok = 1; }
        else for (;;) {
            const char *k; size_t kn;
            j_ws(j);
            if (!j_string(j, &k, &kn)) break;
            j_ws(j);
            if (j->p >= j->end || *j->p != ':') break;
            j->p++;
            int is_shape = (j->depth == 1 && kn == 5 && !memcmp(k, "shape", 5));
            if (is_shape) j->found = 0; /* duplicate keys: serde keeps the last one */
            if (!j_value(j, is_shape)) break;
            j_ws(j);
            if (j->p < j->end && *j->p == ',') { j->p++; continue; }
            if (j->p < j->end && *j->p == '}') { j->p++; ok = 1; }
            break;
        }
    } else if (c == '[') {
        j->p++; j_ws(j);
        if (j->p < j->end && *j->p == ']') { j->p++; ok = 1; }
        else for (int idx = 0;; idx++) {
            j_ws(j);
            if (want_shape && idx == 0) {
                const char *s = j->p; int isu; uint64_t v;
                if (j_number(j, &isu, &v)) { if (isu) { j->found = 1; j->shape0 = v; } }
                else { j->p = s; if (!j_value(j, 0)) break; }
            } else if (!j_value(j, 0)) break;
            j_ws(j);
            if (j->p < j->end && *j->p == ',') { j->p++; continue; }
            if (j->p < j->end && *j->p == ']') { j->p++; ok = 1; }
            break;
        }
    } else if (c == '"') {
        const char *s; size_t n; ok = j_string(j, &s, &n);
    } else if (c == 't') ok = j_lit(j, "true");
    else if (c == 'f') ok = j_lit(j, "false");
    else if (c == 'n') ok = j_lit(j, "null");
    else { int isu; uint64_t v; ok = j_number(j, &isu, &v); }
    j->depth--;
    return ok;
}

/* in-place-ish substring replacement used to munge the python dict (:115-122) */
static char *replace_all(char *s, const char *from, const char *to) {
    size_t fl = strlen(from), tl = strlen(to), n = strlen(s), cnt = 0;
    for (char *p = s; (p = strstr(p, from)); p += fl) cnt++;
    char *out = (char *)malloc(n + cnt * (tl > fl ? tl - fl : 0) + 1), *o = out;
    for (char *p = s;;) {
        char *q = strstr(p, from);
        if (!q) { strcpy(o, p); break; }
        memcpy(o, p, (size_t)(q - p)); o += q - p;
        memcpy(o, to, tl); o += tl;
        p = q + fl;
    }
    free(s);
    return out;
}

static int utf8_valid(const uint8_t *s, size_t n) {
    for (size_t i = 0; i < n;) {
        uint8_t c = s[i];
        size_t extra = c < 0x80 ? 0 : (c >> 5) == 6 ? 1 : (c >> 4) == 14 ? 2 : (c >> 3) == 30 ? 3 : 4;
        if (extra == 4 || i + extra >= n + (extra == 0 ? 1 : 0)) return 0;
        for (size_t t = 1; t <= extra; t++) if ((s[i + t] >> 6) != 2) return 0;
        i += extra + 1;
    }
    return 1;
}

/* rle_bwt.rs:81-155 */
int orc_load_numpy_file(orc_rle_bwt *b, const char *path) {
    struct stat st;
    if (stat(path, &st) != 0) return ORC_ERR_IO;                       /* :84 */
    uint64_t full = (uint64_t)st.st_size;
    FILE *f = fopen(path, "rb");
    if (!f) return ORC_ERR_IO;                                         /* :88 */
    uint8_t fixed[10];
    if (fread(fixed, 1, 10, f) != 10) { fclose(f); return ORC_PANIC_SHORT_FILE; } /* :91-93 */
    size_t header_len = fixed[8] + 256u * fixed[9];                    /* :96, magic/version unchecked */
    size_t skip = 10 + header_len;
    if (skip % 16) skip = (skip / 16 + 1) * 16;                        /* :97-100 */
    size_t hn = skip - 10;
    char *hdr = (char *)malloc(hn + 1);
    if (fread(hdr, 1, hn, f) != hn) { free(hdr); fclose(f); return ORC_ERR_SHORT_HEADER; } /* :101-112 */
    hdr[hn] = 0;
    if (!utf8_valid((const uint8_t *)hdr, hn) || strlen(hdr) != hn) {
        /* from_utf8().unwrap() panics on invalid utf8; an embedded NUL is valid utf8 but
         * is rejected by the JSON parser right after, same outcome class */
        free(hdr); fclose(f); return ORC_PANIC_HEADER_PARSE;
    }
    hdr = replace_all(hdr, "'", "\"");
    hdr = replace_all(hdr, "False", "false");
    hdr = replace_all(hdr, "(", "[");
    hdr = replace_all(hdr, ")", "]");
    hdr = replace_all(hdr, ", }", "}");
    hdr = replace_all(hdr, ", ]", "]");
    hdr = replace_all(hdr, ",]", "]");
    jctx j = { hdr, hdr + strlen(hdr), 0, 0, 0 };
    int ok = j_value(&j, 0);
    if (ok) { j_ws(&j); ok = (j.p == j.end); }
    free(hdr);
    if (!ok || !j.found) { fclose(f); return ORC_PANIC_HEADER_PARSE; } /* :123-125 */
    /* full - skip is an unsigned subtraction in the reference; a file shorter than
     * `skip` already failed read_exact above */
    uint64_t disk = full - skip;
    if (j.shape0 != disk) { fclose(f); return ORC_ERR_SIZE_MISMATCH; } /* :128-136 */
    free(b->bwt);
    b->bwt = (uint8_t *)malloc(disk ? disk : 1);
    if (!b->bwt) { fclose(f); return ORC_ERR_IO; }
    uint64_t got = fread(b->bwt, 1, disk, f);                          /* :139-148 */
    fclose(f);
    if (got != disk) return ORC_ERR_SIZE_MISMATCH;
    b->bwt_len = disk;
    return standard_init(b);                                           /* :152 */
}

uint64_t orc_get_symbol_count(const orc_rle_bwt *b, uint8_t sym) { return b->symbol_counts[sym]; }
uint64_t orc_get_total_size(const orc_rle_bwt *b) { return b->total_size; }
uint64_t orc_start_index(const orc_rle_bwt *b, uint8_t sym) { return b->start_index[sym]; }
uint64_t orc_end_index(const orc_rle_bwt *b, uint8_t sym) { return b->end_index[sym]; }
uint64_t orc_index_length(const orc_rle_bwt *b) { return b->index_length; }
const uint64_t *orc_ref_index(const orc_rle_bwt *b) { return b->ref_index; }
const uint64_t *orc_fm_index(const orc_rle_bwt *b, uint8_t sym) { return b->fm_index[sym]; }
uint64_t orc_rle_len(const orc_rle_bwt *b) { return b->bwt_len; }
const uint8_t *orc_rle_bytes(const orc_rle_bwt *b) { return b->bwt; }

/* One boundary of rle_bwt.rs:202-287: the scan state the reference carries from
 * the low side into the high side when both fall in one bin (:246-249). */
typedef struct {
    uint64_t byte_pos;   /* compressed_index */
    uint64_t bwt_index;  /* symbols fully consumed */
    uint64_t run_len;    /* prev_count */
    uint64_t power;      /* power_multiple */
    uint64_t acc;        /* running result for `sym` */
    uint8_t run_sym;     /* prev_char */
} scan_state;

static inline void scan_seed(const orc_rle_bwt *b, uint8_t sym, uint64_t bin, scan_state *s) {
    s->byte_pos = b->ref_index[bin];                                   /* :205 / :251 */
    uint64_t tot = 0;
    for (int x = 0; x < ORC_VC_LEN; x++) tot += b->fm_index[x][bin];   /* :206-209 / :252-255 */
    s->bwt_index = tot;
    s->acc = b->start_index[sym] + b->fm_index[sym][bin];              /* :211-214 / :257 */
    s->run_sym = 255;
    s->run_len = 0;
    s->power = 1;
}

static inline void scan_until(const orc_rle_bwt *b, uint8_t sym, uint64_t target, scan_state *s) {
    const uint8_t *bw = b->bwt;
    while (s->bwt_index + s->run_len < target) {                       /* :221-238 / :264-281 */
        uint8_t v = bw[s->byte_pos];
        uint8_t c = v & ORC_MASK;
        if (c == s->run_sym) {
            s->run_len += (uint64_t)(v >> ORC_LETTER_BITS) * s->power;
            s->power *= ORC_NUM_POWER;
        } else {
            if (s->run_sym == sym) s->acc += s->run_len;
            s->bwt_index += s->run_len;
            s->run_len = v >> ORC_LETTER_BITS;
            s->run_sym = c;
            s->power = ORC_NUM_POWER;
        }
        s->byte_pos++;
    }
}

orc_range orc_constrain_range(const orc_rle_bwt *b, uint8_t sym, orc_range in) {
    orc_range out;
    scan_state s;
    uint64_t bin_l = in.l >> b->bin_power;                             /* :204 */
    scan_seed(b, sym, bin_l, &s);
    scan_until(b, sym, in.l, &s);
    out.l = s.acc;
    if (s.run_sym == sym) out.l += in.l - s.bwt_index;                 /* :240-243 */
    uint64_t bin_h = in.h >> b->bin_power;                             /* :246 */
    if (bin_h != bin_l) scan_seed(b, sym, bin_h, &s);                  /* :250-262; same bin keeps state */
    scan_until(b, sym, in.h, &s);
    out.h = s.acc;
    if (s.run_sym == sym) out.h += in.h - s.bwt_index;                 /* :283-285 */
    return out;
}

/* msbwt_core.rs:125-161 */
static inline uint64_t count_kmer_unchecked(const orc_rle_bwt *b, const uint8_t *kmer, uint64_t k) {
    orc_range r = { 0, b->total_size };
    for (uint64_t i = k; i-- > 0;) {
        if (r.h == r.l) return 0;                                      /* :151-153 */
        r = orc_constrain_range(b, kmer[i], r);
    }
    return r.h - r.l;
}

int orc_count_kmer(const orc_rle_bwt *b, const uint8_t *kmer, uint64_t k, uint64_t *count) {
    for (uint64_t i = 0; i < k; i++) if (kmer[i] >= ORC_VC_LEN) return ORC_PANIC_BAD_SYMBOL; /* :127 */
    *count = count_kmer_unchecked(b, kmer, k);
    return ORC_OK;
}

typedef struct {
    const orc_rle_bwt *b; const uint8_t *syms; const uint64_t *offsets; uint32_t k;
    uint64_t lo, hi; uint64_t *out; int rc;
} work_t;

static void *worker(void *arg) {
    work_t *w = (work_t *)arg;
    w->rc = ORC_OK;
    for (uint64_t i = w->lo; i < w->hi; i++) {
        const uint8_t *q; uint64_t k;
        if (w->offsets) { q = w->syms + w->offsets[i]; k = w->offsets[i + 1] - w->offsets[i]; }
        else { q = w->syms + i * (uint64_t)w->k; k = w->k; }
        int rc = orc_count_kmer(w->b, q, k, &w->out[i]);
        if (rc != ORC_OK) { w->rc = rc; return NULL; }
    }
    return NULL;
}

static int run_split(const orc_rle_bwt *b, const uint8_t *syms, const uint64_t *offsets, uint32_t k,
                     uint64_t n, uint64_t *out, int threads) {
    if (threads < 1) threads = 1;
    if ((uint64_t)threads > n) threads = n ? (int)n : 1;
    work_t *w = (work_t *)calloc((size_t)threads, sizeof(work_t));
    pthread_t *th = (pthread_t *)calloc((size_t)threads, sizeof(pthread_t));
    int rc = ORC_OK;
    for (int t = 0; t < threads; t++) {
        w[t].b = b; w[t].syms = syms; w[t].offsets = offsets; w[t].k = k; w[t].out = out;
        w[t].lo = n * (uint64_t)t / (uint64_t)threads;
        w[t].hi = n * (uint64_t)(t + 1) / (uint64_t)threads;
    }
    if (threads == 1) worker(&w[0]);
    else {
        for (int t = 0; t < threads; t++) pthread_create(&th[t], NULL, worker, &w[t]);
        for (int t = 0; t < threads; t++) pthread_join(th[t], NULL);
    }
    for (int t = 0; t < threads; t++) if (w[t].rc != ORC_OK) rc = w[t].rc;
    free(w); free(th);
    return rc;
}

int orc_count_kmers_fixed(const orc_rle_bwt *b, const uint8_t *syms, uint32_t k, uint64_t n,
                          uint64_t *out, int threads) {
    return run_split(b, syms, NULL, k, n, out, threads);
}

int orc_count_kmers(const orc_rle_bwt *b, const uint8_t *syms, const uint64_t *offsets, uint64_t n,
                    uint64_t *out, int threads) {
    return run_split(b, syms, offsets, 0, n, out, threads);
}

int orc_count_kmers_stats_skip(const orc_rle_bwt *b, const uint8_t *syms, uint32_t k, uint64_t n,
                               unsigned block_shift, uint32_t skip, uint64_t *steps,
                               uint64_t *two_block_steps, uint64_t *table_hits) {
    uint64_t st = 0, tb = 0, hits = 0;
    for (uint64_t i = 0; i < n; i++) {
        const uint8_t *q = syms + i * (uint64_t)k;
        int eligible = skip > 0 && k >= skip;
        for (uint32_t t = 0; t < k; t++) {
            if (q[t] >= ORC_VC_LEN) return ORC_PANIC_BAD_SYMBOL;
            if (t >= k - (eligible ? skip : 0) && !(q[t] == 1 || q[t] == 2 || q[t] == 3 || q[t] == 5)) eligible = 0;
        }
        hits += (uint64_t)eligible;
        const uint32_t first_counted = eligible ? skip : 0;
        orc_range r = { 0, b->total_size };
        uint32_t done = 0;
        for (uint32_t t = k; t-- > 0; done++) {
            if (r.h == r.l) break;
            if (done >= first_counted) {
                st++;
                if ((r.l >> block_shift) != (r.h >> block_shift)) tb++;
            }
            r = orc_constrain_range(b, q[t], r);
        }
    }
    *steps = st; *two_block_steps = tb;
    if (table_hits) *table_hits = hits;
    return ORC_OK;
}

int orc_count_kmers_stats(const orc_rle_bwt *b, const uint8_t *syms, uint32_t k, uint64_t n,
                          unsigned block_shift, uint64_t *steps, uint64_t *two_block_steps) {
    return orc_count_kmers_stats_skip(b, syms, k, n, block_shift, 0, steps, two_block_steps, NULL);
}

/* Accounting replay of the engine's PAIR path (rust-msbwt_b200/csrc/layout.h): which index lines a batch
 * must touch.  An all-ACGT k-mer starts from the suffix table at depth table_s or table_s-1 (whichever
 * leaves an even number of symbols) and then takes one pair step per two symbols -- one line of
 * `line_syms` positions per boundary; any other k-mer takes the one-step path (blocks of
 * 2^block_shift positions, depth table_s when its last table_s symbols are ACGT).  Ranges evolve
 * exactly as in count_kmer (msbwt_core.rs:125-161); a step is counted when the engine executes it
 * (range non-empty at its start).  out[0..5] = pair steps, pair steps whose l and h fall in
 * different lines, one-step steps, one-step steps over two blocks, table lookups, queries. */
int orc_count_kmers_stats_pair(const orc_rle_bwt *b, const uint8_t *syms, uint32_t k, uint64_t n,
                               uint32_t table_s, uint32_t line_syms, unsigned block_shift, uint64_t *out) {
    uint64_t ps = 0, p2 = 0, os = 0, o2 = 0, hits = 0;
    for (uint64_t i = 0; i < n; i++) {
        const uint8_t *q = syms + i * (uint64_t)k;
        int all_acgt = 1;
        uint32_t na = 0;
        for (uint32_t t = 0; t < k; t++) {
            const uint8_t sy = q[k - 1 - t];
            if (sy >= ORC_VC_LEN) return ORC_PANIC_BAD_SYMBOL;
            const int ok = sy == 1 || sy == 2 || sy == 3 || sy == 5;
            all_acgt &= ok;
            if (t < table_s && na == t && ok) na++;
        }
        uint32_t done = 0;
        int pair = 0;
        if (all_acgt) {
            if (table_s && k >= table_s) done = ((k - table_s) & 1u) ? table_s - 1 : table_s;
            else if (table_s && k + 1 == table_s) done = k;
            pair = ((k - done) & 1u) == 0;
        } else if (table_s && na >= table_s) {
            done = table_s;
        }
        hits += done != 0;
        orc_range r = { 0, b->total_size };
        uint32_t t = k;
        for (uint32_t c = 0; c < done; c++) { t--; if (r.h != r.l) r = orc_constrain_range(b, q[t], r); }
        if (pair) {
            while (t >= 2 && r.h != r.l) {
                ps++;
                if (r.l / line_syms != r.h / line_syms) p2++;
                r = orc_constrain_range(b, q[t - 1], r);
                if (r.h != r.l) r = orc_constrain_range(b, q[t - 2], r);
                t -= 2;
            }
        } else {
            while (t >= 1 && r.h != r.l) {
                os++;
                if ((r.l >> block_shift) != (r.h >> block_shift)) o2++;
                r = orc_constrain_range(b, q[t - 1], r);
                t--;
            }
        }
    }
    out[0] = ps; out[1] = p2; out[2] = os; out[3] = o2; out[4] = hits; out[5] = n;
    return ORC_OK;
}

/* Accounting replay of the engine's QUAD path (layout.h): an all-ACGT k-mer starts from the suffix table
 * at the deepest of the four levels table_s .. table_s-3 that leaves a multiple of four symbols, then
 * takes one quad step per four symbols -- one `sector_syms`-position sector per boundary, `line_sectors`
 * sectors per 128-byte line -- and finishes a remainder with one-step ranks; any other k-mer takes the
 * one-step path.  out[0..7] = quad steps, quad steps whose l and h fall in different sectors, ... in
 * different lines, one-step steps, one-step steps over two blocks, table lookups, queries, 0. */
int orc_count_kmers_stats_quad(const orc_rle_bwt *b, const uint8_t *syms, uint32_t k, uint64_t n,
                               uint32_t table_s, uint32_t sector_syms, uint32_t line_sectors,
                               unsigned block_shift, uint64_t *out) {
    return orc_count_kmers_stats_oct(b, syms, k, n, table_s, sector_syms, line_sectors, block_shift, 0, 0, out);
}

/* The same with an OCT image on top (oct_bucket_shift != 0; oct_syms symbols per line, 8 or 10): the table depth
 * is the one of the four kept levels that leaves the cheapest walk (kernel_common.cuh: one access per oct or
 * quad step, two per one-symbol step); while oct_syms or more symbols are left a step reads one line of the
 * (code, 2^oct_bucket_shift-position bucket) it needs -- when l and h fall in different buckets the same
 * symbols are taken as quad steps (and one-symbol steps for what four does not divide).  out[7] = oct steps,
 * out[8] = two-bucket events.  (Lines the engine answers without the oct image because they overflowed are not
 * modelled: msbwt_oct_overflow_lines reports how many exist.) */
static uint32_t oct_walk_cost(uint32_t rest, uint32_t m) { const uint32_t r = rest % m; return rest / m + r / 4u + 2u * (r % 4u); }

/* fin_bucket_shift != 0: with the (experimental) final-step image on top of the oct image -- exactly fin_syms
 * symbols left and l, h in one bucket of 2^fin_bucket_shift positions: ONE line answers the count and the query
 * ends (a range over two buckets takes the oct steps; overflowed lines are not modelled, as for the oct image).
 * fin[0] = final steps, fin[1] = two-bucket events. */
static int stats_oct_fin(const orc_rle_bwt *b, const uint8_t *syms, uint32_t k, uint64_t n,
                         uint32_t table_s, uint32_t sector_syms, uint32_t line_sectors,
                         unsigned block_shift, unsigned oct_bucket_shift, unsigned oct_syms,
                         unsigned fin_bucket_shift, unsigned fin_syms, uint64_t *out, uint64_t *fin) {
    uint64_t qs = 0, q2s = 0, q2l = 0, os = 0, o2 = 0, hits = 0, es = 0, e2 = 0, fs = 0, f2 = 0;
    const uint64_t line_syms = (uint64_t)sector_syms * line_sectors;
    const uint32_t m = oct_bucket_shift ? (oct_syms ? oct_syms : 8u) : 0u;
    for (uint64_t i = 0; i < n; i++) {
        const uint8_t *q = syms + i * (uint64_t)k;
        int all_acgt = 1;
        uint32_t na = 0;
        for (uint32_t t = 0; t < k; t++) {
            const uint8_t sy = q[k - 1 - t];
            if (sy >= ORC_VC_LEN) return ORC_PANIC_BAD_SYMBOL;
            const int ok = sy == 1 || sy == 2 || sy == 3 || sy == 5;
            all_acgt &= ok;
            if (t < table_s && na == t && ok) na++;
        }
        uint32_t done = 0;
        if (all_acgt) {
            if (table_s && k >= table_s) {
                if (m) {
                    uint32_t best_cost = oct_walk_cost(k, m);
                    for (uint32_t back = 0; back < 4u && back < table_s; back++) {
                        const uint32_t c = oct_walk_cost(k - (table_s - back), m);
                        if (c < best_cost) { best_cost = c; done = table_s - back; }
                    }
                } else {
                    const uint32_t back = (4u - (k - table_s) % 4u) % 4u;
                    done = back < table_s ? table_s - back : 0;
                }
            } else if (table_s && k + 4 > table_s) {
                done = k;
            }
        } else if (table_s && na >= table_s) {
            done = table_s;
        }
        hits += done != 0;
        orc_range r = { 0, b->total_size };
        uint32_t t = k, forced = 0;
        for (uint32_t c = 0; c < done; c++) { t--; if (r.h != r.l) r = orc_constrain_range(b, q[t], r); }
        int fin_tried = 0;
        while (t >= 1 && r.h != r.l) {
            if (all_acgt && m && fin_bucket_shift && t == fin_syms && forced == 0 && !fin_tried) {
                if ((r.l >> fin_bucket_shift) == (r.h >> fin_bucket_shift)) {
                    fs++;
                    for (uint32_t u = 0; u < fin_syms; u++) { t--; if (r.h != r.l) r = orc_constrain_range(b, q[t], r); }
                    continue;
                }
                f2++;
                fin_tried = 1;
            }
            if (all_acgt && m && t >= m && forced == 0) {
                if ((r.l >> oct_bucket_shift) == (r.h >> oct_bucket_shift)) {
                    es++;
                    for (uint32_t u = 0; u < m; u++) { t--; if (r.h != r.l) r = orc_constrain_range(b, q[t], r); }
                    continue;
                }
                e2++;
                forced = m;
            }
            if (all_acgt && t >= 4 && (forced == 0 || forced >= 4)) {
                qs++;
                if (r.l / sector_syms != r.h / sector_syms) q2s++;
                if (r.l / line_syms != r.h / line_syms) q2l++;
                for (int u = 0; u < 4; u++) { t--; if (r.h != r.l) r = orc_constrain_range(b, q[t], r); }
                forced = forced >= 4 ? forced - 4 : 0;
                continue;
            }
            os++;
            if ((r.l >> block_shift) != (r.h >> block_shift)) o2++;
            r = orc_constrain_range(b, q[t - 1], r);
            t--;
            if (forced) forced--;
        }
    }
    out[0] = qs; out[1] = q2s; out[2] = q2l; out[3] = os; out[4] = o2; out[5] = hits; out[6] = n;
    if (oct_bucket_shift) { out[7] = es; out[8] = e2; } else { out[7] = 0; }
    if (fin) { fin[0] = fs; fin[1] = f2; }
    return ORC_OK;
}

int orc_count_kmers_stats_oct(const orc_rle_bwt *b, const uint8_t *syms, uint32_t k, uint64_t n,
                              uint32_t table_s, uint32_t sector_syms, uint32_t line_sectors,
                              unsigned block_shift, unsigned oct_bucket_shift, unsigned oct_syms, uint64_t *out) {
    return stats_oct_fin(b, syms, k, n, table_s, sector_syms, line_sectors, block_shift, oct_bucket_shift, oct_syms, 0, 0, out, NULL);
}

int orc_count_kmers_stats_fin(const orc_rle_bwt *b, const uint8_t *syms, uint32_t k, uint64_t n,
                              uint32_t table_s, uint32_t sector_syms, uint32_t line_sectors,
                              unsigned block_shift, unsigned oct_bucket_shift, unsigned oct_syms,
                              unsigned fin_bucket_shift, unsigned fin_syms, uint64_t *out /* 9 */, uint64_t *fin /* 2 */) {
    return stats_oct_fin(b, syms, k, n, table_s, sector_syms, line_sectors, block_shift, oct_bucket_shift, oct_syms,
                         fin_bucket_shift, fin_syms, out, fin);
}

/* ---- bwt_converter.rs ---- */
static inline uint64_t emit_run(uint8_t sym, uint64_t count, uint8_t *out, uint64_t cap, uint64_t at) {
    /* little-endian base-32 digits, one per byte, zero digits kept (:52-56, :166-171) */
    while (count > 0) {
        if (out && at < cap) out[at] = (uint8_t)(sym | ((count & ORC_COUNT_MASK) << ORC_LETTER_BITS));
        at++;
        count >>= ORC_NUMBER_BITS;
    }
    return at;
}

uint64_t orc_convert_to_vec(const uint8_t *ascii, uint64_t n, uint8_t *out, uint64_t cap) {
    uint8_t translate[256];
    memset(translate, 255, sizeof(translate));
    const char *valid = "$ACGNT";
    for (int x = 0; x < 6; x++) translate[(uint8_t)valid[x]] = (uint8_t)x;
    uint64_t at = 0, count = 0;
    uint8_t curr = '$'; /* any valid symbol: its count is 0 (:35) */
    for (uint64_t i = 0; i < n; i++) {
        uint8_t ch = ascii[i];
        if (ch == curr) count++;
        else if (translate[ch] == 255) { if (ch != 10) return (uint64_t)-1; } /* :41-46 */
        else { at = emit_run(translate[curr], count, out, cap, at); curr = ch; count = 1; }
    }
    return emit_run(translate[curr], count, out, cap, at);
}

uint64_t orc_encode_runs(const uint8_t *syms, const uint64_t *counts, uint64_t nruns, uint8_t *out, uint64_t cap) {
    uint64_t at = 0;
    for (uint64_t i = 0; i < nruns; i++) at = emit_run(syms[i], counts[i], out, cap, at);
    return at;
}

int orc_save_bwt_numpy(const uint8_t *rle, uint64_t len, const char *path) {
    FILE *f = fopen(path, "wb");
    if (!f) return ORC_ERR_IO;
    char hdr[96];
    memset(hdr, 32, 95); hdr[95] = 10;                                 /* :107-108 */
    static const char head[] = "\x93NUMPY\x01\x00\x56\x00{'descr': '|u1', 'fortran_order': False, 'shape': (";
    size_t hl = sizeof(head) - 1;
    memcpy(hdr, head, hl);
    char num[32];
    int nl = snprintf(num, sizeof(num), "%llu", (unsigned long long)len);
    memcpy(hdr + hl, num, (size_t)nl);
    memcpy(hdr + hl + nl, ", ), }", 6);                                /* :120 */
    int ok = fwrite(hdr, 1, 96, f) == 96 && (len == 0 || fwrite(rle, 1, len, f) == len);
    fclose(f);
    return ok ? ORC_OK : ORC_ERR_IO;
}

/* ---- string_util.rs ---- */
const uint8_t ORC_INT_TO_STRING[6] = { '$', 'A', 'C', 'G', 'N', 'T' };
const uint8_t ORC_COMPLEMENT_INT[6] = { 0, 5, 3, 2, 4, 1 };

uint8_t orc_string_to_int(uint8_t a) {
    switch (a) {
        case '$': return 0;
        case 'A': case 'a': return 1;
        case 'C': case 'c': return 2;
        case 'G': case 'g': return 3;
        case 'T': case 't': return 5;
        default: return 4; /* N, n and everything unknown (:16) */
    }
}
void orc_convert_stoi(const uint8_t *ascii, uint64_t n, uint8_t *out) { for (uint64_t i = 0; i < n; i++) out[i] = orc_string_to_int(ascii[i]); }
void orc_convert_itos(const uint8_t *syms, uint64_t n, uint8_t *out) { for (uint64_t i = 0; i < n; i++) out[i] = ORC_INT_TO_STRING[syms[i]]; }
void orc_reverse_complement_i(const uint8_t *syms, uint64_t n, uint8_t *out) { for (uint64_t i = 0; i < n; i++) out[i] = ORC_COMPLEMENT_INT[syms[n - 1 - i]]; }
