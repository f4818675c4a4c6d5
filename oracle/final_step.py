"""TEST INFRASTRUCTURE ONLY (like everything under oracle/): numpy specification of the FINAL-STEP image planned in
DESIGN.md section 7, item 4 -- written before the kernels so that the device builder and the search kernel's final
step have something to be compared with line by line.  Nothing in the product imports this.

Identity (follows from BWT::count_kmer, src/msbwt_core.rs:125-161 of the reference, being k dependent
RleBWT::constrain_range calls, src/rle_bwt.rs:202-287): let [l, h) be the range after the k-mer's last k - m symbols
and c the code of its first m symbols; then count_kmer = #{ j in [l, h) : code_m(j) = c } with
code_m(j) = (B[j], B[LF j], .., B[LF^(m-1) j]), the m text symbols that precede suffix j (undefined when one of them
is `$` or `N`).  The last step therefore needs no checkpoint and no entry for a code that does not occur.

Image: 128-byte lines of 32 u32 words, `1 << lb` lines per bucket of `1 << b` positions (b <= 16, lb >= 12 so that
a tag fits 28 bits for m = 20).  A code goes to line `(bucket << lb) | (mix40(c) & (2^lb - 1))` with tag
`mix40(c) >> lb`, mix40 a bijection of the 2m-bit codes.
  word 0      : number of words in use after it (0..31), or 0xFFFFFFFF when the line's groups do not fit (the query
                then takes its m symbols through the other images, as it does when [l, h) spans two buckets)
  then groups : header `(tag << 4) | nruns` (nruns 1..15) followed by nruns words `(len << 16) | offset` -- a run of
                `len` consecutive positions with that code starting at `offset` inside the bucket.  A code with more
                than 15 runs in a bucket simply has several groups; a run never crosses a bucket boundary.
The groups of one code are adjacent (both builders write a line's groups in (tag, start) order); the order of different
codes inside a line is otherwise unspecified: compare lines as sets of groups.
"""
from __future__ import annotations

import numpy as np

M_SYMS = 20
LINE_WORDS = 32
OVERFLOW = 0xFFFFFFFF
_CODE = np.full(256, -1, dtype=np.int64)
_CODE[[1, 2, 3, 5]] = [0, 1, 2, 3]
_MASK40 = (1 << (2 * M_SYMS)) - 1
_MUL1, _MUL2 = 0x9E3779B97F, 0xC2B2AE3D27  # odd: multiplication mod 2^40 is a bijection


def mix40(c):
    """bijection on [0, 2^40): xorshift by half the width and odd multiplications, each invertible"""
    c = np.asarray(c, dtype=np.uint64) & np.uint64(_MASK40)
    with np.errstate(over="ignore"):  # the products wrap modulo 2^64 on purpose; only their low 40 bits are kept
        c = c ^ (c >> np.uint64(20))
        c = (c * np.uint64(_MUL1)) & np.uint64(_MASK40)
        c = c ^ (c >> np.uint64(20))
        c = (c * np.uint64(_MUL2)) & np.uint64(_MASK40)
        c = c ^ (c >> np.uint64(20))
    return c


def lf_mapping(bwt: np.ndarray) -> np.ndarray:
    """LF[j] = C[B[j]] + rank(B[j], j): one constrain_range boundary per position (src/rle_bwt.rs:202-287)"""
    order = np.argsort(bwt, kind="stable")
    lf = np.empty(bwt.size, dtype=np.int64)
    lf[order] = np.arange(bwt.size, dtype=np.int64)
    return lf


def position_codes(bwt: np.ndarray, m: int = M_SYMS) -> np.ndarray:
    """code_m(j) per position, -1 where one of the m symbols is not A, C, G or T; symbol t of the walk (t = 0: B[j],
    the first one the search consumes) at bits 2 (m-1-t) -- first consumed most significant, the order in which the
    search kernels read a step's symbols off the query word, so that code_20(j) = code_10(j) << 20 | code_10(LF^10 j)"""
    lf = lf_mapping(bwt)
    code = np.zeros(bwt.size, dtype=np.int64)
    ok = np.ones(bwt.size, dtype=bool)
    cur = np.arange(bwt.size, dtype=np.int64)
    for t in range(m):
        c2 = _CODE[bwt[cur]]
        ok &= c2 >= 0
        code |= np.where(c2 >= 0, c2, 0) << (2 * (m - 1 - t))
        cur = lf[cur]
    return np.where(ok, code, -1)


def query_code(kmer_prefix: np.ndarray) -> int:
    """code of the FIRST m symbols of a k-mer as the search consumes them: the last of them first, most significant"""
    c = 0
    m = len(kmer_prefix)
    for t in range(m):
        c |= int(_CODE[kmer_prefix[m - 1 - t]]) << (2 * (m - 1 - t))
    return c


def build_final_image(bwt: np.ndarray, b: int = 16, lb: int = 12, m: int = M_SYMS):
    """-> (lines [nbuck << lb, 32] u32, stats).  Runs are cut at bucket boundaries and at 65535 positions."""
    assert b <= 16 and 2 * m - lb <= 28
    n = bwt.size
    key = position_codes(bwt, m)
    nbuck = (n >> b) + 1
    lines = np.zeros((nbuck << lb, LINE_WORDS), dtype=np.uint32)
    j = np.arange(n, dtype=np.int64)
    head = key >= 0
    head[1:] &= (key[1:] != key[:-1]) | ((j[1:] & ((1 << b) - 1)) == 0)
    starts = np.flatnonzero(head)
    # run end: next position whose key differs, or the bucket boundary
    change = np.ones(n + 1, dtype=bool)
    change[1:n] = (key[1:] != key[:-1]) | ((j[1:] & ((1 << b) - 1)) == 0)
    bounds = np.flatnonzero(change)
    ends = bounds[np.searchsorted(bounds, starts, side="right")]
    codes = key[starts].astype(np.uint64)
    mixed = mix40(codes)
    line_of = ((starts >> b) << lb) | (mixed & np.uint64((1 << lb) - 1)).astype(np.int64)
    tags = (mixed >> np.uint64(lb)).astype(np.int64)
    order = np.lexsort((starts, tags, line_of))
    overflowed = 0
    groups = 0
    i = 0
    while i < order.size:
        ln = line_of[order[i]]
        e = i
        while e < order.size and line_of[order[e]] == ln:
            e += 1
        words = []
        g = i
        while g < e:
            tag = tags[order[g]]
            ge = g
            runs = []
            while ge < e and tags[order[ge]] == tag:
                s, t_end = int(starts[order[ge]]), int(ends[order[ge]])
                while s < t_end:                      # runs longer than 65535 are cut
                    ln_run = min(65535, t_end - s)
                    runs.append((ln_run << 16) | (s & ((1 << b) - 1)))
                    s += ln_run
                ge += 1
            for r0 in range(0, len(runs), 15):
                part = runs[r0:r0 + 15]
                words.append((int(tag) << 4) | len(part))
                words.extend(part)
                groups += 1
            g = ge
        if len(words) > LINE_WORDS - 1:
            lines[ln, 0] = OVERFLOW
            overflowed += 1
        else:
            lines[ln, 0] = len(words)
            lines[ln, 1:1 + len(words)] = np.array(words, dtype=np.uint32)
        i = e
    stats = {"positions": int(n), "coded_positions": int((key >= 0).sum()), "runs": int(starts.size), "groups": groups,
             "lines": int(lines.shape[0]), "overflowed_lines": overflowed, "bucket_shift": b, "lines_per_bucket_log2": lb}
    return lines, stats


def line_groups(line: np.ndarray):
    """one line as a sorted list of (tag, sorted run words) with the groups of one tag merged -- the order-free view
    two builders must share (how a code's runs are split over groups of <= 15 is the builder's business)"""
    if int(line[0]) == OVERFLOW:
        return None
    by_tag, i, used = {}, 1, int(line[0])
    while i <= used:
        tag, nr = int(line[i]) >> 4, int(line[i]) & 15
        by_tag.setdefault(tag, []).extend(int(w) for w in line[i + 1:i + 1 + nr])
        i += 1 + nr
    return sorted((t, tuple(sorted(r))) for t, r in by_tag.items())


def final_count(lines: np.ndarray, code: int, l: int, h: int, b: int = 16, lb: int = 12):
    """#{ j in [l, h) : code_m(j) = code } from ONE line, or None when the caller must fall back (two buckets /
    overflowed line).  What the search kernel's final step computes from its one staged line."""
    if l >= h:
        return 0
    if (l >> b) != (h >> b):          # the kernel's test (h exclusive: a range ending on a boundary falls back too)
        return None
    mixed = int(mix40(np.uint64(code)))
    line = lines[((l >> b) << lb) | (mixed & ((1 << lb) - 1))]
    if int(line[0]) == OVERFLOW:
        return None
    tag, pl, ph = mixed >> lb, l & ((1 << b) - 1), h & ((1 << b) - 1)
    total, i, used = 0, 1, int(line[0])
    while i <= used:
        gt, nr = int(line[i]) >> 4, int(line[i]) & 15
        if gt == tag:
            for w in line[i + 1:i + 1 + nr]:
                off, ln = int(w) & 0xFFFF, int(w) >> 16
                total += min(max(ph - off, 0), ln) - min(max(pl - off, 0), ln)
        i += 1 + nr
    return total
