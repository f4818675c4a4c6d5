"""CPU model of the FINAL-STEP image planned in DESIGN.md section 7 (item 4): numbers before kernels.

The last m symbols a `count_kmer` (src/msbwt_core.rs:125-161) consumes need no rank, only a count: with [l, h)
the range after the k-mer's last k - m symbols, the answer is the number of positions j in [l, h) whose m-symbol
code (B[j], B[LF j], .., B[LF^(m-1) j]) -- the m text symbols that precede suffix j -- spells the rest of the
k-mer.  This script builds that on a synthetic read set in numpy and reports

  * that the identity holds (model count == plain backward search, every sampled query),
  * run statistics of the m-symbol codes (one image entry per run of consecutive positions with one code),
  * the load of a hashed image: lines of `slots` entries, `lines_per_bucket` lines per 2^b positions, the share
    of entries that land on an overflowed line, and the share of sampled queries that would fall back
    (range over two buckets, or an overflowed line),
  * bytes per BWT symbol.

    python tools/final_step_model.py [--reads 200000] [--m 20] [--table 11] [--k 31]
"""
from __future__ import annotations

import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CODE = np.full(6, -1, dtype=np.int64)
CODE[[1, 2, 3, 5]] = [0, 1, 2, 3]


def mix(c: np.ndarray) -> np.ndarray:
    """invertible 64-bit mix (splitmix64 finaliser): the line index takes its low bits, the tag the rest"""
    c = c.astype(np.uint64)
    c ^= c >> np.uint64(30)
    c *= np.uint64(0xBF58476D1CE4E5B9)
    c ^= c >> np.uint64(27)
    c *= np.uint64(0x94D049BB133111EB)
    c ^= c >> np.uint64(31)
    return c


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reads", type=int, default=200_000)
    ap.add_argument("--read-len", type=int, default=150)
    ap.add_argument("--error", type=float, default=0.01)
    ap.add_argument("--m", type=int, default=20, help="symbols answered by the final step")
    ap.add_argument("--table", type=int, default=11, help="symbols answered by the suffix table before it")
    ap.add_argument("--k", type=int, default=31)
    ap.add_argument("--queries", type=int, default=20000)
    ap.add_argument("--slots", type=int, default=16, help="8-byte entries per 128-byte line")
    args = ap.parse_args()
    assert args.table + args.m <= args.k and (args.k - args.table - args.m) == 0, "model covers table + one final step"

    from harness import bwt_build, synth

    reads = synth.make_reads(args.reads, args.read_len, 30.0, args.error, device="cpu")
    B = bwt_build.build_msbwt(reads).numpy().astype(np.uint8)
    N = B.size
    order = np.argsort(B, kind="stable")
    LF = np.empty(N, dtype=np.int64)
    LF[order] = np.arange(N, dtype=np.int64)
    counts = np.bincount(B, minlength=6)
    C = np.concatenate([[0], np.cumsum(counts)[:-1]])
    pos_of = {s: np.flatnonzero(B == s) for s in (1, 2, 3, 5)}

    def rank(sym: np.ndarray, p: np.ndarray) -> np.ndarray:
        out = np.zeros(p.size, dtype=np.int64)
        for s in (1, 2, 3, 5):
            sel = sym == s
            if sel.any():
                out[sel] = np.searchsorted(pos_of[s], p[sel], side="left")
        return out

    # ---- m-symbol code of every position (invalid when a `$` / `N` is among the m symbols)
    code = np.zeros(N, dtype=np.int64)
    ok = np.ones(N, dtype=bool)
    cur = np.arange(N, dtype=np.int64)
    for t in range(args.m):
        c2 = CODE[B[cur]]
        ok &= c2 >= 0
        code |= np.where(c2 >= 0, c2, 0) << (2 * t)
        cur = LF[cur]
    key = np.where(ok, code, -1)
    head = np.ones(N, dtype=bool)
    head[1:] = key[1:] != key[:-1]
    head &= ok
    head |= np.concatenate([[False], ok[1:] & ~ok[:-1]])
    run_start = np.flatnonzero(head)
    run_code = code[run_start]
    n_runs = run_start.size
    res = {"bwt_symbols": int(N), "m": args.m, "positions_with_a_code": int(ok.sum()), "runs": int(n_runs),
           "mean_run": float(ok.sum() / max(1, n_runs)), "images": []}

    # ---- sampled read k-mers: [l, h) after the table, the plain count, the model count
    q = synth.make_queries(reads, args.k, args.queries, 0).numpy()
    q = q[np.isin(q, (1, 2, 3, 5)).all(axis=1)]
    l = np.zeros(q.shape[0], dtype=np.int64)
    h = np.full(q.shape[0], N, dtype=np.int64)
    lt = ht = None
    for t in range(args.k):
        sym = q[:, args.k - 1 - t]
        l, h = C[sym] + rank(sym, l), C[sym] + rank(sym, h)
        if t + 1 == args.table:
            lt, ht = l.copy(), h.copy()
    want = h - l
    qcode = np.zeros(q.shape[0], dtype=np.int64)
    for t in range(args.m):                                   # step t of the final stretch consumes symbol k-1-table-t
        qcode |= CODE[q[:, args.k - 1 - args.table - t]] << (2 * t)
    got = np.array([int(((key[a:b] == c)).sum()) for a, b, c in zip(lt, ht, qcode)], dtype=np.int64)
    res["queries"] = int(q.shape[0])
    res["identity_holds"] = bool((got == want).all())
    res["mean_range_after_table"] = float((ht - lt).mean())

    # ---- hashed image: load and fallback shares.  Two line formats:
    #   flat    : `slots` entries {tag, offset, length} of 8 bytes, one per run
    #   grouped : per code present in the line a header word {tag, run count} + one word {length, offset} per run,
    #             31 words of payload (oracle/final_step.py; runs of one code share the tag)
    hm = mix(run_code)
    for b in (14, 16):
        bucket = run_start >> b
        nb = int(N >> b) + 1
        per_bucket = n_runs / nb
        # groups: distinct (bucket, code) pairs
        gkey = bucket.astype(np.uint64) * np.uint64(1 << 40) + run_code.astype(np.uint64)
        gk, ginv, gruns = np.unique(gkey, return_inverse=True, return_counts=True)
        gbucket = (gk >> np.uint64(40)).astype(np.int64)
        gmix = mix(gk & np.uint64((1 << 40) - 1))
        for fill in (0.5, 0.33, 0.25):
            lines = 1
            while lines * args.slots * fill < per_bucket:
                lines *= 2
            line = bucket * lines + (hm & np.uint64(lines - 1)).astype(np.int64)
            load = np.bincount(line, minlength=nb * lines)
            over = load > args.slots
            gline = gbucket * lines + (gmix & np.uint64(lines - 1)).astype(np.int64)
            # the specified format (oracle/final_step.py): one header word per <= 15 runs of a code + one word per run
            gbytes = np.bincount(gline, weights=4 * (-(-gruns // 15)) + 4 * gruns, minlength=nb * lines)
            gover = gbytes > 124
            qline = (lt >> b) * lines + (mix(qcode) & np.uint64(lines - 1)).astype(np.int64)
            two = (lt >> b) != ((np.maximum(ht, lt + 1) - 1) >> b)
            res["images"].append({
                "bucket_shift": b, "lines_per_bucket": lines, "fill_target": fill,
                "bytes_per_symbol": float(nb * lines * 128 / N),
                "mean_runs_per_line": float(load.mean()), "mean_runs_per_code_and_bucket": float(gruns.mean()),
                "flat_overflowed_lines_share": float(over.mean()),
                "flat_queries_falling_back_share": float((two | over[qline]).mean()),
                "grouped_mean_bytes_per_line": float(gbytes.mean()),
                "grouped_overflowed_lines_share": float(gover.mean()),
                "grouped_queries_falling_back_share": float((two | gover[qline]).mean()),
                "queries_over_two_buckets_share": float(two.mean()),
                "tag_bits": int(2 * args.m - np.log2(lines)), "offset_bits": b,
            })
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
