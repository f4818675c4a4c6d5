"""BASELINE.json configs[3]: k-sweep (k = 15 / 31 / 63 / 101) over the 1.51 Gsymbol BWT of configs[2],
10 M read-sampled queries per k, kernel-only (queries resident in HBM), each k checked against the CPU
oracle on a sample.  Also measures what exact-duplicate grouping of a batch would buy: the batch is
sorted + uniqued with torch on the device (harness-level stand-in for the planned native co-lex
sort/group stage), the engine runs on the unique k-mers, and the time of both variants is reported
(sort time included and shown separately).

    python tools/ksweep.py [--reads 10000000] [--queries 10000000] > gpurun_out/ksweep.json
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rust_msbwt_b200 as M  # noqa: E402
from harness import bwt_build, synth  # noqa: E402
from oracle import oracle as O  # noqa: E402


def timed(fn, reps=5):
    best = None
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        best = ms if best is None else min(best, ms)
    return best


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reads", type=int, default=10_000_000)
    ap.add_argument("--queries", type=int, default=10_000_000)
    ap.add_argument("--error", type=float, default=0.01)
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    torch.cuda.set_device(0)
    t0 = time.time()
    reads = synth.make_reads(args.reads, 150, 30.0, args.error, device=dev)
    rle, total = M.build_rle_bwt(reads.data_ptr(), 0, reads.shape[0], reads.shape[1])  # the library's own builder
    bwt = M.RleBWT.new(devices=[0])
    bwt.load_vector(rle)
    orc = O.RleBWT()
    orc.load_vector(rle)
    print(f"index: {total} symbols, {bwt.index_bytes / 1e6:.0f} MB, table s={bwt.suffix_table_s}, "
          f"lanes={bwt.kernel_lanes}, built in {time.time() - t0:.1f}s", file=sys.stderr)
    stream = torch.cuda.current_stream().cuda_stream
    out = {"bwt_symbols": total, "suffix_table_s": bwt.suffix_table_s, "kernel_lanes": bwt.kernel_lanes,
           "quad_index": bwt.quad_index, "pair_index": bwt.pair_index,
           "queries_per_k": args.queries, "results": []}
    for k in (15, 31, 63, 101):
        q = synth.make_queries(reads, k, args.queries, 0, seed_offset=k)
        n = q.shape[0]
        d_out = torch.zeros(n, dtype=torch.int64, device=dev)
        ms = timed(lambda: bwt.count_kmers_fixed_device(q.data_ptr(), k, n, d_out.data_ptr(), 0, stream))
        m = 100_000
        want = orc.count_kmers_fixed(q[:m].cpu().numpy(), k, threads=os.cpu_count() or 1)
        assert (d_out[:m].cpu().numpy().astype(np.uint64) == want).all(), f"parity failed at k={k}"
        steps, two, hits = orc.count_kmers_stats_skip(q[:m].cpu().numpy(), k, 7, bwt.suffix_table_s)
        # grouped variant: exact duplicates collapse (co-lex sort + group; torch stand-in)
        t_sort = timed(lambda: torch.unique(q, dim=0, return_inverse=True), reps=2)
        uq, inv = torch.unique(q, dim=0, return_inverse=True)
        nu = uq.shape[0]
        u_out = torch.zeros(nu, dtype=torch.int64, device=dev)
        ms_u = timed(lambda: bwt.count_kmers_fixed_device(uq.data_ptr(), k, nu, u_out.data_ptr(), 0, stream))
        assert (u_out[inv] == d_out).all()
        out["results"].append({
            "k": k, "queries": n, "ms_ungrouped": ms, "queries_per_s_ungrouped": n / (ms / 1e3),
            "mean_steps_after_table": steps / m, "two_block_share": two / max(1, steps),
            "ns_per_step_per_query_chain": 1e6 * ms / (steps / m) / 1.0 if steps else None,
            "unique_queries": nu, "ms_search_on_unique": ms_u, "ms_torch_sort_unique": t_sort,
            "queries_per_s_grouped_excl_sort": n / (ms_u / 1e3),
            "queries_per_s_grouped_incl_torch_sort": n / ((ms_u + t_sort) / 1e3),
            "parity_checked": m})
        print(out["results"][-1], file=sys.stderr)
        del q, d_out, uq, inv, u_out
    print(json.dumps(out))


if __name__ == "__main__":
    main()
