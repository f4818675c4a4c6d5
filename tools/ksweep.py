"""BASELINE.json configs[3]: k-sweep (k = 15 / 31 / 63 / 101) over the 1.51 Gsymbol BWT of configs[2],
10 M read-sampled queries per k, kernel-only (queries resident in HBM), each k checked against the CPU
oracle on a sample, with the index traffic of every k counted by the counting build of the search kernel.
(What grouping equal k-mers of a batch would buy is measured on the full 100 M batch by tools/group_bench.py:
profiles/r2_group_cfg3.json.)

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
           "quad_index": bwt.quad_index, "pair_index": bwt.pair_index, "oct_index": bwt.oct_index, "oct_bucket_shift": bwt.oct_bucket_shift,
           "final_index": bwt.final_index, "index_bytes": bwt.index_bytes,
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
        # what the two stages fetch: the pack stage's own counters + the counting build of the search kernel
        d_packed = torch.empty(bwt.packed_bytes(k, n) // 8, dtype=torch.int64, device=dev)
        d_status = torch.zeros(1, dtype=torch.int32, device=dev)
        d_stats = torch.zeros(8, dtype=torch.int64, device=dev)
        d_out2 = torch.zeros(n, dtype=torch.int64, device=dev)
        ms_pack = timed(lambda: bwt.pack_kmers_device(q.data_ptr(), k, n, d_packed.data_ptr(), d_out2.data_ptr(), d_status.data_ptr(), stream))
        ps = bwt.pack_stats(d_packed.data_ptr(), k, n)
        bwt.count_kmers_packed_stats_device(d_packed.data_ptr(), k, n, d_out2.data_ptr(), d_stats.data_ptr(), stream)
        torch.cuda.synchronize()
        assert (d_out2 == d_out).all()
        st = [int(v) for v in d_stats.cpu().tolist()]
        lines = ps["final_lines"] + st[0] + st[1] + st[4]
        out["results"].append({
            "k": k, "queries": n, "suffix_table_depth_used": bwt.table_depth_for_k(k), "ms": ms, "ms_pack_stage": ms_pack,
            "queries_per_s": n / (ms / 1e3),
            "reference_steps_after_table": steps / m,
            "line_fills_per_query": lines / n, "one_step_blocks_per_query": st[6] / n,
            "pack_stage": ps, "search_oct_lines": st[0], "search_final_lines": st[1], "search_quad_lines": st[4],
            "ns_per_query": 1e6 * ms / n, "line_fills_per_s": lines / (ms / 1e3),
            "parity_checked": m})
        del d_packed, d_status, d_stats, d_out2
        print(out["results"][-1], file=sys.stderr)
        del q, d_out
    print(json.dumps(out))


if __name__ == "__main__":
    main()
