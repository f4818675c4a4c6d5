"""north_star item 3 / SURVEY K5 (the reference's planned kmer_cache, src/msbwt_core.rs:133-146): does sorting the
batch and collapsing equal k-mers pay?  Measured on the full configs[2] batch (100 M read-sampled 31-mers over the
1.51 Gsymbol BWT) with device-resident packed keys: radix sort of the 62-bit keys with their indices (torch.sort on a
1-D int64 CUDA tensor is cub::DeviceRadixSort::SortPairs), adjacent-duplicate collapse, the search on the unique set,
the scatter back through the inverse permutation -- against the search on the batch as it is.  Counts must be equal.

    python tools/group_bench.py [--workload cfg3] > profiles/r2_group_cfg3.json
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="cfg3")
    ap.add_argument("--iters", type=int, default=5)
    args = ap.parse_args()
    import torch

    import bench
    import rust_msbwt_b200 as M

    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    cfg = dict(bench.WORKLOADS[args.workload])
    k = cfg["k"]
    rle_host, total, queries, _ = bench.build_workload(cfg, dev, 0)
    n = queries.shape[0]
    keys = bench.encode_u64(queries, k)
    del queries
    bwt = M.RleBWT.new(devices=[0])
    bwt.load_vector(rle_host)
    st = torch.cuda.current_stream().cuda_stream
    d_packed = torch.empty(bwt.packed_bytes(k, n) // 8, dtype=torch.int64, device=dev)
    d_out = torch.empty(n, dtype=torch.int64, device=dev)

    def timed(fn):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r = fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1), r

    def search(kk, out):
        m = kk.shape[0]
        bwt.seed_kmers_u64_device(kk.data_ptr(), k, m, d_packed.data_ptr(), out.data_ptr(), st)
        bwt.count_kmers_packed_device(d_packed.data_ptr(), k, m, out.data_ptr(), st)

    plain, sort_ms, uniq_ms, search_u, scatter_ms = [], [], [], [], []
    uniq_n = 0
    for it in range(args.iters + 1):
        t, _ = timed(lambda: search(keys, d_out))
        t_sort, (skeys, perm) = timed(lambda: torch.sort(keys))
        t_uniq, (ukeys, inverse) = timed(lambda: torch.unique_consecutive(skeys, return_inverse=True))
        uniq_n = int(ukeys.shape[0])
        d_out_u = torch.empty(uniq_n, dtype=torch.int64, device=dev)
        t_su, _ = timed(lambda: search(ukeys, d_out_u))
        def scatter():
            res = torch.empty(n, dtype=torch.int64, device=dev)
            res[perm] = d_out_u[inverse]
            return res
        t_sc, grouped = timed(scatter)
        assert (grouped == d_out).all(), "grouped and ungrouped counts differ"
        if it:
            plain.append(t); sort_ms.append(t_sort); uniq_ms.append(t_uniq); search_u.append(t_su); scatter_ms.append(t_sc)
        del skeys, perm, ukeys, inverse, d_out_u, grouped
    med = statistics.median
    print(json.dumps({
        "workload": cfg["name"], "queries": n, "k": k, "bwt_symbols": int(total),
        "unique_kmers": uniq_n, "unique_fraction": uniq_n / n,
        "ungrouped_seed_plus_search_ms": med(plain),
        "sort_pairs_ms": med(sort_ms), "collapse_ms": med(uniq_ms), "search_unique_set_ms": med(search_u),
        "scatter_back_ms": med(scatter_ms),
        "grouped_total_ms": med(sort_ms) + med(uniq_ms) + med(search_u) + med(scatter_ms),
        "verdict": "grouping pays" if med(sort_ms) + med(uniq_ms) + med(search_u) + med(scatter_ms) < med(plain) else "grouping does not pay: sorting the batch costs more than the duplicate work it removes",
        "parity": "grouped counts == ungrouped counts on every query",
    }), flush=True)


if __name__ == "__main__":
    main()
