"""A/B of the pack/seed kernel (K0) between builds of the library: one process per build, selected with
MSBWT_LIBRARY_PATH (tools/build_variant.sh).  Times the pack/seed launch and the search launch separately with
CUDA events on the launching stream (L2 flushed between iterations) and prints one JSON line with the checksum
of the counts, so that two builds can be compared for speed AND for equality of their results.

    MSBWT_LIBRARY_PATH=build/variants/lib_x.so python tools/pack_ab.py [--workload cfg2] [--iters 20] [--k 31]
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
    ap.add_argument("--workload", default="cfg2")
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--k", type=int, default=0, help="query length (default: the workload's)")
    ap.add_argument("--n", type=int, default=0, help="use only the first n queries of the batch (a rank's slice under strong scaling)")
    args = ap.parse_args()

    import torch

    import bench
    import rust_msbwt_b200 as M

    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    cfg = dict(bench.WORKLOADS[args.workload])
    if args.k:
        cfg["k"] = args.k
    k = cfg["k"]
    rle_host, total, queries, _ = bench.build_workload(cfg, dev, 0)
    if args.n:
        queries = queries[: args.n].contiguous()
    n = queries.shape[0]
    bwt = M.RleBWT.new(devices=[0])
    bwt.load_vector(rle_host)
    stream = torch.cuda.current_stream().cuda_stream
    d_packed = torch.empty(bwt.packed_bytes(k, n) // 8, dtype=torch.int64, device=dev)
    d_out = torch.empty(n, dtype=torch.int64, device=dev)
    d_status = torch.zeros(1, dtype=torch.int32, device=dev)
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)

    pack_ms, search_ms = [], []
    for it in range(args.iters + 3):
        flush.fill_(it & 1)
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        e[0].record()
        bwt.pack_kmers_device(queries.data_ptr(), k, n, d_packed.data_ptr(), d_out.data_ptr(), d_status.data_ptr(), stream)
        e[1].record()
        bwt.count_kmers_packed_device(d_packed.data_ptr(), k, n, d_out.data_ptr(), stream)
        e[2].record()
        torch.cuda.synchronize()
        if it >= 3:
            pack_ms.append(e[0].elapsed_time(e[1]))
            search_ms.append(e[1].elapsed_time(e[2]))
    assert int(d_status.item()) == 0
    print(json.dumps({
        "library": os.environ.get("MSBWT_LIBRARY_PATH", "in-tree"), "workload": args.workload, "k": k, "queries": n,
        "bwt_symbols": int(total), "oct_index": bool(bwt.oct_index), "final_index": bool(bwt.final_index), "suffix_table_s": bwt.suffix_table_s,
        "fixed_k_pack": os.environ.get("MSBWT_PACK_FIXED_K", ""),
        "pack_ms_median": statistics.median(pack_ms), "pack_ms_min": min(pack_ms),
        "search_ms_median": statistics.median(search_ms), "search_ms_min": min(search_ms),
        "checksum": int(d_out.sum().item()), "present": int((d_out > 0).sum().item()),
    }), flush=True)


if __name__ == "__main__":
    main()
