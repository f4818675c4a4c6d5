"""A/B of the pack/seed kernel (K0) between builds of the library: one process per build, selected with
MSBWT_LIBRARY_PATH (tools/build_variant.sh).  Times the pack/seed launch and the search launch separately with
CUDA events on the launching stream (L2 flushed between iterations) and prints one JSON line with the checksum
of the counts, so that two builds can be compared for speed AND for equality of their results.

    MSBWT_LIBRARY_PATH=build/variants/lib_x.so python tools/pack_ab.py [--workload cfg2] [--iters 20] [--k 31]

Also the stress / post-mortem harness of profiles/r2t_convergence.md:
    --also k:n,k:n      further (k, n) runs on the same index;  --superblock-shift s: 64-bit positions (the WIDE kernels)
    --postmortem T      prefill the outputs (--prefill v) before every iteration; when one does not end within T seconds,
                        read the outputs and the dispenser counter from ANOTHER stream and list the unanswered queries
    --verify-pack       check after every pack stage that live list A is a permutation of the batch with the words / seeds
                        of the first iteration;  --trace: synchronise and report after every launch
    --watchdog T        dump the Python stack and exit when a (k, n) run takes longer than T seconds
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
    ap.add_argument("--n", type=int, default=0, help="number of queries (default: the workload's)")
    ap.add_argument("--trace", action="store_true", help="synchronize and report after every launch (finding a stuck kernel)")
    ap.add_argument("--verify-pack", action="store_true", help="check the live list the pack stage leaves, every iteration")
    ap.add_argument("--prefill", type=int, default=-1, help="what --postmortem writes into the output buffer before every iteration")
    ap.add_argument("--postmortem", type=float, default=0, help="seconds an iteration may take before its outputs and counters are examined from another stream")
    ap.add_argument("--watchdog", type=int, default=90, help="seconds a single (k, n) run may take before the process dumps its stack and exits")
    ap.add_argument("--also", default="", help="further runs on the same index: k:n,k:n,...")
    ap.add_argument("--superblock-shift", type=int, default=0, help="cut the index into superblocks of 2^s blocks: 64-bit positions, the WIDE kernels")
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
    from harness import synth
    reads, rle_host, total = bench.build_reads_and_bwt(cfg, dev, True)
    torch.cuda.synchronize()
    torch.cuda.empty_cache()
    opts = {}
    if args.superblock_shift:
        opts["superblock_shift"] = args.superblock_shift  # several superblocks: 64-bit positions (the WIDE kernels)
        opts["oct_index"] = 1
    bwt = M.RleBWT.new(devices=[0], **opts)
    bwt.load_vector(rle_host)
    stream = torch.cuda.current_stream().cuda_stream
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
    runs = [(k, args.n or (cfg["n_read"] + cfg["n_random"]))]
    for item in filter(None, args.also.split(",")):   # further (k, n) pairs on the same index
        kk, nn = item.split(":")
        runs.append((int(kk), int(nn)))
    import faulthandler
    pm = {"t": None, "it": -1}
    if args.postmortem:   # the PRODUCT build, untouched: when an iteration does not end, look at what it left behind
        import threading, time
        import numpy as np
        side_pm = torch.cuda.Stream()

        def fetch(t):
            with torch.cuda.stream(side_pm):
                h = t.to("cpu", non_blocking=True)
            side_pm.synchronize()
            return h

        def watch_pm():
            while True:
                time.sleep(0.5)
                if pm["t"] is not None and time.time() - pm["t"] > args.postmortem:
                    kk, nn, lay_live = pm["k"], pm["n"], pm["live_at"]
                    tail = fetch(pm["d_packed"][lay_live:lay_live + 8]).numpy().view(np.uint64)
                    out1 = fetch(pm["d_out"]).numpy()
                    time.sleep(1.0)
                    out2 = fetch(pm["d_out"]).numpy()
                    missing = np.flatnonzero(out2 == args.prefill)
                    print(f"[postmortem] iteration {pm['it']} stuck: live A {int(tail[0])}, live B {int(tail[1])}, work counter "
                          f"{int(tail[2]) & 0xFFFFFFFF}, split flag {int(tail[2]) >> 32:#x}, unanswered {missing.size} (a second earlier {int((out1 == args.prefill).sum())})",
                          file=sys.stderr, flush=True)
                    if missing.size:
                        print(f"[postmortem] first unanswered queries: {missing[:24].tolist()}", file=sys.stderr)
                        qs = fetch(pm["queries"][torch.as_tensor(missing[:6], device=dev)]) if False else None
                        np.save("gpurun_out/postmortem_missing.npy", missing[:100000])
                    pipe = os.environ.get("CUDA_COREDUMP_PIPE")
                    if pipe and os.environ.get("CUDA_ENABLE_USER_TRIGGERED_COREDUMP") == "1":   # GPU core dump of the stuck kernel
                        core = os.environ.get("CUDA_COREDUMP_FILE", "")
                        print(f"[postmortem] asking the driver for a GPU core dump through {pipe} (exists: {os.path.exists(pipe)})", file=sys.stderr, flush=True)
                        try:
                            with open(pipe, "w") as f:
                                f.write("dump\n")
                        except Exception as ex:
                            print(f"[postmortem] pipe: {ex}", file=sys.stderr, flush=True)
                        last, same = -1, 0
                        for _ in range(240):
                            time.sleep(0.5)
                            sz = os.path.getsize(core) if core and os.path.exists(core) else -1
                            same = same + 1 if (sz == last and sz > 0) else 0
                            last = sz
                            if same >= 6:
                                break
                        print(f"[postmortem] core file {core}: {last} bytes", file=sys.stderr, flush=True)
                    os._exit(3)

        threading.Thread(target=watch_pm, daemon=True).start()
    for k, n in runs:
        faulthandler.dump_traceback_later(args.watchdog, exit=True)  # a stuck run says where, and costs no more than this
        print(f"[pack_ab] k={k} n={n}", file=sys.stderr, flush=True)
        queries = synth.make_queries(reads, k, n if not cfg["n_random"] else n // 2, 0 if not cfg["n_random"] else n - n // 2)
        n = queries.shape[0]
        d_packed = torch.empty(bwt.packed_bytes(k, n) // 8, dtype=torch.int64, device=dev)
        d_out = torch.empty(n, dtype=torch.int64, device=dev)
        d_status = torch.zeros(1, dtype=torch.int32, device=dev)
        print(f"[pack_ab] buffers: flush {flush.data_ptr():#x} queries {queries.data_ptr():#x} (+{queries.numel():#x}) packed {d_packed.data_ptr():#x} "
              f"(+{d_packed.numel() * 8:#x}) out {d_out.data_ptr():#x} (+{d_out.numel() * 8:#x}) status {d_status.data_ptr():#x} reads {reads.data_ptr():#x}",
              file=sys.stderr, flush=True)
        pack_ms, search_ms = [], []
        for it in range(args.iters + 3):
            flush.fill_(it & 1)
            e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
            if args.postmortem:
                d_out.fill_(args.prefill)
                torch.cuda.synchronize()
                lay_live = bwt.packed_bytes(k, n) // 8 - 8
                pm.update(k=k, n=n, live_at=lay_live, d_packed=d_packed, d_out=d_out, queries=queries, it=it, t=time.time())
            e[0].record()
            bwt.pack_kmers_device(queries.data_ptr(), k, n, d_packed.data_ptr(), d_out.data_ptr(), d_status.data_ptr(), stream)
            e[1].record()
            if args.verify_pack:   # is what the pack stage left a permutation of the batch, with the words / seeds of iteration 0?
                torch.cuda.synchronize()
                live = bwt.pack_stats(d_packed.data_ptr(), k, n)["live_a"]
                qi = d_packed[2 * n: 2 * n + (n + 1) // 2].view(torch.int32)[:live] & ((1 << 30) - 1)
                order = torch.argsort(qi)
                qs = qi[order]
                perm_ok = bool((qs == torch.arange(live, device=dev, dtype=qs.dtype)).all().item()) if live == n else False
                w_now, s_now = d_packed[:live][order].clone(), d_packed[n:n + live][order].clone()
                if it == 0:
                    ref_pack = (w_now, s_now)
                    print(f"[pack_ab] it 0: list A is a permutation: {perm_ok}", file=sys.stderr, flush=True)
                else:
                    dw = int((w_now != ref_pack[0]).sum().item()) if live == n else -1
                    ds = int((s_now != ref_pack[1]).sum().item()) if live == n else -1
                    print(f"[pack_ab] it {it}: list A is a permutation: {perm_ok}, words differing from it 0: {dw}, seeds: {ds}", file=sys.stderr, flush=True)
                    if not perm_ok:
                        cnt = torch.bincount(qi.long(), minlength=n)
                        print(f"   queries absent from the list: {torch.nonzero(cnt == 0).flatten()[:24].tolist()}; twice: {torch.nonzero(cnt > 1).flatten()[:24].tolist()}",
                              file=sys.stderr, flush=True)
                    del w_now, s_now
                del qi, order, qs
            if args.trace:
                torch.cuda.synchronize()
                print(f"[pack_ab] it {it}: packed {bwt.pack_stats(d_packed.data_ptr(), k, n)}", file=sys.stderr, flush=True)
            bwt.count_kmers_packed_device(d_packed.data_ptr(), k, n, d_out.data_ptr(), stream)
            e[2].record()
            torch.cuda.synchronize()
            pm["t"] = None
            if args.trace:
                print(f"[pack_ab] it {it}: searched, checksum {int(d_out.sum().item())}", file=sys.stderr, flush=True)
            if it == 0:
                print(f"[pack_ab] first iteration: pack {e[0].elapsed_time(e[1]):.3f} ms, search {e[1].elapsed_time(e[2]):.3f} ms", file=sys.stderr, flush=True)
            if it >= 3:
                pack_ms.append(e[0].elapsed_time(e[1]))
                search_ms.append(e[1].elapsed_time(e[2]))
        assert int(d_status.item()) == 0
        total_ms = statistics.median([a + b for a, b in zip(pack_ms, search_ms)])
        print(json.dumps({
            "library": os.environ.get("MSBWT_LIBRARY_PATH", "in-tree"), "workload": args.workload, "k": k, "queries": n,
            "bwt_symbols": int(total), "oct_index": bool(bwt.oct_index), "final_index": bool(bwt.final_index),
            "quad_index": bool(bwt.quad_index), "suffix_table_s": bwt.suffix_table_s,
            "superblock_shift": args.superblock_shift, "index_bytes": bwt.index_bytes,
            "final_fast": os.environ.get("MSBWT_FINAL_FAST", ""), "fixed_k_pack": os.environ.get("MSBWT_PACK_FIXED_K", ""),
            "pack_ms_median": statistics.median(pack_ms), "pack_ms_min": min(pack_ms),
            "search_ms_median": statistics.median(search_ms), "search_ms_min": min(search_ms),
            "queries_per_s": n / (total_ms / 1e3),
            "checksum": int(d_out.sum().item()), "present": int((d_out > 0).sum().item()),
            "split_flag": hex((int(d_packed[bwt.packed_bytes(k, n) // 8 - 6].item()) >> 32) & 0xFFFFFFFF),
        }), flush=True)
        del queries, d_packed, d_out
    faulthandler.cancel_dump_traceback_later()


if __name__ == "__main__":
    main()
