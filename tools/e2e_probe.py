"""Per-chunk timeline of the packed-integer end-to-end route (MSBWT_TRACE_PIPE=1) on a workload's batch, and the
call's wall time without tracing: where the copy engines and the kernels of the lanes actually overlap."""
import ctypes
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import numpy as np
    import torch

    import bench
    import rust_msbwt_b200 as M
    wl = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
    dev = torch.device("cuda", 0)
    cfg = dict(bench.WORKLOADS[wl])
    k = cfg["k"]
    rle_host, total, queries, _ = bench.build_workload(cfg, dev, 0)
    n = queries.shape[0]
    km = torch.empty(n, dtype=torch.int64, pin_memory=True)
    km.copy_(bench.encode_u64(queries, k))
    del queries
    out = torch.empty(n, dtype=torch.int64, pin_memory=True)
    bwt = M.RleBWT.new(devices=[0])
    bwt.load_vector(rle_host)
    lib = M.load_library()
    km_np, out_np = km.numpy().view(np.uint64), out.numpy().view(np.uint64)

    def call():
        rc = lib.msbwt_count_kmers_u64(bwt.handle, ctypes.c_void_p(km_np.ctypes.data), k, n, ctypes.c_void_p(out_np.ctypes.data))
        assert rc == 0
    call()
    for _ in range(3):
        t0 = time.perf_counter()
        call()
        print(f"wall {1e3 * (time.perf_counter() - t0):.2f} ms for {n} queries", file=sys.stderr, flush=True)
    # the same call on PAGEABLE host memory (what a caller's Vec<u64> / numpy array is): cudaMemcpyAsync then stages
    # through the driver and holds the calling thread
    km_pg, out_pg = km_np.copy(), np.zeros(n, dtype=np.uint64)

    def call_pageable():
        rc = lib.msbwt_count_kmers_u64(bwt.handle, ctypes.c_void_p(km_pg.ctypes.data), k, n, ctypes.c_void_p(out_pg.ctypes.data))
        assert rc == 0
    call_pageable()
    for _ in range(3):
        t0 = time.perf_counter()
        call_pageable()
        print(f"pageable wall {1e3 * (time.perf_counter() - t0):.2f} ms for {n} queries", file=sys.stderr, flush=True)
    assert (out_pg == out_np).all()
    if len(sys.argv) > 2 and sys.argv[2] == "bytes":
        q_pg = None
    os.environ["MSBWT_TRACE_PIPE"] = "1"
    call()


if __name__ == "__main__":
    main()
