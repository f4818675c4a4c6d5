"""What the host link of this box sustains: pinned-memory cudaMemcpyAsync H2D alone, D2H alone, and both at once
(GB/s), in chunks like the library's pipelines use.  The end-to-end entry points cannot beat these numbers:
an 8-byte-in / 8-byte-out query needs 8 bytes of each direction at the same time."""
import json
import sys
import time

import torch


def main():
    dev = torch.device("cuda", 0)
    nbytes = 1 << 30
    h_in = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
    h_out = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
    d_in = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    d_out = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    res = {}
    for chunk in (1 << 24, 1 << 27):
        def h2d():
            with torch.cuda.stream(s1):
                for a in range(0, nbytes, chunk):
                    d_in[a:a + chunk].copy_(h_in[a:a + chunk], non_blocking=True)

        def d2h():
            with torch.cuda.stream(s2):
                for a in range(0, nbytes, chunk):
                    h_out[a:a + chunk].copy_(d_out[a:a + chunk], non_blocking=True)

        for name, fns in (("h2d", (h2d,)), ("d2h", (d2h,)), ("both", (h2d, d2h))):
            best = 1e9
            for _ in range(3):
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                for f in fns:
                    f()
                torch.cuda.synchronize()
                best = min(best, time.perf_counter() - t0)
            res[f"{name}_chunk{chunk >> 20}MB_GBps_each_direction"] = nbytes / best / 1e9
    # the library's lane pattern: L streams, each chunk = H2D -> a kernel -> D2H on ONE stream, chunks round-robin
    for lanes in (2, 5, 8):
        for chunk in (1 << 22, 1 << 24):
            streams = [torch.cuda.Stream() for _ in range(lanes)]
            best = 1e9
            nb = 800_000_000 // chunk * chunk
            for _ in range(3):
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                for c, a in enumerate(range(0, nb, chunk)):
                    with torch.cuda.stream(streams[c % lanes]):
                        d_in[a:a + chunk].copy_(h_in[a:a + chunk], non_blocking=True)
                        d_out[a:a + chunk].copy_(d_in[a:a + chunk])          # stands in for the kernels
                        h_out[a:a + chunk].copy_(d_out[a:a + chunk], non_blocking=True)
                torch.cuda.synchronize()
                best = min(best, time.perf_counter() - t0)
            res[f"lanes{lanes}_chunk{chunk >> 20}MB_GBps_each_direction"] = nb / best / 1e9
    print(json.dumps(res), flush=True)


if __name__ == "__main__":
    sys.exit(main())
