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
    print(json.dumps(res), flush=True)


if __name__ == "__main__":
    sys.exit(main())
