"""Random-gather rate over very large buffers (does the TLB / page-table walk throttle an index of
tens of GB?).  Prints one JSON object; run on a B200:  python tools/gather_big.py"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rust_msbwt_b200 as M  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    torch.cuda.set_device(0)
    out = {"gpu": torch.cuda.get_device_properties(0).name, "results": []}
    sink = torch.zeros(1, dtype=torch.int64, device=dev)
    stream = torch.cuda.current_stream().cuda_stream
    big = torch.empty(128 << 30, dtype=torch.uint8, device=dev)
    big[::4096].zero_()  # touch every page
    for gb in (2, 16, 64, 128):
        nbytes = gb << 30
        for gran in (32, 128):
            n = 1 << 28
            best = None
            for it in range(3):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                M.gather_bench(0, big.data_ptr(), nbytes, gran, n, 77 + it, sink.data_ptr(), stream)
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1)
                if it and (best is None or ms < best):
                    best = ms
            out["results"].append({"buffer_gb": gb, "granule": gran, "reads_per_s": n / (best / 1e3),
                                   "gb_per_s": n * gran / (best / 1e3) / 1e9})
    print(json.dumps(out))


if __name__ == "__main__":
    main()
