"""Random-gather roofline sweep (K4): sustained rate of independent, uniformly
random, granule-aligned reads of 32 / 64 / 128 bytes over buffers from L2-resident to
DRAM-resident sizes.  Prints one JSON object; run on a B200:  python tools/gather_sweep.py"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rust_msbwt_b200 as M  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    torch.cuda.set_device(0)
    props = torch.cuda.get_device_properties(0)
    out = {"gpu": props.name, "sms": props.multi_processor_count, "l2_bytes": props.L2_cache_size,
           "l2_fetch_granularity_default": M.l2_fetch_granularity(0), "results": []}
    sink = torch.zeros(1, dtype=torch.int64, device=dev)
    stream = torch.cuda.current_stream().cuda_stream
    big = torch.empty(4 << 30, dtype=torch.uint8, device=dev)
    big.random_(0, 255)
    for fetch in (0, 32, 64, 128):
        eff = M.l2_fetch_granularity(0, fetch)
        for mb in (16, 32, 64, 128, 256, 1024, 4096):
            nbytes = mb << 20
            for gran in (32, 64, 128):
                n = 1 << 27
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
                out["results"].append({"l2_fetch_limit": eff, "buffer_mb": mb, "granule": gran,
                                       "reads_per_s": n / (best / 1e3), "gb_per_s": n * gran / (best / 1e3) / 1e9})
    print(json.dumps(out))


if __name__ == "__main__":
    main()
