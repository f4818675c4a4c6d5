"""K4 with the bulk copy: random 128-byte line reads issued as ONE cp.async.bulk per lane (granule code 129) against
eight lanes x 16-byte loads per line (granule 128), same buffer, same index stream."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rust_msbwt_b200 as M  # noqa: E402

dev = torch.device("cuda", 0)
st = torch.cuda.current_stream().cuda_stream
res = {}
for gb in (2, 24):
    buf = torch.empty(gb << 30, dtype=torch.uint8, device=dev)
    sink = torch.zeros(1, dtype=torch.int64, device=dev)
    for gran in (128, 129):
        ng = 1 << 27
        best = None
        for it in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            M.gather_bench(0, buf.data_ptr(), buf.numel(), gran, ng, 99 + it, sink.data_ptr(), st)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
            best = ms if best is None else min(best, ms)
        res[f"{gb}GB_{'bulk_per_lane' if gran == 129 else 'ldg_8_lanes'}"] = {"lines_per_s": ng / (best / 1e3), "gb_per_s": ng * 128 / (best / 1e3) / 1e9}
    del buf
print(json.dumps(res))
