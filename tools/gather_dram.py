"""Run the K4 gather kernel once per granule (32/64/128 B) over a 2 GiB buffer so that
`ncu --metrics dram__bytes_read.sum,gpu__time_duration.sum -k regex:gather_kernel` can show how many DRAM
bytes one random read of each size costs on this GPU."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rust_msbwt_b200 as M  # noqa: E402

torch.cuda.set_device(0)
buf = torch.empty(2 << 30, dtype=torch.uint8, device="cuda")
buf.random_(0, 255)
sink = torch.zeros(1, dtype=torch.int64, device="cuda")
stream = torch.cuda.current_stream().cuda_stream
for gran in (32, 64, 128):
    for it in range(2):
        M.gather_bench(0, buf.data_ptr(), buf.numel(), gran, 1 << 26, 99 + it, sink.data_ptr(), stream)
torch.cuda.synchronize()
print("gather launches done: 2 x (32, 64, 128 B), 2^26 reads each")
