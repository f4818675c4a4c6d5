"""Host-side packer throughput vs worker count (msbwt_debug_host_pack), and raw host read bandwidth
(tools/membw.cpp) on the same box: how much of the end-to-end path the CPU side can feed."""
import ctypes as C
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rust_msbwt_b200 as M  # noqa: E402

L = M.load_library()
n, k = 16_000_000, 31
rng = np.random.default_rng(1)
q = rng.choice(np.array([1, 2, 3, 5], dtype=np.uint8), size=(n, k))
words = np.zeros((1, n), dtype=np.uint64)
exc = np.zeros(16, dtype=np.uint64)
ne = C.c_uint64(0)
res = {"pack": {}, "membw": {}}
for th in (1, 2, 4, 8, 12, 16, 24, 32):
    if th > 2 * (os.cpu_count() or 1):
        break
    best = 1e9
    for it in range(3):
        t = time.perf_counter()
        L.msbwt_debug_host_pack(C.c_void_p(q.ctypes.data), k, n, th, C.c_void_p(words.ctypes.data),
                                C.c_void_p(exc.ctypes.data), 16, C.byref(ne))
        best = min(best, time.perf_counter() - t)
    res["pack"][th] = {"ms": best * 1e3, "GBps_in": n * k / best / 1e9, "Mqps": n / best / 1e6}
exe = "/tmp/membw"
subprocess.run(["g++", "-O2", "-pthread", os.path.join(ROOT, "tools", "membw.cpp"), "-o", exe], check=True)
for th in (1, 4, 8, 16):
    out = subprocess.run([exe, str(th)], capture_output=True, text=True).stdout.strip().splitlines()[-1]
    res["membw"][th] = out
print(json.dumps(res, indent=1))
