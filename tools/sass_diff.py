"""Which kernels does a source change touch?  Compiles every CUDA translation unit of the engine at a git revision
and in the working tree to sm_100a cubins (extra nvcc flags apply to the working tree only: a tuning variant against
the committed default) and compares the SASS function by function.

Used when GPU time is short: a change whose diff is only NEW functions cannot have altered what the last GPU run
validated.

    python tools/sass_diff.py [--rev HEAD] [--flags=-DMSBWT_X=1 ...] [kernels.cu quad_kernels.cu ...]
"""
from __future__ import annotations

import argparse
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "rust-msbwt_b200"))
import build as B  # noqa: E402

NVCC = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-cubin"]


def functions(cubin: str) -> dict[str, list[str]]:
    out = subprocess.run(["cuobjdump", "-sass", cubin], capture_output=True, text=True, check=True).stdout
    d, cur = {}, None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            d[cur] = []
        elif cur and re.match(r"\s+/\*[0-9a-f]{4}\*/", line):
            d[cur].append(re.sub(r"/\*[0-9a-f]{4}\*/", "", line.split(";")[0]).strip())
    return d


def demangle(name: str) -> str:
    """full demangled signature: the key functions are matched by (anonymous-namespace symbols carry a hash of the
    source path in their mangled names, which differs between the checkout of the revision and the working tree)"""
    r = subprocess.run(["cu++filt", name], capture_output=True, text=True)
    return r.stdout.strip() or name


def short(sig: str) -> str:
    return sig[:sig.index(">(") + 1] if ">(" in sig else sig.split("(")[0]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rev", default="HEAD")
    ap.add_argument("--flags", action="append", default=[])
    ap.add_argument("units", nargs="*", default=[s for s in B.SOURCES])
    args = ap.parse_args()
    with tempfile.TemporaryDirectory() as tmp:
        files = subprocess.run(["git", "ls-tree", "-r", "--name-only", args.rev, "--", "rust-msbwt_b200/csrc", "include"],
                               cwd=ROOT, capture_output=True, text=True, check=True).stdout.split()
        for f in files:
            dst = os.path.join(tmp, "rev", f)
            os.makedirs(os.path.dirname(dst), exist_ok=True)
            with open(dst, "wb") as fh:
                fh.write(subprocess.run(["git", "show", f"{args.rev}:{f}"], cwd=ROOT, capture_output=True, check=True).stdout)
        changed = False
        for unit in args.units:
            cub = {}
            for side, base, flags in (("rev", os.path.join(tmp, "rev"), []), ("tree", ROOT, args.flags)):
                cub[side] = os.path.join(tmp, f"{side}_{unit}.cubin")
                src = os.path.join(base, "rust-msbwt_b200", "csrc", unit)
                if not os.path.exists(src):
                    cub[side] = None
                    continue
                subprocess.run([B.nvcc_path(), *NVCC, *flags, "-o", cub[side], unit], cwd=os.path.dirname(src), check=True)
            a = {demangle(k): v for k, v in (functions(cub["rev"]) if cub["rev"] else {}).items()}
            b = {demangle(k): v for k, v in (functions(cub["tree"]) if cub["tree"] else {}).items()}
            for k in sorted(set(a) | set(b)):
                tag = "new " if k not in a else ("gone" if k not in b else ("same" if a[k] == b[k] else "DIFF"))
                if tag != "same":
                    changed = True
                    print(f"{tag} {unit:18s} {short(k)[:110]}  ({len(a.get(k, []))} -> {len(b.get(k, []))} instructions)")
            print(f"{unit}: {len(b)} functions, {sum(1 for k in b if k in a and a[k] == b[k])} identical to {args.rev}")
        return 1 if changed else 0


if __name__ == "__main__":
    sys.exit(main())
