"""Per-instruction hot spots of an .ncu-rep source page (needs -lineinfo / --import-source on).
Usage: python tools/ncu_hot.py prof.ncu-rep [min_share_pct]"""
import csv
import subprocess
import sys


def main():
    rep = sys.argv[1]
    thr = float(sys.argv[2]) if len(sys.argv) > 2 else 1.5
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr = rows[1]
    si, ai, ei = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
    stall_cols = [i for i, c in enumerate(hdr) if c.startswith("stall_") and "Not Issued" not in c]
    data = [r for r in rows[2:] if len(r) > ei and r[ai] != ""]
    tot = sum(float(r[ai] or 0) for r in data)
    ex_max = max(float(r[ei] or 0) for r in data)
    print(f"instructions: {len(data)}  total samples: {tot:.0f}  max executed: {ex_max:.0f}")
    agg = {}
    for r in data:
        for i in stall_cols:
            agg[hdr[i]] = agg.get(hdr[i], 0) + float(r[i] or 0)
    print("stall totals:", {k: round(100 * v / tot, 1) for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v > tot * 0.01})
    for n, r in enumerate(data):
        v = float(r[ai] or 0)
        if v >= tot * thr / 100:
            top = sorted(((float(r[i] or 0), hdr[i]) for i in stall_cols), reverse=True)[:2]
            print(f"{n:4d} {100 * v / tot:5.1f}%  exec {float(r[ei]) / ex_max:4.2f}  {r[si].strip()[:64]:64s} {[(t[1], round(100 * t[0] / max(v, 1))) for t in top]}")


if __name__ == "__main__":
    main()
