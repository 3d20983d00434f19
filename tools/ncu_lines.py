#!/usr/bin/env python
"""Executed warp instructions and stall samples per source line of an ncu report captured with --import-source on.

    python tools/ncu_lines.py report.ncu-rep [kernel-substring] [top-n]
"""
import csv
import subprocess
import sys


def main():
    path = sys.argv[1]
    want = sys.argv[2] if len(sys.argv) > 2 else ""
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    raw = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
    fn = fp = hdr = None
    agg = {}
    for r in csv.reader(raw.splitlines()):
        if not r:
            continue
        if r[0] == "File Path":
            fp = r[1]; continue
        if r[0] == "Function Name":
            fn = r[1]; continue
        if r[0] == "Line No":
            hdr = r; continue
        if r[0] != "" and hdr and len(r) == len(hdr):
            try:
                line = int(r[0])
                inst = int(r[hdr.index("Instructions Executed")] or 0)
                samp = int(r[hdr.index("# Samples")] or 0)
            except ValueError:
                continue
            key = (fn.split("(")[0], fp.split("/")[-1], line, r[1].strip()[:120])
            a = agg.setdefault(key, [0, 0]); a[0] += inst; a[1] += samp
    kernels = sorted({k[0] for k in agg})
    for f in kernels:
        if want not in f:
            continue
        items = [(k, v) for k, v in agg.items() if k[0] == f]
        tot = sum(v[0] for _, v in items) or 1
        ts = sum(v[1] for _, v in items) or 1
        print(f"===== {f}: {tot} warp instructions, {ts} samples")
        for k, v in sorted(items, key=lambda kv: -kv[1][0])[:top]:
            print(f"{v[0] / tot * 100:5.1f}% inst {v[1] / ts * 100:5.1f}% samp  {k[1]}:{k[2]}  {k[3]}")


if __name__ == "__main__":
    main()
