#!/usr/bin/env python
"""Shared-memory wavefronts (actual / ideal) per source line of an ncu report captured with --import-source on.

    python tools/ncu_wavefronts.py report.ncu-rep [kernel-substring] [top-n]
"""
import csv
import subprocess
import sys


def main():
    path = sys.argv[1]
    want = sys.argv[2] if len(sys.argv) > 2 else ""
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 20
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
                w = int(r[hdr.index("L1 Wavefronts Shared")] or 0)
                wi = int(r[hdr.index("L1 Wavefronts Shared Ideal")] or 0)
                g = int(r[hdr.index("L2 Theoretical Sectors Global")] or 0)
            except ValueError:
                continue
            key = (fn.split("(")[0], fp.split("/")[-1], line, r[1].strip()[:110])
            a = agg.setdefault(key, [0, 0, 0]); a[0] += w; a[1] += wi; a[2] += g
    for f in sorted({k[0] for k in agg}):
        if want not in f:
            continue
        items = [(k, v) for k, v in agg.items() if k[0] == f]
        tot = sum(v[0] for _, v in items) or 1
        print(f"===== {f}: {tot} shared wavefronts, {sum(v[1] for _, v in items)} ideal, {sum(v[2] for _, v in items)} global sectors")
        for k, v in sorted(items, key=lambda kv: -kv[1][0])[:top]:
            print(f"{v[0] / tot * 100:5.1f}%  {v[0]:10d} actual {v[1]:10d} ideal  {k[1]}:{k[2]}  {k[3]}")


if __name__ == "__main__":
    main()
