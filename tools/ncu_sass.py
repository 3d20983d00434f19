#!/usr/bin/env python
"""Hot SASS of one kernel of an ncu report (--import-source on), in address order, with executed counts and stall samples.

    python tools/ncu_sass.py report.ncu-rep kernel-substring [min-fraction-of-max]
"""
import csv
import subprocess
import sys


def main():
    path, want = sys.argv[1], sys.argv[2]
    frac = float(sys.argv[3]) if len(sys.argv) > 3 else 0.25
    raw = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
    fn = fp = hdr = cur = None
    out = []
    for r in csv.reader(raw.splitlines()):
        if not r:
            continue
        if r[0] == "Function Name":
            fn = r[1]; continue
        if r[0] == "File Path":
            fp = r[1]; continue
        if r[0] == "Line No":
            hdr = r; continue
        if hdr and len(r) == len(hdr):
            if r[0] != "":
                cur = (fp.split("/")[-1], r[0]); continue
            if want in fn:
                try:
                    out.append((r[2], r[3].strip(), int(r[hdr.index("Instructions Executed")]), int(r[hdr.index("# Samples")]), cur))
                except ValueError:
                    pass
    out.sort(key=lambda x: x[0])
    mx = max(o[2] for o in out)
    for o in out:
        if o[2] > mx * frac:
            print(o[0][-5:], f"{o[2]:9d} {o[3]:5d}", f"{o[1][:80]:80s}", f"{o[4][0]}:{o[4][1]}")


if __name__ == "__main__":
    main()
