#!/usr/bin/env python
"""Turns the files a `tools/run_final.sh` + `tools/run_o.sh` pair left under gpurun_out/ into the committed summaries under
profiles/:   python tools/make_profiles.py r02za r02     (input tag, output prefix)"""
import collections
import csv
import json
import shutil
import subprocess
import sys

tag, out = sys.argv[1], sys.argv[2]
G, P = "gpurun_out/", "profiles/"

# ---- DRAM traffic (ncu --cache-control none) ----
rows = [r for r in csv.reader(open(f"{G}{tag}_traffic.csv")) if len(r) > 10 and r[0].isdigit()]
d = collections.OrderedDict()
for r in rows:
    d.setdefault((int(r[0]), r[4]), {})[r[12]] = float(r[14].replace(",", ""))
steps, cur = [], []
for (i, name), v in d.items():
    short = "k_sample_meta" if "k_sample_meta" in name else name.split("(")[0].split("::")[-1]
    if "k_sample_meta" in name and cur:
        steps.append(cur); cur = []
    cur.append((short, v))
steps.append(cur)
last = [s for s in steps if len(s) >= 5][-1]
lines = ["# DRAM traffic of one headline step (256 x ~1M events, 640x480, 5 bins + sum plane, 4 B packed layout), per launch.",
         "# ncu --cache-control none --clock-control none --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct",
         "# command: python tools/quick_bin.py --batch 256 --packed4 --methods tiled --steps 1   (tools/run_final.sh)",
         "kernel,dram_read_bytes,dram_write_bytes,duration_ns,l2_hit_pct"]
tot = 0
for short, v in last:
    lines.append(f"{short},{int(v['dram__bytes_read.sum'])},{int(v['dram__bytes_write.sum'])},{int(v['gpu__time_duration.sum'])},{v['lts__t_sector_hit_rate.pct']}")
    tot += v["dram__bytes_read.sum"] + v["dram__bytes_write.sum"]
lines.append(f"# total DRAM bytes per step: {int(tot)}")
open(f"{P}{out}_traffic_tiled.csv", "w").write("\n".join(lines) + "\n")
per = {short: int(v["dram__bytes_read.sum"] + v["dram__bytes_write.sum"]) for short, v in last}
json.dump({"source": f"profiles/{out}_traffic_tiled.csv: ncu --cache-control none --clock-control none on the headline batch and layout (256 x ~1M events, "
                     "640x480, 5 bins + sum plane, 4 B packed), DRAM read + write bytes per launch; one step = one launch of each kernel",
           "dram_bytes_per_launch": per, "launches_per_step": {k: 1 for k in per}, "dram_bytes_per_step": int(tot),
           "how": f"ncu --cache-control none (caches not flushed between the kernels of a step), dram__bytes_read.sum + dram__bytes_write.sum, per launch x launches per step; profiles/{out}_traffic_tiled.csv",
           "note": "compulsory bytes of this design: events 1.02 GB read by the route, routed records 1.02 GB written and read back, outputs 1.89 GB written = 4.95 GB; SURVEY 8(d) algorithmic bytes 5.20 GB"},
          open(f"{P}ncu_traffic.json", "w"), indent=1)
print("\n".join(lines))

# ---- launch list of the bench command ----
rows = [r for r in csv.reader(open(f"{G}{tag}_launches.csv")) if len(r) > 10 and r[0].isdigit()]
agg, seq = collections.OrderedDict(), []
for r in rows:
    short = r[4].replace("unnamed>::", "").replace("void ", "")[:70]
    t = float(r[14].replace(",", "")) / 1000
    a = agg.setdefault(short, [0, 0.0]); a[0] += 1; a[1] += t
    seq.append((short, t))
bench = json.load(open(f"{G}{tag}_bench_1gpu.json"))
kr, ks = bench["roofline"]["kernels"]["k_route_ms_per_step"], bench["roofline"]["kernels"]["k_sweep_ms_per_step"]
lines = ["# ncu launch list of `python bench.py --steps 2 --warmup 1` (1 GPU), restricted to the library's kernels (-k regex:k_...),",
         "# --metrics gpu__time_duration.sum --clock-control none -c 400: the warm-up + timed headline steps (k_sample_meta, k_tiled_setup,",
         "# k_tiled_desc, k_route, k_sweep per step), the statistics variant (+ k_stats_slices, k_stats_final), the per-kernel profiling",
         "# pass, then the global-RED legs of extra.layouts (k_scatter / k_finalize_voxel, 43 groups per step) until the capture limit.",
         "# Per-launch times under ncu are cold-cache and serialised; what compares with the CUDA-event figures of the bench line is the",
         f"# SHARE of k_route vs k_sweep within a headline step (CUDA events: {kr:.3f} ms / {ks:.3f} ms = {kr / (kr + ks) * 100:.1f} % / {ks / (kr + ks) * 100:.1f} %).",
         f"{'kernel':72s} {'launches':>8s} {'total us':>12s} {'us/launch':>10s}"]
for k, v in sorted(agg.items(), key=lambda x: -x[1][1]):
    lines.append(f"{k:72s} {v[0]:8d} {v[1]:12.1f} {v[1] / v[0]:10.1f}")
r = [v for k, v in agg.items() if k.startswith("k_route<0, 0>")]
s = [v for k, v in agg.items() if k.startswith("k_sweep<1, 1>") or k.startswith("k_sweep<1>")]
if r and s:
    a, b = r[0][1] / r[0][0], s[0][1] / s[0][0]
    lines.append(f"# headline step under ncu: k_route {a:.1f} us/launch, k_sweep {b:.1f} us/launch -> shares {a / (a + b) * 100:.1f} % / {b / (a + b) * 100:.1f} %")
lines.append("# first 12 launches in order:")
for k, t in seq[:12]:
    lines.append(f"#   {k:60s} {t:10.1f} us")
open(f"{P}{out}_bench_launches.txt", "w").write("\n".join(lines) + "\n")
print("\n".join(lines[6:20]))

# ---- full ncu captures ----
import os
caps = [(f"{G}{tag}_tiled.ncu-rep", f"{P}{out}_binning_tiled_ncu.txt", ""), (f"{G}{tag}_evrep.ncu-rep", f"{P}{out}_evrep_ncu.txt", "k_evrep"),
        (f"{G}{tag}_plane.ncu-rep", f"{P}{out}_plane_ncu.txt", "k_plane")]
for rep, name, filt in [c for c in caps if os.path.exists(c[0])]:
    txt = subprocess.run([sys.executable, "tools/ncu_summary.py", rep], capture_output=True, text=True).stdout
    txt += subprocess.run([sys.executable, "tools/ncu_lines.py", rep, filt, "30"], capture_output=True, text=True).stdout
    open(name, "w").write(txt)
shutil.copy(f"{G}{tag}_bench_1gpu.json", f"{P}{out}_bench_1gpu.json")
shutil.copy(f"{G}{tag}_bench_ref.json", f"{P}{out}_bench_reference_arm.json")

# ---- DRAM traffic of the whole-plane path (reference-res 224x224 step) ----
if os.path.exists(f"{G}{tag}_traffic_plane.csv"):
    rows = [r for r in csv.reader(open(f"{G}{tag}_traffic_plane.csv")) if len(r) > 10 and r[0].isdigit()]
    d = collections.OrderedDict()
    for r in rows:
        d.setdefault((int(r[0]), r[4]), {})[r[12]] = float(r[14].replace(",", ""))
    steps, cur = [], []
    for (i, name), v in d.items():
        short = "k_sample_meta" if "k_sample_meta" in name else name.split("(")[0].split("::")[-1]
        if "k_sample_meta" in name and cur:
            steps.append(cur); cur = []
        cur.append((short, v))
    steps.append(cur)
    last = [s_ for s_ in steps if len(s_) >= 5][-1]
    lines = ["# DRAM traffic of one reference-res step (256 x ~1M events, 640x480 -> 224x224 fused, 5 bins + sum plane, 4 B packed layout), per launch:",
             "# the whole-plane kernels and, behind them, the stand-by route + sweep launches that return at once (sorted input).",
             "# ncu --cache-control none --clock-control none --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct",
             "# command: python tools/quick_bin.py --batch 256 --packed4 --size 224x224 --methods plane --steps 1   (tools/run_final.sh)",
             "kernel,dram_read_bytes,dram_write_bytes,duration_ns,l2_hit_pct"]
    tot = 0
    for short, v in last:
        lines.append(f"{short},{int(v['dram__bytes_read.sum'])},{int(v['dram__bytes_write.sum'])},{int(v['gpu__time_duration.sum'])},{v['lts__t_sector_hit_rate.pct']}")
        tot += v["dram__bytes_read.sum"] + v["dram__bytes_write.sum"]
    lines.append(f"# total DRAM bytes per step: {int(tot)}  (compulsory: events 1.02 GB once + outputs 0.31 GB; algorithmic, SURVEY 8(d): 3.62 GB)")
    open(f"{P}{out}_traffic_plane.csv", "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))
