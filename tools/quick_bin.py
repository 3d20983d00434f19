#!/usr/bin/env python
"""Quick timing / cross-check of the binning paths on a slice of the bench workload (development tool).

    python tools/quick_bin.py [--batch 64] [--steps 5] [--methods tiled,global] [--check] [--skewed] [--compact]

Prints per method: ms/step, Gevents/s and the library's per-kernel device times (pass 1 = route | scatter,
pass 2 = sweep | finalize).  --check compares the methods bit for bit.
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--methods", default="tiled,global")
    ap.add_argument("--check", action="store_true")
    ap.add_argument("--skewed", action="store_true")
    ap.add_argument("--compact", action="store_true")
    ap.add_argument("--packed", action="store_true")
    ap.add_argument("--packed4", action="store_true")
    ap.add_argument("--size", default="", help="HxW output grid with the fused events_reshape scale, e.g. 224x224")
    ap.add_argument("--bins", type=int, default=bench.BINS)
    ap.add_argument("--count", type=int, default=0)
    ap.add_argument("--stats", action="store_true", help="ep_bin_events_stats: batch statistics as a by-product of the sweep")
    args = ap.parse_args()
    import eventpretrain_b200 as ep
    from eventpretrain_b200 import _lib
    dev = torch.device("cuda", 0)
    L = ep.load_library()
    ev = bench.make_batch_gpu(0, dev, skewed=args.skewed, batch=args.batch)
    if args.compact:
        host = ep.RaggedEvents(ev.x.cpu(), ev.y.cpu(), ev.t.cpu(), ev.p.cpu(), ev.offsets.cpu(), ev.offsets_host, ev.t_div)
        ev = host.compact().to(dev)
    if args.packed:
        host = ep.RaggedEvents(ev.x.cpu(), ev.y.cpu(), ev.t.cpu(), ev.p.cpu(), ev.offsets.cpu(), ev.offsets_host, ev.t_div)
        ev = host.packed().to(dev)
    if args.packed4:
        host = ep.RaggedEvents(ev.x.cpu(), ev.y.cpu(), ev.t.cpu(), ev.p.cpu(), ev.offsets.cpu(), ev.offsets_host, ev.t_div)
        ev = host.packed(4).to(dev)
    n = ev.num_events
    H, W = bench.H, bench.W
    scale = (1.0, 1.0)
    if args.size:
        H, W = (int(v) for v in args.size.split("x"))
        scale = (W / bench.W, H / bench.H)
    results = {}
    for m in args.methods.split(","):
        kw = dict(num_bins=args.bins, voxel_sum=True, count_channels=args.count, method=m, scale=scale, stats=args.stats)
        out = ep.bin_events(ev, (H, W), check=True, **kw)
        for _ in range(2):
            ep.bin_events(ev, (H, W), out=out, **kw)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            ep.bin_events(ev, (H, W), out=out, **kw)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.steps
        L.ep_profile_enable(1)
        for _ in range(args.steps):
            ep.bin_events(ev, (H, W), out=out, **kw)
        prof = _lib.ProfileStats()
        L.ep_profile_read(prof)
        L.ep_profile_enable(0)
        alg = 13 * n + 4 * (args.bins + 1) * H * W * args.batch
        print(f"{m:7s} batch {args.batch} events {n}: {ms:.3f} ms/step  {n / ms / 1e6:.1f} Gev/s  "
              f"{alg / ms / 1e6:.0f} GB/s algorithmic | pass1 {prof.ms[0] / args.steps:.3f} ms ({prof.launches[0] // args.steps} launches) "
              f"pass2 {prof.ms[1] / args.steps:.3f} ms other {prof.ms[2] / args.steps:.3f} ms", flush=True)
        results[m] = {k: v.clone() for k, v in out.items()}
    if args.check and len(results) > 1:
        names = list(results)
        for k in results[names[0]]:
            same = all(torch.equal(results[names[0]][k], results[o][k]) for o in names[1:])
            print(f"check {k}: {'identical' if same else 'DIFFERENT'}")
            if not same:
                d = (results[names[0]][k] - results[names[1]][k]).abs()
                print("   max abs diff", float(d.max()), "n diff", int((d > 0).sum()))


if __name__ == "__main__":
    main()
