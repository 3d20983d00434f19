#!/usr/bin/env python
"""EvRep timing on the C4 shape (DSEC-like 640x440, ~2 M events per sample): routed path (4 B transport layout) against the
global counting sort (canonical layout).   python tools/quick_evrep.py [--batch 32] [--steps 5] [--check]"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--mean", type=int, default=2_000_000)
    ap.add_argument("--check", action="store_true")
    ap.add_argument("--only-tiled", action="store_true")
    args = ap.parse_args()
    import eventpretrain_b200 as ep
    dev = torch.device("cuda", 0)
    h, w = 440, 640
    e4 = bench.make_batch_gpu(0, dev, batch=args.batch, mean=args.mean, size=(h, w), spread=0.1, seed=4000, window_us=100_000)
    host = ep.RaggedEvents(e4.x.cpu(), e4.y.cpu(), e4.t.cpu(), e4.p.cpu(), e4.offsets.cpu(), e4.offsets_host, e4.t_div)
    t4 = host.packed(4).to(dev)
    out = torch.empty((args.batch, 3, h, w), dtype=torch.float64, device=dev)

    def timed(fn):
        fn(); fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / args.steps
    ms = timed(lambda: ep.evrep(t4, (h, w), out=out))
    print(f"routed  batch {args.batch} events {e4.num_events}: {ms:.3f} ms  {e4.num_events / ms / 1e6:.1f} Gev/s", flush=True)
    if not args.only_tiled:
        ref = torch.empty_like(out)
        ms2 = timed(lambda: ep.evrep(e4, (h, w), out=ref))
        print(f"global  batch {args.batch} events {e4.num_events}: {ms2:.3f} ms  {e4.num_events / ms2 / 1e6:.1f} Gev/s")
        if args.check:
            print("identical" if torch.equal(out, ref) else "DIFFERENT")


if __name__ == "__main__":
    main()
