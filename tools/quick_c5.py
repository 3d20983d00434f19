#!/usr/bin/env python
"""Times the pieces of bench.py's C5 entry (MVSEC-shaped: 346x260, 9 bins, B = 512, ~100 k events per sample) one by one:
voxel grid on each kernel family, the bilinear 224x224 copy of the pair (development tool)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import bench  # noqa: E402


def main():
    import eventpretrain_b200 as ep
    from eventpretrain_b200.view_augment import ViewChoice
    dev = torch.device("cuda", 0)
    h, w, bins, Bc = 260, 346, 9, 512
    if len(sys.argv) > 1:
        h, w, bins, Bc = (int(v) for v in sys.argv[1:5])
    e5 = bench.make_batch_gpu(0, dev, batch=Bc, mean=100_000, size=(h, w), spread=0.5, seed=5000, window_us=50_000)
    host = ep.RaggedEvents(e5.x.cpu(), e5.y.cpu(), e5.t.cpu(), e5.p.cpu(), e5.offsets.cpu(), e5.offsets_host, e5.t_div)
    t5 = host.transport().to(dev)
    o = {"voxel": torch.empty((Bc, bins, h, w), dtype=torch.float32, device=dev)}
    full = [ViewChoice(0, 0, w, h, False, False, False)] * Bc

    def timed(fn, steps=10):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(steps):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / steps

    ref = None
    for m in ("global", "tiled", "plane", None):
        try:
            ms = timed(lambda: ep.bin_events(t5, (h, w), num_bins=bins, out=o, method=m))
        except RuntimeError as e:
            print(f"bin {m}: {e}")
            continue
        if ref is None:
            ref = o["voxel"].clone()
        print(f"bin {str(m):7s} {ms:.3f} ms  identical={torch.equal(ref, o['voxel'])}  ({t5.num_events} events, {o['voxel'].numel() * 4 / 1e9:.2f} GB out)", flush=True)
    ms = timed(lambda: ep.apply_views(o["voxel"], full, (224, 224), "bilinear"))
    print(f"view bilinear 224x224 (list of choices converted per call): {ms:.3f} ms")
    prep = ep.prepare_views(full, h, w, dev)
    ms = timed(lambda: ep.apply_views(o["voxel"], prep, (224, 224), "bilinear"))
    print(f"view bilinear 224x224 (prepared views): {ms:.3f} ms")


if __name__ == "__main__":
    main()
