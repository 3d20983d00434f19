#!/usr/bin/env python
"""Secondary measurements for the BASELINE.json configs other than the bench.py headline (C2):
C1 single N-Caltech101-shaped sample, C3 masked-modelling input pipeline, C4 DSEC-shaped, C5 MVSEC-shaped.
Prints one JSON object per line; CUDA-event timing, median of `reps` after warm-up; algorithmic bytes per SURVEY §8(d).
    python tools/bench_configs.py > profiles/r01_configs.jsonl
"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import eventpretrain_b200 as ep  # noqa: E402

PEAK = 6552.3
try:
    PEAK = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass
dev = torch.device("cuda", 0)


def timeit(fn, reps=30, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))


def make_events(B, n_mean, spread, H, W, seed, t_span):
    rng = np.random.default_rng(seed)
    counts = np.maximum(1, np.round(n_mean * rng.uniform(1 - spread, 1 + spread, B))).astype(np.int64)
    off = np.zeros(B + 1, np.int64)
    np.cumsum(counts, out=off[1:])
    n = int(off[-1])
    g = torch.Generator(device=dev).manual_seed(seed)
    x = torch.randint(0, W, (n,), device=dev, generator=g, dtype=torch.int32).to(torch.uint16)
    y = torch.randint(0, H, (n,), device=dev, generator=g, dtype=torch.int32).to(torch.uint16)
    p = torch.randint(0, 2, (n,), device=dev, generator=g, dtype=torch.uint8)
    t = torch.empty(n, dtype=torch.int64, device=dev)
    for b in range(B):
        lo, hi = int(off[b]), int(off[b + 1])
        t[lo:hi] = torch.sort(torch.randint(0, t_span, (hi - lo,), device=dev, generator=g)).values
    return ep.RaggedEvents(x, y, t, p, torch.from_numpy(off).to(dev), off, t_div=1e6)


def report(name, ms, alg_bytes, units, unit_name, **extra):
    line = {"config": name, "ms": ms, "GBps_algorithmic": alg_bytes / ms / 1e6, "frac_of_measured_hbm_peak": alg_bytes / ms / 1e6 / PEAK,
            unit_name + "_per_s": units / (ms * 1e-3)}
    line.update(extra)
    print(json.dumps(line), flush=True)


def c1():
    H, W, N = 180, 240, 200_000
    ev = make_events(1, N, 0.0, H, W, 1001, 300_000)
    out = {}
    ms = timeit(lambda: ep.bin_events(ev, (H, W), num_bins=5, count_channels=2, out=out), reps=100)
    report("C1 single 240x180 sample, 200k events -> 2-ch count + 5-bin voxel (fused, one call)", ms, 13 * N + 4 * 7 * H * W, N, "events",
           note="latency-bound by construction (3.8 MB)")
    # CPU oracle port beside it (1 core)
    from oracle import events as oe
    s = np.stack([ev.x.cpu().numpy(), ev.y.cpu().numpy(), ev.t.cpu().numpy() / 1e6, ev.p.cpu().numpy()], 1).astype(np.float64)
    t0 = time.perf_counter(); oe.voxel_grid(s, 5, (H, W)); oe.count_frame(s, (H, W), 2); cpu_ms = (time.perf_counter() - t0) * 1e3
    print(json.dumps({"config": "C1 CPU port (1 core)", "ms": cpu_ms, "events_per_s": N / cpu_ms * 1e3}), flush=True)


def c3():
    B, C, L, K, D, p = 128, 5, 196, 49, 384, 16
    torch.manual_seed(3000)
    x = torch.randn(B, C, 224, 224, device=dev)
    noise = torch.rand(B, L, device=dev)
    tokens = torch.randn(B, L, D, device=dev)
    pos = torch.randn(L, D, device=dev)
    sub = torch.randn(B, 1, 224, 224, device=dev)
    ids = {}

    def f_mask():
        ids["k"], ids["m"], ids["r"] = ep.mask_from_noise(noise, K)
    ms = timeit(f_mask)
    report("C3 mask_from_noise B=128 L=196 keep=49", ms, B * (4 * L + 8 * K + 12 * L), B, "samples", launches=1)
    ms = timeit(lambda: ep.gather_tokens(tokens, ids["k"], pos))
    report("C3 token gather (+pos) (128,196,384)->(128,49,384)", ms, B * 2 * 4 * K * D + 4 * L * D, B, "samples", launches=1)
    ms = timeit(lambda: ep.patchify_gather(x, p, ids["k"], "cpq"))
    report("C3 pre-embed patch gather (128,5,224,224)->(128,49,1280)", ms, B * 2 * 4 * K * C * p * p, B, "samples", launches=1)
    ms = timeit(lambda: ep.target_normpix(sub, p))
    report("C3 target patchify+norm_pix (128,1,224,224)->(128,196,256)", ms, B * 2 * 4 * 224 * 224, B, "samples", launches=1)

    def pipeline():
        k, m, r = ep.mask_from_noise(noise, K)
        ep.patchify_gather(x, p, k, "cpq")
        ep.target_normpix(sub, p)
    ms = timeit(pipeline)
    report("C3 input pipeline: mask + patch gather + target (3 launches)", ms,
           B * (4 * L + 8 * K + 12 * L + 2 * 4 * K * C * p * p + 2 * 4 * 224 * 224), B, "samples", launches=3)
    pipe = ep.MaskedInputPipeline(B, C, (224, 224), p, 0.75, dev)
    pipe.noise.copy_(noise); pipe.x.copy_(x); pipe.sub_frame.copy_(sub)
    ms = timeit(lambda: pipe.run(draw_noise=False))
    report("C3 input pipeline as one CUDA graph (mask + patch gather + target)", ms,
           B * (4 * L + 8 * K + 12 * L + 2 * 4 * K * C * p * p + 2 * 4 * 224 * 224), B, "samples", launches=3)
    # reference formulation in stock torch on the same GPU, for scale
    def torch_ref():
        ids_shuffle = torch.argsort(noise, dim=1)
        ids_restore = torch.argsort(ids_shuffle, dim=1)
        k = ids_shuffle[:, :K]
        torch.gather(tokens + pos, 1, k.unsqueeze(-1).repeat(1, 1, D))
        t = sub.reshape(B, 1, 14, p, 14, p)
        t = torch.einsum("bchpwq->bhwpqc", t).reshape(B, L, p * p)
        (t - t.mean(-1, keepdim=True)) / (t.var(-1, keepdim=True) + 1e-6) ** .5
    print(json.dumps({"config": "C3 stock torch ops on the same GPU (argsort x2, gather, einsum patchify, mean/var)",
                      "ms": timeit(torch_ref)}), flush=True)


def c4():
    B, H, W, bins, N = 32, 440, 640, 15, 2_000_000
    ev = make_events(B, N, 0.1, H, W, 4000, 100_000)
    out = {}
    ms = timeit(lambda: ep.bin_events(ev, (H, W), num_bins=bins, out=out), reps=10)
    report("C4 DSEC-shaped B=32, 640x440, 15 bins, ~2M events/sample: voxel", ms, 13 * ev.num_events + 4 * bins * H * W * B, ev.num_events, "events")
    ms = timeit(lambda: ep.time_surface(ev, (H, W), tau=0.03), reps=10)
    report("C4 time surface (2,440,640) [unpinned]", ms, 13 * ev.num_events + 4 * 2 * H * W * B, ev.num_events, "events")
    ms = timeit(lambda: ep.evrep(ev, (H, W)), reps=5, warm=2)
    report("C4 EvRep (3,440,640) f64 [pinned time-surface stand-in]", ms, 13 * ev.num_events + 12 * H * W * B, ev.num_events, "events")


def c5():
    B, H, W, bins, N = 512, 260, 346, 9, 100_000
    ev = make_events(B, N, 0.5, H, W, 5000, 50_000)
    out = {}
    ms = timeit(lambda: ep.bin_events(ev, (H, W), num_bins=bins, out=out), reps=10)
    report("C5 MVSEC-shaped B=512, 346x260, 9 bins, ~100k events/sample: voxel (org grid)", ms,
           13 * ev.num_events + 4 * bins * H * W * B, ev.num_events, "events")
    mask = (torch.rand(B, 196, device=dev) < 0.75).float()
    ms = timeit(lambda: ep.convvit_keep_masks(mask))
    report("C5 ConvViT block masks (512,196)->(512,1,56,56),(512,1,28,28)", ms, B * 4 * (196 + 56 * 56 + 28 * 28), B, "samples", launches=2)
    xs = torch.randn(B, 3136, 96, device=dev)
    m49 = torch.zeros(B, 49, device=dev)
    m49[:, torch.randperm(49, device=dev)[:37]] = 1
    ms = timeit(lambda: ep.swin_apply_mask(xs, m49, (56, 56), n_vis=12 * 64))
    report("C5 Swin apply_mask (512,3136,96)->(512,768,96)", ms, B * 2 * 4 * 768 * 96, B, "samples", launches=2)


def f_rows():
    """SURVEY.md §8(f) "next" rows that are built: f1 view augmentation, f3 decoder un-shuffle + fused patch loss, f4 Swin grouping."""
    from types import SimpleNamespace
    from eventpretrain_b200.view_augment import ViewChoice
    B, C, H, W = 256, 5, 480, 640
    # a8: frame-side difference map at the C2 batch (two frames in, one map out; SURVEY 8(d): 4*H*W*(inputs+1) bytes per sample)
    f0 = torch.rand(B, 1, H, W, device=dev) + 0.1
    f1 = torch.rand(B, 1, H, W, device=dev) + 0.1
    flips = [int(i % 2) for i in range(B)]
    for mode in ("linear", "log"):
        ms = timeit(lambda: ep.diffmap_frames(f0, f1, mode, negate=flips), reps=20)
        report(f"a8 diff-map target from frames (256,1,480,640), {mode}, per-sample sign flip", ms, B * 4 * H * W * 3, B, "samples", launches=1)
    del f0, f1
    x = torch.randn(B, C, H, W, device=dev)
    rng = np.random.default_rng(7)
    choices = []
    for _ in range(B):
        cw, ch = int(rng.integers(500, 640)), int(rng.integers(380, 480))
        choices.append(ViewChoice(int(rng.integers(0, W - cw + 1)), int(rng.integers(0, H - ch + 1)), cw, ch, bool(rng.integers(0, 2)),
                                  bool(rng.integers(0, 2)), bool(rng.integers(0, 2))))
    touched = sum(c.crop_w * c.crop_h for c in choices) * C * 4
    for mode in ("nearest", "bilinear"):
        ms = timeit(lambda: ep.apply_views(x, choices, (224, 224), mode), reps=20)
        report(f"f1 evg_augment batch (256,5,480,640) -> crop -> {mode} 224x224 -> flips (one launch)", ms,
               (touched if mode != "nearest" else B * C * 224 * 224 * 4) + B * C * 224 * 224 * 4, B, "samples", launches=1)
    fr = torch.randn(B, 1, H, W, device=dev)
    ms = timeit(lambda: ep.apply_views(fr, choices, (224, 224), "bicubic"), reps=20)
    report("f1 frame_augment batch (256,1,480,640) -> crop -> bicubic 224x224 -> flip/negate", ms,
           touched // C + B * 224 * 224 * 4, B, "samples", launches=1)
    Bm, L, K, D = 128, 196, 49, 256
    emb = torch.randn(Bm, K, D, device=dev)
    mt = torch.randn(1, 1, D, device=dev)
    pos = torch.randn(1, L, D, device=dev)
    ids_restore = torch.argsort(torch.rand(Bm, L, device=dev), 1)
    ms = timeit(lambda: ep.unshuffle_tokens(emb, mt, ids_restore, pos))
    report("f3 decoder un-shuffle (128,49,256) + mask token + pos -> (128,196,256)", ms, Bm * 4 * (K * D + L * D) + 4 * L * D, Bm, "samples", launches=1)
    pred = torch.randn(Bm, L, 256, device=dev)
    sub = torch.randn(Bm, 1, 224, 224, device=dev)
    ms = timeit(lambda: ep.target_patch_loss(pred, sub, 16))
    report("f3 fused target (patchify + norm_pix) + per-patch MSE (128,196,256)", ms, Bm * 4 * (224 * 224 + L * 256 + L), Bm, "samples", launches=1)
    # f4: grouping plan of a fresh mask (host DP in native code + torch index ops), then the cached call
    keep = 24
    m = torch.zeros(49)
    m[torch.randperm(49)[:49 - keep]] = 1
    mm = m.reshape(7, 7)[:, None, :, None].expand(7, 8, 7, 8).reshape(-1).bool()
    ii, jj = torch.meshgrid(torch.arange(56), torch.arange(56), indexing="ij")
    coords = torch.stack([ii, jj], -1).reshape(1, -1, 2)[:, ~mm].to(dev)
    t0 = time.perf_counter()
    for i in range(20):
        ep.GroupingModule._cache.clear()
        ep.GroupingModule(7, 3).prepare(coords, coords.shape[1])
    torch.cuda.synchronize()
    fresh = (time.perf_counter() - t0) / 20 * 1e3
    t0 = time.perf_counter()
    for i in range(200):
        ep.GroupingModule(7, 3).prepare(coords, coords.shape[1])
    cached = (time.perf_counter() - t0) / 200 * 1e3
    wt = [24] * 48 + [12] * 16
    t0 = time.perf_counter()
    for i in range(200):
        ep.group_windows(49, wt)
    dp = (time.perf_counter() - t0) / 200 * 1e3
    print(json.dumps({"config": "f4 Swin GroupingModule.prepare, 1536 visible tokens, window 7 shift 3 (host wall time)",
                      "ms_fresh_mask": fresh, "ms_cached_mask": cached, "ms_native_group_windows_64_windows": dp}), flush=True)


if __name__ == "__main__":
    for f in (c1, c3, c4, c5, f_rows):
        f()
        torch.cuda.empty_cache()
