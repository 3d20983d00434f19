#!/usr/bin/env python
"""Benchmark of the EventPretrain input hot path on B200 (contract: see the task statement / DESIGN.md section 6).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload at every N (weak scaling: the same per-GPU batch on each rank): BASELINE.json configs[1] —
an N-ImageNet-shaped ragged batch of 256 samples, 640x480 sensor, ~1M events/sample (+-20 %), time-sorted microsecond stamps
uniform in a 0.3 s window, binned at sensor resolution into a 5-bin voxel grid plus the event-side
difference-map target voxel.sum(0) (the variant that also emits the per-channel batch statistics feeding the path's one
collective is timed under extra.with_batch_statistics).  The batch is
resident in the layout the collate step ships over PCIe: the densest lossless transport layout the stream fits
(RaggedEvents.transport(): here 4 B/event; results bit-identical to the 13 B/event canonical SoA, whose number is under "extra").
A "step" is one pass of the hot path over the batch.  Prints ONE JSON line on rank 0.

  value     events/s with the batch already resident in HBM (CUDA events on the launching stream, max over ranks)
  roofline  SURVEY 8(d) algorithmic bytes of a step / the device time of the step — the same quotient at every N —
            against the measured HBM copy peak; frac_actual_layout_bytes = the same with the bytes the kernels really have
            to move in the resident layout; traffic = DRAM bytes per step from the committed ncu capture (profiles/)
  e2e       the same metric through the public API from pinned HOST buffers in the transport layout: H2D of the batch in
            slices (copy of slice i+1 overlaps the binning of slice i) and a D2H read of a per-sample checksum inside the
            timed region; the tensors themselves stay on the device, where the encoder consumes them.  Sub-entries:
            from_reference_format = the whole host side too, starting from the reference's own per-sample (N,4) float64 arrays
            (threaded one-pass collate into the transport layout of slice i+1 overlapping H2D + binning of slice i; bounded
            sample, rate-normalised);
            h2d_only = the same buffers copied with no kernels (the ceiling the interconnect sets at this N)
  cpu_baseline  the oracle C port of the reference routine on the host cores, bounded sample, rank 0 / N=1 only
  extra.reference_res_224  the reference's own pre-training order (events rescaled to 224x224, then binned) on the three kernel
            families (whole-plane kernels by default), with a bit-identity check
  extra.configs  BASELINE.json configs[0], [2], [3], [4] at their own sizes (per-GPU share at N > 1), each with
            throughput, SURVEY 8(d) bytes, roofline fraction and a parity check against the oracle

--impl reference times that CPU port alone (the reference is pure Python and /root/reference does not exist
on the GPU box; its arithmetic is restated in oracle/ep_oracle.c and pinned bit-exact to the reference).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

H, W, BINS = 480, 640, 5
BATCH, MEAN_EVENTS = 256, 1_000_000
WORKLOAD = "N-ImageNet-shaped ragged batch 256 x ~1M events, 640x480 -> 5-bin voxel grid + voxel.sum(0) diff-map target (sensor-res)"
BYTES_PER_EVENT = 13
WINDOW_US = 300_000      # stamps uniform in [0, 0.3 s)


def counts_for(rank, batch=BATCH, mean=MEAN_EVENTS, spread=0.2):
    rng = np.random.default_rng(2000 + rank)
    return np.maximum(1, np.round(mean * rng.uniform(1 - spread, 1 + spread, batch))).astype(np.int64)


def algorithmic_bytes(n_events, batch, h=H, w=W, planes=BINS + 1):
    # 13 B/event read once + every output element written once (voxel bins + the sum plane)
    return BYTES_PER_EVENT * n_events + 4 * planes * h * w * batch


def make_batch_gpu(rank, device, skewed=False, batch=BATCH, mean=MEAN_EVENTS, window_us=WINDOW_US, size=(H, W), spread=0.2,
                   bursty=False, seed=2000):
    """Synthetic streams generated on the device: uniform pixels (or the skewed mix: 70 % on 64 line segments,
    0.1 % on 32 hot pixels), time-sorted int64 microsecond stamps uniform in a window (0.3 s) or bursty
    (90 % of the events inside ten 3 ms bursts, the rest spread over the window), p in {0,1}."""
    import torch
    import eventpretrain_b200 as ep
    h, w = size
    counts = counts_for(rank + seed - 2000, batch, mean, spread)
    off = np.zeros(batch + 1, np.int64)
    np.cumsum(counts, out=off[1:])
    n = int(off[-1])
    g = torch.Generator(device=device).manual_seed(seed + rank)
    x = torch.randint(0, w, (n,), device=device, generator=g, dtype=torch.int32)
    y = torch.randint(0, h, (n,), device=device, generator=g, dtype=torch.int32)
    if skewed:
        u = torch.rand(n, device=device, generator=g)
        seg = torch.randint(0, 64, (n,), device=device, generator=g)
        gs = torch.Generator(device=device).manual_seed(99)
        ends = torch.rand(64, 4, device=device, generator=gs) * torch.tensor([w - 1, h - 1, w - 1, h - 1], device=device)
        a = torch.rand(n, device=device, generator=g)
        lx = ends[seg, 0] + a * (ends[seg, 2] - ends[seg, 0]) + 1.5 * torch.randn(n, device=device, generator=g)
        ly = ends[seg, 1] + a * (ends[seg, 3] - ends[seg, 1]) + 1.5 * torch.randn(n, device=device, generator=g)
        on_line = u < 0.7
        x = torch.where(on_line, lx.round().clamp_(0, w - 1).int(), x)
        y = torch.where(on_line, ly.round().clamp_(0, h - 1).int(), y)
        hot = torch.randint(0, 32, (n,), device=device, generator=g)
        hp = torch.randint(0, w * h, (32,), device=device, generator=gs)
        is_hot = u > 0.999
        x = torch.where(is_hot, (hp[hot] % w).int(), x)
        y = torch.where(is_hot, (hp[hot] // w).int(), y)
        del u, seg, a, lx, ly, hot
    p = torch.randint(0, 2, (n,), device=device, generator=g, dtype=torch.uint8)
    t = torch.empty(n, dtype=torch.int64, device=device)
    for b in range(batch):
        lo, hi = int(off[b]), int(off[b + 1])
        tt = torch.randint(0, window_us, (hi - lo,), device=device, generator=g)
        if bursty:
            k = torch.randint(0, 10, (hi - lo,), device=device, generator=g)
            inside = torch.rand(hi - lo, device=device, generator=g) < 0.9
            tt = torch.where(inside, k * (window_us // 10) + tt % 3000, tt)
        t[lo:hi] = torch.sort(tt).values
    ev = ep.RaggedEvents(x.to(torch.uint16), y.to(torch.uint16), t, p, torch.from_numpy(off).to(device), off, t_div=1e6)
    return ev


class ClockSampler(threading.Thread):
    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.proc = index, [], None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.FIELDS}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE, text=True)
            for line in self.proc.stdout:
                self.rows.append([c.strip() for c in line.split(",")])
        except Exception:
            pass

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm, mx, reasons = [], 0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = max(mx, float(r[1]))
                for nme, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(nme)
            except (ValueError, IndexError):
                continue
        top = sorted(sm)[len(sm) // 2:] or [0.0]   # upper half = samples under load
        return {"sm_mhz": float(np.median(top)), "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def measured_peak():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic():
    """DRAM bytes per step of the headline layout from the committed ncu capture (--cache-control none), or None."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
    except Exception:
        return None


def aos_sample(ev_host_soa, off, b):
    lo, hi = int(off[b]), int(off[b + 1])
    x, y, t, p = ev_host_soa
    out = np.empty((hi - lo, 4), np.float64)
    out[:, 0] = x[lo:hi]; out[:, 1] = y[lo:hi]; out[:, 2] = t[lo:hi]; out[:, 2] /= 1e6; out[:, 3] = p[lo:hi]
    return out


def cpu_port_setup(n_samples, rank=0):
    """The same synthetic workload (same generator family), host side, as the (N,4) float64 arrays the reference
    functions take; bounded to n_samples samples."""
    counts = counts_for(rank)[:n_samples]
    rng = np.random.default_rng(7000 + rank)
    samples = []
    for n in counts:
        n = int(n)
        samples.append(np.stack([rng.integers(0, W, n), rng.integers(0, H, n),
                                 np.sort(rng.integers(0, WINDOW_US, n)) / 1e6, rng.integers(0, 2, n)], 1).astype(np.float64))
    off = np.cumsum([0] + [len(s) for s in samples]).astype(np.int64)
    return np.ascontiguousarray(np.concatenate(samples, 0)), off


def cpu_port_step(ev, off, threads):
    from oracle import events as oe
    t0 = time.perf_counter()
    vox = oe.voxel_grid_batch(ev, off, BINS, (H, W), num_threads=threads)
    vox.sum(axis=1, keepdims=True)          # voxel.sum(0)[None] per sample
    return time.perf_counter() - t0


def run_reference(args):
    """--impl reference: the CPU port of the reference routine on all host cores, bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import events as oe
    oe.build()
    threads = os.cpu_count() or 1
    n_samples = min(BATCH, max(8, min(threads, 64)))
    ev, off = cpu_port_setup(n_samples)
    for _ in range(args.warmup):
        cpu_port_step(ev, off, threads)
    times = [cpu_port_step(ev, off, threads) for _ in range(args.steps)]
    total = float(np.sum(times))
    value = int(off[-1]) * args.steps / total / 1e9
    sample = f"{n_samples} of {BATCH} samples ({int(off[-1])} events) per step"
    line = {"impl": "reference", "metric": "events_binned_per_s", "value": value, "unit": "Gevents/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64 time / f32 sequential accumulate (reference arithmetic)",
            "data": "synthetic", "config": {"workload": WORKLOAD, "sample": sample},
            "cpu_baseline": {"value": value, "unit": "Gevents/s", "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "Gevents/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


def close(a, b, rel=1e-5, abs_=1e-6):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return bool(np.all(np.abs(a - b) <= rel * np.abs(b) + abs_))


def layout_name_of(host):
    bpe = host.nbytes() / max(host.num_events, 1)
    name = ("canonical SoA: x,y u16 | t i64 ticks | p u8" if host.t_base is None else
            "compact SoA: x,y u16 | u32 ticks relative to the sample | polarity << 31" if host.y is not None else
            "packed SoA: u32 x | y << 11 | p << 22 | ticks << 23 (9-bit ticks relative to a base per 256 events)" if host.t is None else
            "packed SoA: u32 x | y << 11 | p << 22 | ticks >> 8 << 23 + u8 ticks & 0xff (ticks relative to a base per 1024 events)")
    return f"{name} ({bpe:.2f} B/event)", bpe


def run_ours(args):
    import torch
    import torch.distributed as dist
    import eventpretrain_b200 as ep
    from eventpretrain_b200 import _lib
    from eventpretrain_b200 import dist as epd
    from eventpretrain_b200.dist import init_from_env

    rank, world, local = init_from_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device; there is no CPU fallback (use --impl reference for the CPU port)")
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    L = ep.load_library()
    cores = os.cpu_count() or 1
    host_threads = max(1, cores // world)            # host-side work of this rank (collate / pack)
    peak, peak_src = measured_peak()

    def timed_ms(fn, steps, warm=3):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        for _ in range(steps):
            fn()
        s1.record()
        torch.cuda.synchronize()
        return s0.elapsed_time(s1) / steps

    def allmax(v):
        t_ = torch.tensor([v], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t_, op=dist.ReduceOp.MAX)
        return float(t_.item())

    def allsum(v):
        t_ = torch.tensor([v], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t_, op=dist.ReduceOp.SUM)
        return float(t_.item())

    def to_host(e):
        return ep.RaggedEvents(e.x.cpu(), e.y.cpu(), e.t.cpu(), e.p.cpu(), e.offsets.cpu(), e.offsets_host, e.t_div)

    ev13 = make_batch_gpu(rank, dev)
    n_events = ev13.num_events
    host13 = to_host(ev13)
    host = host13.transport(threads=host_threads).pin_memory()
    ev = host.to(dev)
    layout_name, bpe = layout_name_of(host)
    layout_name += "; the lossless transport layout, results bit-identical to the 13 B/event canonical SoA"
    host_gb = host.nbytes() / 1e9
    torch.cuda.synchronize()
    out = {"voxel": torch.empty((BATCH, BINS, H, W), dtype=torch.float32, device=dev),
           "voxel_sum": torch.empty((BATCH, 1, H, W), dtype=torch.float32, device=dev),
           "stats": torch.zeros((BINS + 1, 4), dtype=torch.float64, device=dev)}
    reduced = torch.zeros((BINS + 1, 4), dtype=torch.float64, device=dev)
    side = torch.cuda.Stream(device=dev)

    def step(comm=True):
        # the hot path: events -> voxel grid + voxel.sum(0), one C-ABI call (no collective: samples are independent)
        ep.bin_events(ev, (H, W), num_bins=BINS, voxel_sum=True, out=out, method=args.method)

    def stats_step(comm=True):
        # the same call with the per-channel batch statistics as a by-product of the sweep's flush, and (N > 1) the path's only
        # collective: the small all-reduce of those statistics (SUM of count / sum / sum of squares, MAX of max) on a side
        # stream, so that it never gates the next binning call
        ep.bin_events(ev, (H, W), num_bins=BINS, voxel_sum=True, out=out, method=args.method, stats=True)
        if world > 1 and comm:
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                reduced.copy_(out["stats"])
                epd.allreduce_statistics(reduced)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()

    # ---- timed region: K steps, CUDA events on the launching stream, clocks sampled meanwhile ----------------
    sampler = ClockSampler(local)
    sampler.start()
    time.sleep(0.3)
    launches0 = L.ep_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    ms_total = e0.elapsed_time(e1)
    launches = int(L.ep_launch_count() - launches0)
    # keep the GPU busy a little longer so the 100 ms clock sampler sees the loaded state
    t_end = time.time() + 1.0
    while time.time() < t_end:
        step(comm=False)          # time-based loop: no collectives here (iteration counts differ between ranks)
    torch.cuda.synchronize()
    clocks = sampler.stop()
    ms_total = allmax(ms_total)
    total_events = int(allsum(n_events))
    value = total_events * args.steps / (ms_total * 1e-3) / 1e9
    ms_step = ms_total / args.steps

    # the statistics variant of the step (+ its all-reduce at N > 1), timed the same way
    barrier()
    ms_stats = allmax(timed_ms(stats_step, args.steps, warm=2))
    ms_stats_no_comm = None
    if world > 1:
        barrier()
        ms_stats_no_comm = allmax(timed_ms(lambda: stats_step(comm=False), args.steps, warm=1))
        torch.cuda.synchronize()

    # ---- roofline: the same quotient at every N: algorithmic bytes of one rank's step / device time of the step ----
    L.ep_profile_enable(1)
    for _ in range(args.steps):
        step(comm=False)
    prof = _lib.ProfileStats()
    L.ep_profile_read(prof)
    L.ep_profile_enable(0)
    route_ms, sweep_ms, other_ms = (prof.ms[i] / args.steps for i in range(3))
    alg = algorithmic_bytes(n_events, BATCH)
    actual = host.nbytes() + 4 * (BINS + 1) * H * W * BATCH
    achieved = alg / (ms_step * 1e-3) / 1e9
    traffic = ncu_traffic()
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": None if traffic is None else traffic.get("dram_bytes_per_step"),
                "traffic_source": None if traffic is None else traffic.get("how"),
                "peak_source": peak_src, "frac_of_8TBs_spec": achieved / 8000.0,
                "frac_actual_layout_bytes": actual / (ms_step * 1e-3) / 1e9 / peak,
                "actual_layout_bytes_per_step": actual,
                "denominator": "device time of the whole step (CUDA events around the K timed steps), the same rule at every N",
                "kernels": {"k_route_ms_per_step": route_ms, "k_sweep_ms_per_step": sweep_ms, "setup_ms_per_step": other_ms,
                            "note": "each kernel timed alone with CUDA events around its launch (library instrumentation, profiling run "
                                    "after the timed region); they run back to back, their sum is the step"},
                "algorithmic_bytes_per_step": alg,
                "algorithmic_bytes_rule": f"SURVEY.md 8(d): 13 B/event canonical record + 4 B per output element, whatever the resident layout ({bpe:.1f} B/event here)"}

    extra = {}
    extra["with_batch_statistics"] = {
        "what": "ep_bin_events_stats: the same step plus per-channel (count, sum, sum of squares, max) of the six output channels as a "
                "by-product of the sweep's flush (bit-reproducible), feeding the path's one collective: the (6,4) fp64 all-reduce "
                "(NCCL SUM + MAX) on a side stream at N > 1",
        "ms_per_step": ms_stats, "Gevents_per_s": total_events / (ms_stats * 1e-3) / 1e9,
        "ms_per_step_without_the_allreduce": ms_stats_no_comm}

    def gev(evx, method, size=(H, W), scale=(1.0, 1.0), o=None):
        o = out if o is None else o
        ms = timed_ms(lambda: ep.bin_events(evx, size, num_bins=BINS, voxel_sum=True, out=o, method=method, scale=scale), args.steps)
        return evx.num_events / (ms * 1e-3) / 1e9, ms

    # ---- the same batch in the other layouts / kernel family, on the skewed and on a bursty stream, and binned the way the
    #      reference's own pre-training does it (events_reshape to 224x224 fused, pr_n_imagenet_dataset.py:85-87) ----
    if world == 1:
        lay = {}
        lay["packed_4B_tiled_path_Gevents_per_s (the headline step again)"] = gev(ev, args.method)[0]
        lay["packed_4B_global_RED_path_Gevents_per_s"] = gev(ev, "global")[0]
        lay["canonical_13B_layout_Gevents_per_s (global-RED path)"] = gev(ev13, args.method)[0]
        ev8 = host13.compact(threads=host_threads).to(dev)
        lay["compact_8B_layout_Gevents_per_s (global-RED path)"] = gev(ev8, args.method)[0]
        del ev8
        extra["layouts"] = lay
        o224 = {"voxel": torch.empty((BATCH, BINS, 224, 224), dtype=torch.float32, device=dev),
                "voxel_sum": torch.empty((BATCH, 1, 224, 224), dtype=torch.float32, device=dev)}
        g224, ms224 = gev(ev, args.method, (224, 224), (224 / W, 224 / H), o224)
        alg224 = algorithmic_bytes(n_events, BATCH, 224, 224)
        extra["reference_res_224"] = {"what": "the reference's pre-training order: events_reshape 640x480 -> 224x224 fused, then the voxel grid "
                                              "+ voxel.sum(0) (dataset/pretrain/pr_n_imagenet_dataset.py:85-87)",
                                      "Gevents_per_s": g224, "ms_per_step": ms224, "algorithmic_bytes_per_step": alg224,
                                      "roofline_frac": alg224 / (ms224 * 1e-3) / 1e9 / peak,
                                      "kernels": "whole-plane path (k_plane_bounds + k_plane): a 224x224 plane fits one SM's shared memory, so a "
                                                 "CTA owns one output plane of one sample and streams the two time-contiguous event slices "
                                                 "that feed it; no route pass, no routed records (DESIGN.md 3.0)"}
        ref224 = {k: v.clone() for k, v in o224.items()}
        r224 = extra["reference_res_224"]
        r224["route_sweep_path_Gevents_per_s"] = gev(ev, "tiled", (224, 224), (224 / W, 224 / H), o224)[0]
        same = all(torch.equal(ref224[k], o224[k]) for k in ref224)
        r224["global_RED_path_Gevents_per_s"] = gev(ev, "global", (224, 224), (224 / W, 224 / H), o224)[0]
        r224["bit_identical_across_the_three_paths"] = bool(same and all(torch.equal(ref224[k], o224[k]) for k in ref224))
        ms224s = timed_ms(lambda: ep.bin_events(ev, (224, 224), num_bins=BINS, voxel_sum=True, out=o224, method=args.method,
                                                scale=(224 / W, 224 / H), stats=True), args.steps)
        r224["with_batch_statistics_ms_per_step"] = ms224s
        del o224, ref224
    del ev13
    torch.cuda.empty_cache()
    if world == 1:
        sk13 = make_batch_gpu(rank, dev, skewed=True)
        skh = to_host(sk13)
        del sk13
        skt = skh.transport(threads=host_threads)
        sk = skt.to(dev)
        extra["skewed_distribution"] = {"what": "70 % of the events on 64 line segments, 0.1 % on 32 hot pixels", "layout": layout_name_of(skt)[0],
                                        "Gevents_per_s": gev(sk, args.method)[0]}
        del sk, skh, skt
        torch.cuda.empty_cache()
        bu13 = make_batch_gpu(rank, dev, bursty=True, batch=64)
        buh = to_host(bu13)
        del bu13
        but = buh.transport(threads=host_threads)
        bu = but.to(dev)
        o64 = {"voxel": out["voxel"][:64], "voxel_sum": out["voxel_sum"][:64]}
        extra["bursty_stream"] = {"what": "64 samples; 90 % of each sample's events inside ten 3 ms bursts, the rest spread over 0.3 s "
                                          "(quiet stretches of ~0.4 Mevents/s: a 256-event block then spans more than 2^9 ticks)",
                                  "layout_chosen_by_transport": layout_name_of(but)[0], "Gevents_per_s": gev(bu, args.method, o=o64)[0]}
        del bu, buh, but
        torch.cuda.empty_cache()

    # ---- e2e: pinned host transport buffers -> H2D -> bin -> D2H of the per-sample checksum, all inside the timed region ----
    # The batch crosses PCIe in E2E_SLICES slices of consecutive samples: the copy of slice i+1 (copy stream) overlaps the binning
    # of slice i (compute stream); two device staging sets, guarded by events.
    del ev
    torch.cuda.empty_cache()
    E2E_SLICES = 8
    bounds = [(BATCH * i) // E2E_SLICES for i in range(E2E_SLICES + 1)]
    # every slice is packed on its own (the collate step would produce them like this): block offsets start at 0
    slices = [host13.take(bounds[i], bounds[i + 1]).transport(threads=host_threads).pin_memory() for i in range(E2E_SLICES)]
    if any((sl.t is None) != (slices[0].t is None) or (sl.y is None) != (slices[0].y is None) for sl in slices):
        slices = [host13.take(bounds[i], bounds[i + 1]).packed(5).pin_memory() for i in range(E2E_SLICES)]
    h2d = sum(sl.nbytes() for sl in slices)
    fields = tuple(f for f in ("x", "y", "t", "p", "offsets", "t_base") if getattr(slices[0], f) is not None)
    stage = [{f: torch.empty(int(1.05 * max(getattr(sl, f).numel() for sl in slices)) + 64, dtype=getattr(slices[0], f).dtype, device=dev)
              for f in fields} for _ in range(2)]
    copy_stream, comp_stream = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
    copied = [torch.cuda.Event() for _ in range(E2E_SLICES)]
    binned = [torch.cuda.Event() for _ in range(E2E_SLICES)]

    def upload(i, sl, kernels=True):
        st_ = stage[i % 2]
        with torch.cuda.stream(copy_stream):
            if i >= 2:
                copy_stream.wait_event(binned[i - 2])                 # the staging set is free again
            view = {}
            for f in fields:
                src = getattr(sl, f)
                view[f] = st_[f][:src.numel()]
                view[f].copy_(src, non_blocking=True)
            copied[i].record(copy_stream)
        with torch.cuda.stream(comp_stream):
            comp_stream.wait_event(copied[i])
            if kernels:
                d = ep.RaggedEvents(view["x"], view.get("y"), view.get("t"), view.get("p"), view["offsets"], sl.offsets_host, sl.t_div,
                                    view["t_base"])
                o = {"voxel": out["voxel"][bounds[i]:bounds[i + 1]], "voxel_sum": out["voxel_sum"][bounds[i]:bounds[i + 1]]}
                ep.bin_events(d, (H, W), num_bins=BINS, voxel_sum=True, out=o, method=args.method)
            binned[i].record(comp_stream)

    def e2e_step(kernels=True):
        for i, sl in enumerate(slices):
            upload(i, sl, kernels)
        with torch.cuda.stream(comp_stream):
            chk_ = out["voxel_sum"].sum(dim=(1, 2, 3)).cpu() if kernels else None   # (B,) fp32 per-sample checksum (synchronises)
        if not kernels:
            comp_stream.synchronize()
        return chk_

    def wall(fn, steps):
        r_ = None
        for _ in range(2):
            fn()
        barrier()
        t0_ = time.perf_counter()
        for _ in range(steps):
            r_ = fn()
        barrier()
        return allmax(time.perf_counter() - t0_) / steps, r_

    e2e_s, chk = wall(e2e_step, args.steps)
    h2d_s, _ = wall(lambda: e2e_step(kernels=False), args.steps)
    e2e = {"value": total_events / e2e_s / 1e9, "unit": "Gevents/s", "h2d_bytes_per_step": h2d,
           "d2h_bytes_per_step": int(chk.numel() * chk.element_size()), "ms_per_step": 1e3 * e2e_s,
           "host_layout": "pinned " + layout_name + "; include/eventpretrain_b200.h",
           "pipelining": f"{E2E_SLICES} slices of consecutive samples; H2D of slice i+1 overlaps the binning of slice i",
           "outputs": "the voxel grids and sum planes (1.9 GB per step and GPU) stay on the device, where the encoder consumes them "
                      "(the reference's trainer moves them there, pr_trainer.py:27-28); the D2H read is a per-sample checksum of the result",
           "h2d_only": {"Gevents_per_s": total_events / h2d_s / 1e9, "ms_per_step": 1e3 * h2d_s, "GBps_all_ranks": world * h2d / h2d_s / 1e9,
                        "what": "the same pinned buffers copied with no kernels: the ceiling the host interconnect sets at this N"},
           "frac_of_h2d_only_ceiling": h2d_s / e2e_s}

    # ---- e2e from the reference's own input format: per-sample (N,4) float64 arrays -> threaded collate + pack (slice i+1 on
    #      worker threads) -> H2D -> bin; bounded sample (REF_SAMPLES of the batch), rate-normalised ----
    REF_SAMPLES, REF_SLICE = 64, 8
    hx, hy, ht, hp = (a.numpy() for a in (host13.x, host13.y, host13.t, host13.p))
    off13 = host13.offsets_host
    aos = [aos_sample((hx, hy, ht, hp), off13, b) for b in range(REF_SAMPLES)]
    n_ref = int(off13[REF_SAMPLES] - off13[0])
    ref_bounds = list(range(0, REF_SAMPLES + 1, REF_SLICE))
    n_ref_slices = len(ref_bounds) - 1

    def host_side(i):
        # one pass over the rows straight into the 4 B layout (ep_collate_transport4_host; the two-step form is its fallback)
        return ep.collate_transport(aos[ref_bounds[i]:ref_bounds[i + 1]], 1e6, pin=True, threads=host_threads)

    def host_side_two_steps(i):
        return ep.collate_events(aos[ref_bounds[i]:ref_bounds[i + 1]], 1e6, pin=True, threads=host_threads).transport(threads=host_threads)

    pool = ThreadPoolExecutor(max_workers=2)

    def ref_step():
        futs = [pool.submit(host_side, 0)]
        if n_ref_slices > 1:
            futs.append(pool.submit(host_side, 1))
        for i in range(n_ref_slices):
            sl = futs[i].result()
            if i + 2 < n_ref_slices:
                futs.append(pool.submit(host_side, i + 2))
            d = sl.to(dev, non_blocking=True)
            o = {"voxel": out["voxel"][ref_bounds[i]:ref_bounds[i + 1]], "voxel_sum": out["voxel_sum"][ref_bounds[i]:ref_bounds[i + 1]]}
            ep.bin_events(d, (H, W), num_bins=BINS, voxel_sum=True, out=o, method=args.method)
        return out["voxel_sum"][:REF_SAMPLES].sum(dim=(1, 2, 3)).cpu()

    try:
        ref_s, _ = wall(ref_step, max(2, args.steps // 3))
        t0 = time.perf_counter()
        for i in range(n_ref_slices):
            host_side(i)
        host_only_s = time.perf_counter() - t0
        for i in range(n_ref_slices):                 # (once untimed: the pinned-buffer cache holds the one-pass shapes so far)
            host_side_two_steps(i)
        t0 = time.perf_counter()
        for i in range(n_ref_slices):
            host_side_two_steps(i)
        host_two_s = time.perf_counter() - t0
        e2e["from_reference_format"] = {
            "Gevents_per_s": world * n_ref / ref_s / 1e9, "ms_per_sample_step": 1e3 * ref_s,
            "what": "per-sample (N,4) float64 x,y,t,p arrays (the reference's event format, 32 B/event) -> ep.collate_transport (one "
                    "pass over the rows into the 4 B layout) -> H2D -> ep.bin_events; host work of slice i+1 / i+2 on worker threads "
                    "while slice i is copied and binned",
            "sample": f"first {REF_SAMPLES} of {BATCH} samples per rank ({n_ref} events), slices of {REF_SLICE}, rate-normalised",
            "host_threads_per_rank": host_threads, "host_side_alone_Gevents_per_s": world * n_ref / host_only_s / 1e9,
            "host_side_alone_two_step_form_Gevents_per_s (ep.collate_events + transport(): 62 B of host memory traffic per event against 36)": world * n_ref / host_two_s / 1e9,
            "note": "bounded by the host: 32 B/event of float64 rows have to be read from host memory before anything is shipped; "
                    "the pinned output buffers are allocated inside the timed host step"}
    except Exception as e:      # a side measurement: never at the expense of the result line
        e2e["from_reference_format"] = {"error": repr(e)}
    pool.shutdown()
    del aos

    # ---- CPU baseline beside it (rank 0, N=1): oracle port on the host cores, bounded sample + parity check ----
    cpu = None
    if rank == 0 and world == 1:
        from oracle import events as oe
        oe.build()
        e2e_step()                                   # the whole batch again (the reference-format leg rewrote the first samples)
        n_s = min(BATCH, max(8, min(cores, 64)))
        pick = [int(v) for v in np.linspace(0, BATCH - 1, n_s).round()]          # spread over the whole batch
        sam = [aos_sample((hx, hy, ht, hp), off13, b) for b in pick]
        aos_cat = np.ascontiguousarray(np.concatenate(sam, 0))
        soff = np.cumsum([0] + [len(s) for s in sam]).astype(np.int64)
        t0 = time.perf_counter()
        ref = oe.voxel_grid_batch(aos_cat, soff, BINS, (H, W), num_threads=cores)
        ref_sum = ref.sum(axis=1, keepdims=True)
        cpu_s = time.perf_counter() - t0
        idx = torch.tensor(pick, device=dev)
        got = out["voxel"][idx].cpu().numpy()
        err = np.abs(got - ref)
        ok = close(got, ref)
        sum_ok = close(out["voxel_sum"][idx].cpu().numpy(), ref_sum, 1e-5, 2e-6)
        sample = f"{n_s} of {BATCH} samples spread over the batch ({int(soff[-1])} events), one pass"
        cpu = {"value": int(soff[-1]) / cpu_s / 1e9, "unit": "Gevents/s", "cores": cores, "kind": "port", "sample": sample,
               "parity_vs_port": {"voxel_within_1e-5rel_1e-6abs": ok, "voxel_sum_ok": sum_ok, "max_abs_err": float(err.max())}}
        del ref, ref_sum, got, err, aos_cat, sam
    del hx, hy, ht, hp, host13, slices, stage
    out.clear()
    torch.cuda.empty_cache()

    # ---- the other BASELINE configs at their own sizes (per-GPU share), each with bytes, roofline fraction and a parity check ----
    try:
        extra["configs"] = run_configs(ep, torch, dev, rank, world, peak, allmax, allsum, timed_ms)
    except Exception as e:
        extra["configs"] = {"error": repr(e)}

    if rank == 0:
        line = {"metric": "events_binned_per_s", "value": value, "unit": "Gevents/s", "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "integer-tick time arithmetic, Q24 fixed-point int32/int64 accumulate -> f32",
                "data": "synthetic",
                "config": {"workload": WORKLOAD, "per_gpu_batch": BATCH, "events_per_gpu": n_events, "layout": layout_name,
                           "stamps": "time-sorted int64 microseconds, uniform in a 0.3 s window per sample (SURVEY.md 8d)",
                           "cache": f"inputs ({host_gb:.1f} GB/GPU) and outputs (1.9 GB/GPU) exceed the 126 MB L2; no flush needed",
                           "parallelism": f"shard-by-sample x{world}, no collective in the step (samples are independent); the statistics variant with its "
                                          "(6,4) fp64 all-reduce is timed under extra.with_batch_statistics",
                           "method": args.method, "step": "ep_bin_events: voxel grid + voxel.sum(0), one C-ABI call"},
                "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches, "clocks": clocks, "extra": extra}
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def run_configs(ep, torch, dev, rank, world, peak, allmax, allsum, timed_ms):
    """BASELINE.json configs[0], [2], [3], [4] (C1, C3, C4, C5), per-GPU share, weak scaling.  Values are whole-job
    (all ranks); ms is the slowest rank's; parity = rank 0's tensors against the CPU oracle on a bounded sample."""
    from oracle import events as oe
    from oracle import stage3_np as s3
    if rank == 0:
        oe.build()
    res = {}

    def entry(name, ms, alg_bytes, units, unit, parity, **kw):
        ms = allmax(ms)
        tot = allsum(units)
        e = {"ms": ms, unit + "_per_s": tot / (ms * 1e-3), "algorithmic_bytes_per_gpu": int(alg_bytes),
             "roofline_frac": alg_bytes / (ms * 1e-3) / 1e9 / peak, "parity": parity}
        e.update(kw)
        res[name] = e

    def sample_aos(evh, b):
        lo, hi = int(evh.offsets_host[b]), int(evh.offsets_host[b + 1])
        return np.stack([evh.x[lo:hi].numpy(), evh.y[lo:hi].numpy(), evh.t[lo:hi].numpy() / 1e6, evh.p[lo:hi].numpy()], 1).astype(np.float64)

    def host_of(e):
        return ep.RaggedEvents(e.x.cpu(), e.y.cpu(), e.t.cpu(), e.p.cpu(), e.offsets.cpu(), e.offsets_host, e.t_div)

    # ---- C1: single N-Caltech101-shaped sample, 2-ch count frame + 5-bin voxel grid, one fused call ----
    h, w, n = 180, 240, 200_000
    e1 = make_batch_gpu(rank, dev, batch=1, mean=n, size=(h, w), spread=0.0, seed=1001)
    o = {}
    ms = timed_ms(lambda: ep.bin_events(e1, (h, w), num_bins=5, count_channels=2, out=o), 50)
    par = None
    if rank == 0:
        s = sample_aos(host_of(e1), 0)
        par = bool(close(o["voxel"][0].cpu().numpy(), oe.voxel_grid(s, 5, (h, w))) and
                   np.array_equal(o["count"][0].cpu().numpy(), oe.count_frame(s, (h, w), 2)))
    entry("C1 single 240x180 sample, 200k events -> 2-ch count frame + 5-bin voxel grid (one call)", ms, 13 * n + 4 * 7 * h * w, n, "events", par,
          note="3.8 MB of work: latency-bound by construction (SURVEY 8d); the CPU port takes ~3.4 ms for the same sample on one core")
    del e1, o

    # ---- C3: masked-modelling input pipeline, ViT-S/16 @224, 75 % mask, B = 128 per GPU (1024 over 8) ----
    B, C, Lp, K, p = 128, 5, 196, 49, 16
    pipe = ep.MaskedInputPipeline(B, C, (224, 224), p, 0.75, dev)
    g = torch.Generator(device=dev).manual_seed(3000 + rank)
    pipe.x.copy_(torch.randn(pipe.x.shape, device=dev, generator=g))
    pipe.sub_frame.copy_(torch.randn(pipe.sub_frame.shape, device=dev, generator=g))
    ms = timed_ms(lambda: pipe.run(), 50)
    alg = B * (4 * Lp + 8 * K + 12 * Lp + 2 * 4 * K * C * p * p + 2 * 4 * 224 * 224)
    par = None
    if rank == 0:
        r = pipe.run(draw_noise=False)
        noise = pipe.noise[:4].cpu().numpy()
        rk, rm, rr = s3.mask_from_noise(noise, K)
        xs = pipe.x[:4].cpu().numpy()
        par = bool(np.array_equal(r["ids_keep"][:4].cpu().numpy(), rk) and np.array_equal(r["mask"][:4].cpu().numpy(), rm) and
                   np.array_equal(r["ids_restore"][:4].cpu().numpy(), rr) and
                   np.array_equal(r["visible_patches"][:4].cpu().numpy(), s3.patchify_gather(xs, p, rk, "cpq")) and
                   close(r["target"][:4].cpu().numpy(), s3.target_normpix(pipe.sub_frame[:4].cpu().numpy(), p), 1e-5, 1e-5))
    entry("C3 pretraining input pipeline: ViT-S/16 @224, C=5, 75 % random mask, B=128 per GPU: noise + mask + visible-patch gather + "
          "norm_pix target (one CUDA graph)", ms, alg, B, "samples", par, launches_per_step=3)
    del pipe

    # ---- C4: DSEC-shaped, 640x440 after the crop, 15 bins, ~2M events per sample, B = 32 per GPU; voxel grid + EvRep ----
    h, w, bins, Bc = 440, 640, 15, 32
    e4 = make_batch_gpu(rank, dev, batch=Bc, mean=2_000_000, size=(h, w), spread=0.1, seed=4000, window_us=100_000)
    h4 = host_of(e4)
    h4t = h4.transport()
    t4 = h4t.to(dev)
    o = {"voxel": torch.empty((Bc, bins, h, w), dtype=torch.float32, device=dev)}
    ms = timed_ms(lambda: ep.bin_events(t4, (h, w), num_bins=bins, out=o), 5)
    par = bool(close(o["voxel"][1].cpu().numpy(), oe.voxel_grid(sample_aos(h4, 1), bins, (h, w)))) if rank == 0 else None
    entry("C4 DSEC-shaped B=32 per GPU, 640x440, 15 bins, ~2M events/sample: voxel grid", ms, 13 * e4.num_events + 4 * bins * h * w * Bc,
          e4.num_events, "events", par, layout=layout_name_of(h4t)[0])
    del o
    keep = {"ev": torch.empty((Bc, 3, h, w), dtype=torch.float64, device=dev)}

    def do_evrep():
        ep.evrep(t4, (h, w), out=keep["ev"])       # the transport layout: routed shared-memory path
    ms = timed_ms(do_evrep, 5, warm=2)
    ms_canon = timed_ms(lambda: ep.evrep(e4, (h, w), out=keep["ev"]), 2, warm=1)
    do_evrep()
    par = None
    if rank == 0:
        s = sample_aos(h4, 1)
        # stamp value = ticks / t_div on both sides (the same fp64 quotient)
        par = bool(np.array_equal(keep["ev"][1].cpu().numpy(), oe.evrep(s[:, 0], s[:, 1], s[:, 2], s[:, 3], (w, h)), equal_nan=True))
    entry("C4 EvRep (3,440,640) f64, the pinned stand-in for the time surface (events_to_image.py:77-125)", ms,
          13 * e4.num_events + 12 * h * w * Bc, e4.num_events, "events", par, layout=layout_name_of(h4t)[0],
          canonical_layout_global_sort_ms=allmax(ms_canon),
          bytes_rule="13 B/event + 12 B per pixel (three planes at 4 B, as in round 1; the kernel writes the reference's float64: 24 B)")
    keep.clear()
    del t4, h4t
    ms = timed_ms(lambda: ep.time_surface(e4, (h, w), tau=0.03), 5)
    entry("C4 exponential time surface (2,440,640) [parity unpinned: no reference routine, SURVEY F5]", ms,
          13 * e4.num_events + 8 * h * w * Bc, e4.num_events, "events", None)
    del e4, h4
    torch.cuda.empty_cache()

    # ---- C5: MVSEC-shaped, 346x260, 9 bins, B = 512 per GPU: paired output (org grid + bilinear 224x224), block-mask variants ----
    h, w, bins, Bc = 260, 346, 9, 512
    e5 = make_batch_gpu(rank, dev, batch=Bc, mean=100_000, size=(h, w), spread=0.5, seed=5000, window_us=50_000)
    h5 = host_of(e5)
    h5t = h5.transport()
    t5 = h5t.to(dev)
    o = {"voxel": torch.empty((Bc, bins, h, w), dtype=torch.float32, device=dev)}
    from eventpretrain_b200.view_augment import ViewChoice
    full = ep.prepare_views([ViewChoice(0, 0, w, h, False, False, False)] * Bc, h, w, dev)      # the pair's resize is the whole frame: prepared once
    paired = {}

    def do_paired():
        ep.bin_events(t5, (h, w), num_bins=bins, out=o)
        paired["resized"] = ep.apply_views(o["voxel"], full, (224, 224), "bilinear")      # ft_mvsec_dataset.py:229-239
    ms = timed_ms(do_paired, 5)
    par = None
    if rank == 0:
        import torch.nn.functional as F
        ok_v = close(o["voxel"][3].cpu().numpy(), oe.voxel_grid(sample_aos(h5, 3), bins, (h, w)))
        ref_r = F.interpolate(o["voxel"][:8], size=(224, 224), mode="bilinear")
        par = bool(ok_v and torch.allclose(paired["resized"][:8], ref_r, rtol=1e-5, atol=1e-6))
    entry("C5 MVSEC-shaped B=512 per GPU, 346x260, 9 bins, ~100k events/sample: paired output = org voxel grid + bilinear 224x224 copy",
          ms, 13 * e5.num_events + 4 * bins * Bc * (h * w + 224 * 224), e5.num_events, "events", par,
          samples_per_s=allsum(Bc) / (allmax(ms) * 1e-3), layout=layout_name_of(h5t)[0])
    del o, paired, t5, e5, h5, h5t
    mask = (torch.rand(Bc, 196, device=dev) < 0.75).float()
    ms = timed_ms(lambda: ep.convvit_keep_masks(mask), 50)
    par = None
    if rank == 0:
        m1, m2 = ep.convvit_keep_masks(mask)
        r1, r2 = s3.block_mask_expand(mask.cpu().numpy(), 14, 4), s3.block_mask_expand(mask.cpu().numpy(), 14, 2)
        par = bool(np.array_equal(m1.cpu().numpy(), r1) and np.array_equal(m2.cpu().numpy(), r2))
    entry("C5 ConvViT block masks (512,196) -> (512,1,56,56), (512,1,28,28)", ms, Bc * 4 * (196 + 56 * 56 + 28 * 28), Bc, "samples", par)
    xs = torch.randn(Bc, 3136, 96, device=dev)
    m49 = torch.zeros(Bc, 49, device=dev)
    m49[:, torch.randperm(49, device=dev)[:37]] = 1
    ms = timed_ms(lambda: ep.swin_apply_mask(xs, m49, (56, 56), n_vis=12 * 64), 20)
    par = None
    if rank == 0:
        got = ep.swin_apply_mask(xs, m49, (56, 56), n_vis=12 * 64)
        rx, rc = s3.swin_apply_mask(xs[:2].cpu().numpy(), m49[:1].cpu().numpy() > 0, (56, 56))[:2]
        par = bool(np.array_equal(got[0][:2].cpu().numpy(), rx) and np.array_equal(got[1].cpu().numpy().reshape(rc.shape), rc))
    entry("C5 Swin apply_mask (512,3136,96) -> (512,768,96) + coords", ms, Bc * 2 * 4 * 768 * 96, Bc, "samples", par)
    return res


_RESULT_FD = None


def emit(line):
    """The one JSON line of the contract, on the process's real stdout."""
    data = (json.dumps(line) + "\n").encode()
    if _RESULT_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_RESULT_FD, data)


def main():
    # stdout carries exactly one JSON line: everything else a library prints there (NCCL's version banner, for one)
    # is sent to stderr by pointing fd 1 at fd 2 for the lifetime of the process; emit() writes to the saved descriptor
    global _RESULT_FD
    sys.stdout.flush()
    _RESULT_FD = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--method", default="auto", choices=["auto", "global", "tiled"], help="binning kernel family")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
