"""Stage 1 on the GPU (through the C ABI) against the golden vectors of the reference and the CPU oracle.

Tolerances: count frames, hot-pixel filter, normalisers and EvRep are bit-exact; voxel grids (float,
fixed-point accumulation) satisfy |a-b| <= 1e-5*|b| + 1e-6 (north_star's 1e-5 relative, with the
absolute floor SURVEY.md §7 shows any implementation needs)."""
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from conftest import reshaped
from test_oracle_golden import STAGE1

pytestmark = pytest.mark.gpu


def close(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return bool(np.all(np.abs(a - b) <= 1e-5 * np.abs(b) + 1e-6))


@pytest.fixture(scope="module")
def ep(native_lib):
    import eventpretrain_b200 as ep
    assert torch.cuda.is_available()
    return ep


@pytest.mark.parametrize("name", STAGE1)
def test_dropins_vs_golden(ep, golden_stage1, name):
    c = golden_stage1[name]
    ev = reshaped(c)
    size, bins = tuple(int(v) for v in c["size"]), int(c["bins"])
    args = SimpleNamespace(num_bins=bins)
    v = ep.events_to_voxel_grid(args, ev.copy(), size)
    assert v.dtype == torch.float32 and tuple(v.shape) == (bins,) + size and not v.is_cuda
    assert close(v.numpy(), c["voxel"]), np.abs(v.numpy() - c["voxel"]).max()
    e = ep.events_to_image_ecdp(args, ev.copy(), size)
    m = ep.events_to_image_mem(args, ev.copy(), size)
    assert np.array_equal(e.numpy(), c["ecdp"])
    assert np.array_equal(m.numpy(), c["mem"])
    hot = ep.remove_hot_pixel_mem(m / 255)
    assert np.array_equal(hot.numpy(), c["mem_hot"])
    en = ep.normalise(e.cuda(), "count").cpu().numpy()
    assert np.array_equal(en, c["ecdp_norm"])
    if "mem_norm" in c:
        assert np.array_equal(ep.normalise(hot.cuda(), "mem").cpu().numpy(), c["mem_norm"])
    if "evrep" in c:
        xs, ys = ev[:, 0].astype(np.int16), ev[:, 1].astype(np.int16)
        t = ev[:, 2].astype(np.float64)
        r = ep.events_to_EvRep(xs, ys, t, ev[:, 3], (size[1], size[0]))
        assert r.dtype == np.float64
        assert np.array_equal(r, c["evrep"], equal_nan=True)
        assert np.array_equal(ep.events_to_EvRep(xs, ys, t * 1e6, ev[:, 3], (size[1], size[0])), c["evrep_us"],
                              equal_nan=True)


def test_mem_guard_normaliser(ep):
    """EP_NORM_MEM_GUARD against the statements of ft_mvsec_dataset.py:244-249 (3-channel MEM image; channels 0 and 2 are scaled
    by 1 / max, by 1 / 0.001 when that max is 0), executed in torch on the same tensors: bit-exact, batch of mixed cases."""
    g = torch.Generator().manual_seed(77)
    frames = torch.randint(0, 40, (5, 3, 33, 47), generator=g).float()
    frames[1] = 0                              # an all-zero frame: the guard
    frames[2, 0::2] = 0                        # only the untouched middle channel is populated
    frames[3, 0::2] *= -1                      # max of the scaled channels is 0 with negative entries elsewhere
    frames[3, 0, 0, 0] = 0
    frames[4] /= 7                             # non-integer values: the division result is not exact
    want = frames.clone()
    for b in range(frames.shape[0]):
        v = want[b]
        if v[0::2, :, :].max() != 0:
            factor = 1.0 / v[0::2, :, :].max()
        else:
            factor = 1.0 / 0.001
        v[0::2, :, :] = v[0::2, :, :] * factor
    got = ep.normalise(frames.cuda(), "mem_guard").cpu()
    assert torch.equal(got, want)
    one = ep.normalise(frames[4].clone().cuda(), "mem_guard").cpu()          # the (C,H,W) form
    assert torch.equal(one, want[4])


def test_fused_reshape_scale(ep, golden_stage1):
    """events_reshape fused as scale=(sx, sy): x*sx in fp64 then truncation (the 640->224 trap)."""
    for name in ("reshape_trap", "reshape_mvsec"):
        c = golden_stage1[name]
        sw, sh, iw, ih = (int(v) for v in c["reshape"])
        size, bins = tuple(int(v) for v in c["size"]), int(c["bins"])
        raw = torch.from_numpy(c["events"]).cuda()
        out = ep.bin_events_aos(raw, size, num_bins=bins, count_channels=2, scale=ep.reshape_scale(sw, sh, iw, ih))
        assert np.array_equal(out["count"].cpu().numpy(), c["ecdp"])
        assert close(out["voxel"].cpu().numpy(), c["voxel"])
        # canonical SoA batch of the same sample twice
        batch = ep.pack_events([c["events"], c["events"]]).to("cuda")
        o = ep.bin_events(batch, size, num_bins=bins, count_channels=2, scale=ep.reshape_scale(sw, sh, iw, ih),
                          voxel_sum=True, check=True)
        for b in range(2):
            assert np.array_equal(o["count"][b].cpu().numpy(), c["ecdp"])
            assert close(o["voxel"][b].cpu().numpy(), c["voxel"])
            assert close(o["voxel_sum"][b].cpu().numpy(), c["voxel_sum"])


def test_ragged_batch_vs_oracle(ep):
    """Canonical SoA (u16,u16,i64 us,u8) ragged batch, incl. an empty sample, vs the oracle per sample."""
    from oracle import events as oe
    rng = np.random.default_rng(77)
    H, W, bins = 60, 80, 5
    counts = [5000, 0, 1, 12345, 2, 777]
    xs, ys, ts, ps, samples = [], [], [], [], []
    for n in counts:
        x = rng.integers(0, W, n); y = rng.integers(0, H, n); p = rng.integers(0, 2, n)
        t_us = np.sort(rng.integers(0, 300000, n)).astype(np.int64)
        xs.append(x); ys.append(y); ts.append(t_us); ps.append(p)
        samples.append(np.stack([x, y, t_us.astype(np.float64) / 1e6, p], 1).astype(np.float64))
    off = np.cumsum([0] + counts)
    ev = ep.from_soa(np.concatenate(xs).astype(np.uint16), np.concatenate(ys).astype(np.uint16),
                     np.concatenate(ts), np.concatenate(ps).astype(np.uint8), off, t_div=1e6).to("cuda")
    out = ep.bin_events(ev, (H, W), num_bins=bins, count_channels=3, voxel_sum=True, check=True)
    for b, s in enumerate(samples):
        if len(s) == 0:
            assert not out["voxel"][b].any() and not out["count"][b].any()
            continue
        assert close(out["voxel"][b].cpu().numpy(), oe.voxel_grid(s, bins, (H, W))), b
        assert np.array_equal(out["count"][b].cpu().numpy(), oe.count_frame(s, (H, W), 3)), b
    # a shard of the same batch (offsets[0] > 0) gives the same rows
    sh = ev.shard(1, 2)
    o2 = ep.bin_events(sh, (H, W), num_bins=bins, check=True)
    assert torch.equal(o2["voxel"], out["voxel"][3:])


def test_group_chain_orderings(ep, monkeypatch):
    """The group loop is a chain of dependent launches (scatter -> finalize voxel -> finalize count -> next scatter) over
    shared accumulator slots.  One sample per group (EP_L2_GROUP_MB=1), empty samples in between (their groups launch no
    scatter), voxel + count + sum: every grouping must give the tensors of the single-group call bit for bit."""
    rng = np.random.default_rng(5)
    H, W, bins = 96, 128, 7
    counts = [3000, 0, 0, 9000, 1, 0, 20000, 512, 0]
    xs, ys, ts, ps = [], [], [], []
    for n in counts:
        xs.append(rng.integers(0, W, n)); ys.append(rng.integers(0, H, n)); ps.append(rng.integers(0, 2, n))
        ts.append(np.sort(rng.integers(0, 50000, n)).astype(np.int64))
    off = np.cumsum([0] + counts)
    ev = ep.from_soa(np.concatenate(xs).astype(np.uint16), np.concatenate(ys).astype(np.uint16), np.concatenate(ts),
                     np.concatenate(ps).astype(np.uint8), off, t_div=1e6).to("cuda")
    kw = dict(num_bins=bins, count_channels=2, voxel_sum=True, check=True)
    ref = {k: v.clone() for k, v in ep.bin_events(ev, (H, W), **kw).items()}
    monkeypatch.setenv("EP_L2_GROUP_MB", "1")
    for rep in range(3):                                  # repeated: the slots must come back zeroed every time
        got = ep.bin_events(ev, (H, W), **kw)
        for k in ref:
            assert torch.equal(got[k], ref[k]), (k, rep)
    got = ep.bin_events(ev, (H, W), num_bins=0, count_channels=3, check=True)        # count-only chain
    assert torch.equal(got["count"][:, 0], ref["count"][:, 0]) and torch.equal(got["count"][:, 2], ref["count"][:, 1])


def test_bin_events_in_cuda_graph(ep):
    """The whole binning call (memset + dependent-launch chain) can be captured into a CUDA graph and replayed on new data."""
    rng = np.random.default_rng(9)
    H, W, bins, n = 64, 96, 5, 40000
    def make(seed):
        r = np.random.default_rng(seed)
        off = np.array([0, n // 4, n // 4, n])
        t = np.concatenate([np.sort(r.integers(0, 90000, c)) for c in np.diff(off)]).astype(np.int64)
        return ep.from_soa(r.integers(0, W, n).astype(np.uint16), r.integers(0, H, n).astype(np.uint16), t,
                           r.integers(0, 2, n).astype(np.uint8), off, t_div=1e6)
    static = make(1).to("cuda")
    out = ep.bin_events(static, (H, W), num_bins=bins, count_channels=2, voxel_sum=True)         # warm-up: workspace, outputs
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        ep.bin_events(static, (H, W), num_bins=bins, count_channels=2, voxel_sum=True, out=out)
    torch.cuda.current_stream().wait_stream(s)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        ep.bin_events(static, (H, W), num_bins=bins, count_channels=2, voxel_sum=True, out=out)
    for seed in (2, 3):
        fresh = make(seed).to("cuda")
        for name in ("x", "y", "t", "p"):
            getattr(static, name).copy_(getattr(fresh, name))
        g.replay()
        ref = ep.bin_events(fresh, (H, W), num_bins=bins, count_channels=2, voxel_sum=True)
        for k in ref:
            assert torch.equal(out[k], ref[k]), (k, seed)


def test_generic_layouts_vs_oracle(ep, golden_stage1):
    """Non-canonical SoA dtype tags (float coords, fp32 time, int8 polarity) take the scalar-load kernel."""
    from oracle import events as oe
    c = golden_stage1["fractional_xy"]
    ev = c["events"]
    off = np.array([0, len(ev)])
    r = ep.from_soa(ev[:, 0], ev[:, 1], ev[:, 2], ev[:, 3], off).to("cuda")
    o = ep.bin_events(r, (48, 64), num_bins=5, count_channels=2, check=True)
    assert close(o["voxel"][0].cpu().numpy(), c["voxel"]) and np.array_equal(o["count"][0].cpu().numpy(), c["ecdp"])
    c = golden_stage1["f32_events"]
    ev = c["events"]
    r = ep.from_soa(ev[:, 0].astype(np.int32), ev[:, 1].astype(np.int32), ev[:, 2].astype(np.float32),
                    ev[:, 3].astype(np.uint8), np.array([0, len(ev)])).to("cuda")
    o = ep.bin_events(r, (48, 64), num_bins=5, time_f32=True, check=True)
    assert close(o["voxel"][0].cpu().numpy(), c["voxel"])
    c = golden_stage1["pm1_polarity"]
    ev = c["events"]
    r = ep.from_soa(ev[:, 0].astype(np.int16), ev[:, 1].astype(np.int16), ev[:, 2], ev[:, 3].astype(np.int8),
                    np.array([0, len(ev)])).to("cuda")
    o = ep.bin_events(r, (24, 32), num_bins=5, count_channels=2, check=True)
    assert close(o["voxel"][0].cpu().numpy(), c["voxel"]) and np.array_equal(o["count"][0].cpu().numpy(), c["ecdp"])


def test_error_behaviour(ep):
    args = SimpleNamespace(num_bins=5)
    with pytest.raises(IndexError):      # events[0, 2] on an empty array
        ep.events_to_voxel_grid(args, np.zeros((0, 4)), (4, 4))
    with pytest.raises(AssertionError):  # shape[1] == 4
        ep.events_to_voxel_grid(args, np.zeros((3, 3)), (4, 4))
    bad = np.array([[0, 0, 0.0, 1], [3, 99, 1.0, 1]], np.float64)
    with pytest.raises(IndexError):
        ep.events_to_voxel_grid(args, bad, (4, 4))
    with pytest.raises(IndexError):
        ep.events_to_image_ecdp(args, bad, (4, 4))
    with pytest.raises(IndexError):
        ep.events_to_EvRep(np.array([9], np.int16), np.array([0], np.int16), np.array([0.0]), np.array([1.0]), (4, 4))


def test_is_txyp(ep, golden_stage1):
    c = golden_stage1["c1_small"]
    ev = c["events"]
    txyp = ev[:, [2, 0, 1, 3]].copy()
    v = ep.events_to_voxel_grid(SimpleNamespace(num_bins=5), txyp, (180, 240), is_txyp=True)
    assert close(v.numpy(), c["voxel"])


def test_deterministic_and_order_independent(ep):
    """Fixed-point accumulation: bit-identical run to run and under any permutation of the events that
    keeps the first and last rows (which define the time window) in place."""
    rng = np.random.default_rng(5)
    n, H, W = 200000, 180, 240
    ev = np.stack([rng.integers(0, W, n), rng.integers(0, H, n), np.sort(rng.uniform(0, 0.3, n)),
                   rng.integers(0, 2, n)], 1).astype(np.float64)
    ev[n // 2: n // 2 + 3000, :2] = (7, 9)                      # a hot pixel
    d = torch.from_numpy(ev).cuda()
    a = ep.bin_events_aos(d, (H, W), num_bins=5, count_channels=2)
    b = ep.bin_events_aos(d, (H, W), num_bins=5, count_channels=2)
    assert torch.equal(a["voxel"], b["voxel"]) and torch.equal(a["count"], b["count"])
    perm = np.concatenate([[0], 1 + rng.permutation(n - 2), [n - 1]])
    c = ep.bin_events_aos(torch.from_numpy(ev[perm]).cuda(), (H, W), num_bins=5, count_channels=2)
    assert torch.equal(a["voxel"], c["voxel"]) and torch.equal(a["count"], c["count"])
    # conservation: every in-window event contributes p*(1-d) + p*d = p in total
    pol = np.where(ev[:, 3] == 0, -1.0, 1.0)
    assert abs(float(a["voxel"].double().sum()) - pol.sum()) < 1e-2


def test_evrep_batch(ep):
    from oracle import events as oe
    rng = np.random.default_rng(9)
    H, W = 44, 64
    counts = [3000, 1, 9000]
    parts = []
    for n in counts:
        parts.append((rng.integers(0, W, n).astype(np.int16), rng.integers(0, H, n).astype(np.int16),
                      np.sort(rng.uniform(0, 5e4, n)), rng.integers(0, 2, n).astype(np.float64)))
    parts[2][0][100:1500] = 3
    parts[2][1][100:1500] = 4                                    # hot pixel: long segment -> heap sort path
    off = np.cumsum([0] + counts)
    ev = ep.from_soa(*(np.concatenate([p[i] for p in parts]) for i in range(4)), off).to("cuda")
    out = ep.evrep(ev, (H, W), check=True).cpu().numpy()
    for b, p in enumerate(parts):
        assert np.array_equal(out[b], oe.evrep(p[0], p[1], p[2], p[3], (W, H))), b


def test_evrep_multi_chunk_scan(ep):
    """Grids larger than one 4096-pixel scan chunk, an empty sample in the middle, unsorted stamps, a hot pixel: bit-exact."""
    from oracle import events as oe
    rng = np.random.default_rng(19)
    H, W = 100, 131                                              # 13100 pixels: 4 chunks, the last one partial
    counts = [20000, 0, 7, 40000]
    parts = []
    for n in counts:
        parts.append((rng.integers(0, W, n).astype(np.int16), rng.integers(0, H, n).astype(np.int16),
                      rng.permutation(np.sort(rng.uniform(0, 9e4, n))) if n == 7 else np.sort(rng.uniform(0, 9e4, n)),
                      rng.integers(0, 2, n).astype(np.float64)))
    parts[3][0][500:900] = 130
    parts[3][1][500:900] = 99                                    # hot pixel in the last (partial) chunk
    off = np.cumsum([0] + counts)
    ev = ep.from_soa(*(np.concatenate([p[i] for p in parts]) for i in range(4)), off).to("cuda")
    for rep in range(2):                                         # the workspace is reused: second call must not see stale words
        out = ep.evrep(ev, (H, W), check=True).cpu().numpy()
        for b, p in enumerate(parts):
            if len(p[0]) == 0:
                assert not out[b][:2].any()
                continue
            assert np.array_equal(out[b], oe.evrep(p[0], p[1], p[2], p[3], (W, H))), (b, rep)


def _random_batch(ep, rng, counts, H, W, t_span=50_000, hot=None, unsorted=False):
    xs, ys, ts, ps = [], [], [], []
    for n in counts:
        x = rng.integers(0, W, n); y = rng.integers(0, H, n); p = rng.integers(0, 2, n)
        t = rng.integers(0, t_span, n).astype(np.int64)
        if not unsorted:
            t = np.sort(t)
        if hot is not None and n > 4 * hot:
            x[n // 4: n // 4 + hot] = 3
            y[n // 4: n // 4 + hot] = 2
            p[n // 4: n // 4 + hot] = 1
        xs.append(x); ys.append(y); ts.append(t); ps.append(p)
    off = np.cumsum([0] + list(counts))
    ev = ep.from_soa(np.concatenate(xs).astype(np.uint16), np.concatenate(ys).astype(np.uint16), np.concatenate(ts),
                     np.concatenate(ps).astype(np.uint8), off, t_div=1e6)
    samples = [np.stack([xs[b], ys[b], ts[b].astype(np.float64) / 1e6, ps[b]], 1).astype(np.float64) for b in range(len(counts))]
    return ev.to("cuda"), samples


def test_time_surface_self_oracle(ep):
    """No reference routine exists: checked against the numpy self-oracle."""
    from oracle import stage3_np as s3
    rng = np.random.default_rng(12)
    H, W = 30, 40
    ev, samples = _random_batch(ep, rng, [4000, 1, 900], H, W)
    out = ep.time_surface(ev, (H, W), tau=0.01, check=True).cpu().numpy()
    for b, s in enumerate(samples):
        ref = s3.time_surface(s[:, 0], s[:, 1], s[:, 2], s[:, 3], (H, W), 0.01)
        assert np.allclose(out[b], ref, rtol=1e-6, atol=1e-7), b


def test_compact_transport_layout(ep):
    """8 B/event transport layout (u16 x,y + u32 relative ticks | polarity << 31) gives bit-identical tensors."""
    rng = np.random.default_rng(41)
    H, W = 60, 80
    ev, _ = _random_batch(ep, rng, [30000, 0, 1, 5003], H, W)
    host = ep.RaggedEvents(ev.x.cpu(), ev.y.cpu(), ev.t.cpu() + 1_700_000_000_000_000, ev.p.cpu(), ev.offsets.cpu(),
                           ev.offsets_host, ev.t_div)
    comp = host.compact()
    assert comp.p is None and comp.t.dtype == torch.uint32 and comp.nbytes() < 0.65 * host.nbytes()
    a = ep.bin_events(host.to("cuda"), (H, W), num_bins=5, count_channels=2, voxel_sum=True, check=True)
    b = ep.bin_events(comp.to("cuda"), (H, W), num_bins=5, count_channels=2, voxel_sum=True, check=True)
    for key in ("voxel", "voxel_sum", "count"):
        assert torch.equal(a[key], b[key]), key
    c = ep.bin_events(comp.to("cuda").shard(1, 2), (H, W), num_bins=5, check=True)
    assert torch.equal(c["voxel"], a["voxel"][2:])


def test_packed_transport_layout(ep):
    """5 B/event transport layout (uint32 x | y << 11 | p << 22 | ticks >> 8 << 23 + one tick byte, block-relative
    ticks) gives bit-identical tensors; fused scale and count frames included."""
    rng = np.random.default_rng(43)
    H, W = 60, 80
    ev, _ = _random_batch(ep, rng, [30000, 0, 1, 5003, 1024, 2047, 4097], H, W)
    host = ep.RaggedEvents(ev.x.cpu(), ev.y.cpu(), ev.t.cpu() + 1_700_000_000_000_000, ev.p.cpu(), ev.offsets.cpu(),
                           ev.offsets_host, ev.t_div)
    pk = host.packed()
    assert pk.y is None and pk.x.dtype == torch.uint32 and pk.nbytes() < 0.42 * host.nbytes()
    a = ep.bin_events(host.to("cuda"), (H, W), num_bins=5, count_channels=2, voxel_sum=True, check=True)
    b = ep.bin_events(pk.to("cuda"), (H, W), num_bins=5, count_channels=2, voxel_sum=True, check=True)
    for key in ("voxel", "voxel_sum", "count"):
        assert torch.equal(a[key], b[key]), key
    sc = (0.5, 0.75)
    a = ep.bin_events(host.to("cuda"), (H, W), num_bins=9, scale=sc, check=True)
    b = ep.bin_events(pk.to("cuda"), (H, W), num_bins=9, scale=sc, check=True)
    assert torch.equal(a["voxel"], b["voxel"])
    c = ep.bin_events(pk.to("cuda").shard(1, 2), (H, W), num_bins=5, check=True)     # offsets[0] > 0 on the device side
    a5 = ep.bin_events(host.to("cuda"), (H, W), num_bins=5, check=True)
    assert torch.equal(c["voxel"], a5["voxel"][3:])
    # 4 B/event form (9-bit ticks per 256-event block) on a dense stream
    ev, _ = _random_batch(ep, rng, [30000, 0, 1, 5003, 256, 255, 257], H, W, t_span=300)
    host = ep.RaggedEvents(ev.x.cpu(), ev.y.cpu(), ev.t.cpu() + 1_700_000_000_000_000, ev.p.cpu(), ev.offsets.cpu(),
                           ev.offsets_host, ev.t_div)
    p4 = host.packed(4)
    assert p4.t is None and p4.nbytes() < 0.33 * host.nbytes()
    a = ep.bin_events(host.to("cuda"), (H, W), num_bins=5, count_channels=3, voxel_sum=True, check=True)
    b = ep.bin_events(p4.to("cuda"), (H, W), num_bins=5, count_channels=3, voxel_sum=True, check=True)
    for key in ("voxel", "voxel_sum", "count"):
        assert torch.equal(a[key], b[key]), key
    assert torch.equal(ep.evrep(p4.to("cuda"), (H, W), check=True), ep.evrep(host.to("cuda"), (H, W), check=True))   # routed EvRep
    with pytest.raises(RuntimeError):
        ep.evrep(pk.to("cuda"), (H, W))                 # the 5 B and 8 B transport layouts are for ep_bin_events only


def test_full_size_properties(ep):
    """BASELINE configs[1] sizes (640x480, ~1M events/sample, 5 bins; 24 samples here to keep the suite short): size-independent
    properties instead of an oracle pass — conservation (every in-window event adds p in total, so the grid sums to the
    sample's net polarity and voxel.sum(0) to the same), run-to-run bit identity, identical bits across the resident layouts
    and both kernel families (the 4 B transport layout takes the tiled shared-memory path by default), shards equal to the rows of the whole batch."""
    import bench
    dev = torch.device("cuda", 0)
    ev = bench.make_batch_gpu(0, dev, batch=24)
    H, W, bins = bench.H, bench.W, bench.BINS
    host = ep.RaggedEvents(ev.x.cpu(), ev.y.cpu(), ev.t.cpu(), ev.p.cpu(), ev.offsets.cpu(), ev.offsets_host, ev.t_div)
    a = ep.bin_events(ev, (H, W), num_bins=bins, voxel_sum=True, count_channels=2, check=True)
    # conservation, per sample (events exactly at t_first .. t_last all lie inside the bins)
    off = ev.offsets_host
    pol = ev.p.to(torch.float64) * 2 - 1
    for b in range(ev.batch):
        net = float(pol[int(off[b]):int(off[b + 1])].sum())
        assert abs(float(a["voxel"][b].double().sum()) - net) < 0.5
        assert abs(float(a["voxel_sum"][b].double().sum()) - net) < 0.5
        assert float(a["count"][b].sum()) == float(off[b + 1] - off[b])
    b2 = ep.bin_events(ev, (H, W), num_bins=bins, voxel_sum=True, count_channels=2)
    for key in ("voxel", "voxel_sum", "count"):
        assert torch.equal(a[key], b2[key]), key
    for other, method in ((host.transport().to(dev), None), (host.transport().to(dev), "global"), (host.packed(5).to(dev), None),
                          (host.compact().to(dev), None)):
        o = ep.bin_events(other, (H, W), num_bins=bins, voxel_sum=True, count_channels=2, check=True, method=method)
        for key in ("voxel", "voxel_sum", "count"):
            assert torch.equal(a[key], o[key]), (key, method)
    sh = ep.bin_events(ev.shard(1, 3), (H, W), num_bins=bins, check=True)
    assert torch.equal(sh["voxel"], a["voxel"][8:16])
