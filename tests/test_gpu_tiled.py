"""The default fast path of ep_bin_events for the 4 B/event transport layout (route + two-plane shared-memory sweep,
csrc/ep_binning_tiled.cu) against the global-RED kernels (bit for bit: same Q24 integers, one rounding) and against the
CPU oracle of events_to_voxel_grid (dataset/dataset_utils/events_to_voxel_grid.py:4-61) within 1e-5*|ref| + 1e-6."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


PLANE_CELLS = 54272     # csrc/ep_binning_tiled.cu: cells of one CTA's plane tile in the whole-plane kernels


def plane_tiles(H, W):
    """Row tiles the whole-plane kernels cut an H x W grid into (1 = the plane fits one SM)."""
    if W > PLANE_CELLS:
        return 1 << 20
    rows = min(PLANE_CELLS // W, H)
    T = -(-H // rows)
    rows = -(-H // T)
    return -(-H // rows)


def plane_takes(H, W, cells, n_events):
    """The rule of tiled_plan(): up to 2 row tiles, 3 when the output outweighs the events (cells >= 2 x events)."""
    t = plane_tiles(H, W)
    return t <= 2 or (t == 3 and cells >= 2 * n_events)


def close(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return bool(np.all(np.abs(a - b) <= 1e-5 * np.abs(b) + 1e-6))


@pytest.fixture(scope="module")
def ep(native_lib):
    import eventpretrain_b200 as ep
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    return ep


def dense_batch(ep, rng, counts, H, W, Ws=None, Hs=None, ticks_per_event=0.25, hot=None, block_shuffle=False, t0=0, edit=None):
    """Host batch whose stamps fit the 4 B layout: sorted integer ticks, ~4 events per tick.  block_shuffle permutes the
    256-event blocks of each sample (dense inside a block, out of order between blocks: first/last rows are then not
    min/max and chunks span several temporal intervals)."""
    Ws, Hs = Ws or W, Hs or H
    xs, ys, ts, ps = [], [], [], []
    for n in counts:
        x = rng.integers(0, Ws, n); y = rng.integers(0, Hs, n); p = rng.integers(0, 2, n)
        t = np.sort(rng.integers(0, max(int(n * ticks_per_event), 1), n)).astype(np.int64) + t0
        if hot is not None and n > 4 * hot:
            x[n // 4: n // 4 + hot] = 3
            y[n // 4: n // 4 + hot] = 2
            p[n // 4: n // 4 + hot] = 1
        xs.append(x); ys.append(y); ts.append(t); ps.append(p)
    off = np.cumsum([0] + list(counts))
    if block_shuffle:
        # blocks are cut at array positions (i // 256), so permute whole array-aligned blocks inside each sample
        for b, n in enumerate(counts):
            lo = int(off[b])
            first = (-lo) % 256 + 256          # the block where the sample starts counts from the sample's smallest stamp: kept
            nb = (n - first) // 256
            if nb > 1:
                perm = rng.permutation(nb)
                for arr in (xs, ys, ts, ps):
                    body = arr[b][first:first + nb * 256].reshape(nb, 256)[perm].reshape(-1)
                    arr[b] = np.concatenate([arr[b][:first], body, arr[b][first + nb * 256:]])
    if edit is not None:
        edit(xs, ys, ts, ps)
    ev = ep.from_soa(np.concatenate(xs).astype(np.uint16), np.concatenate(ys).astype(np.uint16), np.concatenate(ts),
                     np.concatenate(ps).astype(np.uint8), off, t_div=1e6, pin=False)
    samples = [np.stack([xs[b], ys[b], ts[b].astype(np.float64) / 1e6, ps[b]], 1).astype(np.float64) for b in range(len(counts))]
    return ev, samples


def both(ep, p4, size, **kw):
    a = ep.bin_events(p4, size, method="global", **kw)
    b = ep.bin_events(p4, size, method="tiled", **kw)
    c = ep.bin_events(p4, size, **kw)                 # default = whole-plane kernels where the plane fits an SM, else tiled
    for key in a:
        assert torch.equal(a[key], b[key]), key
        assert torch.equal(a[key], c[key]), key
    if not kw.get("stats") and plane_takes(size[0], size[1], p4.batch * kw.get("num_bins", 0) * size[0] * size[1], p4.num_events):
        d = ep.bin_events(p4, size, method="plane", **kw)
        for key in a:
            assert torch.equal(a[key], d[key]), key
    elif not kw.get("stats"):
        with pytest.raises(RuntimeError):
            ep.bin_events(p4, size, method="plane", **kw)
    return b


@pytest.mark.parametrize("H,W,bins", [(224, 224, 5), (480, 640, 5), (44, 64, 15), (65, 87, 9), (180, 240, 2), (33, 50, 1),
                                      (700, 36, 3), (260, 346, 9), (261, 347, 4), (300, 500, 12)])
def test_tiled_equals_global_and_oracle(ep, H, W, bins):
    """Ragged batch with an empty sample, a 1-event sample, samples that start and end inside tick blocks and route chunks,
    and a hot pixel of 2000 same-polarity events (int32 plane words wrap: exercises the spill list)."""
    from oracle import events as oe
    rng = np.random.default_rng(H * 1000 + bins)
    counts = [30000, 0, 1, 70001, 4099, 12288, 8192, 255, 8193]
    ev, samples = dense_batch(ep, rng, counts, H, W, hot=2000)
    p4 = ev.packed(4).to("cuda")
    out = both(ep, p4, (H, W), num_bins=bins, voxel_sum=True, check=True)
    for i, s in enumerate(samples):
        if len(s) == 0:
            assert not out["voxel"][i].any() and not out["voxel_sum"][i].any()
            continue
        ref = oe.voxel_grid(s, bins, (H, W))
        got = out["voxel"][i].cpu().numpy()
        if len(s) > 8000:
            # the hot cell: the reference's own sequential fp32 sum of 2000 terms drifts by ~1e-5 relative from the exact
            # value (measured: 1.2e-5 at 2 bins), the fixed-point sum does not: compared at 1e-4 there
            assert np.allclose(got[:, 2, 3], ref[:, 2, 3], rtol=1e-4, atol=1e-6), i
            got[:, 2, 3] = ref[:, 2, 3]
        assert close(got, ref), i
        gs = out["voxel_sum"][i, 0].cpu().numpy()
        rs = ref.sum(0)
        gs[2, 3] = rs[2, 3]
        assert close(gs, rs), i
    # hot cell: 2000 events of weight up to 2^24 each do not fit an int32 word
    assert float(out["voxel_sum"][3, 0, 2, 3]) > 1500
    # voxel only (no sum plane), shard with offsets[0] > 0
    v = ep.bin_events(p4, (H, W), num_bins=bins, method="tiled", check=True)
    assert torch.equal(v["voxel"], out["voxel"])
    sh = ep.bin_events(p4.shard(1, 2), (H, W), num_bins=bins, method="tiled", check=True)
    assert torch.equal(sh["voxel"], out["voxel"][4:])
    sh = ep.bin_events(p4.shard(1, 2), (H, W), num_bins=bins, voxel_sum=True, check=True)        # default path of this shape, offsets[0] > 0
    assert torch.equal(sh["voxel"], out["voxel"][4:]) and torch.equal(sh["voxel_sum"], out["voxel_sum"][4:])
    # run-to-run bit identity
    again = ep.bin_events(p4, (H, W), num_bins=bins, voxel_sum=True, method="tiled")
    assert torch.equal(again["voxel"], out["voxel"]) and torch.equal(again["voxel_sum"], out["voxel_sum"])


def test_tiled_reference_res_scale(ep):
    """The reference's own pre-training order: events_reshape 640x480 -> 224x224 (fp64 product, truncation), then the voxel
    grid (pr_n_imagenet_dataset.py:85-87); includes the x = 180, 340, 360 cases where 0.35 * x lands below the integer."""
    from oracle import events as oe
    rng = np.random.default_rng(5)
    counts = [50000, 9000]
    def trap(xs, ys, ts, ps):
        xs[0][:6] = [180, 340, 360, 639, 0, 20]
    ev, samples = dense_batch(ep, rng, counts, 224, 224, Ws=640, Hs=480, edit=trap)
    p4 = ev.packed(4).to("cuda")
    sc = (224 / 640, 224 / 480)
    out = both(ep, p4, (224, 224), num_bins=5, voxel_sum=True, scale=sc, check=True)
    for i, s in enumerate(samples):
        r = s.copy()
        r[:, 0] *= sc[0]
        r[:, 1] *= sc[1]
        assert close(out["voxel"][i].cpu().numpy(), oe.voxel_grid(r, 5, (224, 224))), i


def test_tiled_unsorted_blocks(ep):
    """Stamps out of order between tick blocks: first / last rows are not min / max (events before the first row or
    after the last are dropped, events_to_voxel_grid.py:44-45,51-52) and a chunk's events span several intervals."""
    from oracle import events as oe
    rng = np.random.default_rng(31)
    H, W, bins = 48, 64, 5
    ev, samples = dense_batch(ep, rng, [40000, 9000, 20000], H, W, block_shuffle=True)
    p4 = ev.packed(4).to("cuda")
    out = both(ep, p4, (H, W), num_bins=bins, voxel_sum=True, check=True)
    for i, s in enumerate(samples):
        assert close(out["voxel"][i].cpu().numpy(), oe.voxel_grid(s, bins, (H, W))), i


def test_tiled_degenerate_time(ep):
    """deltaT == 0 (all stamps equal: the reference divides by 1.0, events_to_voxel_grid.py:24-25) and last row earlier
    than the first: samples without integer-time constants take the fp64 expression inside the sweep."""
    from oracle import events as oe
    H, W, bins = 40, 56, 5
    rng = np.random.default_rng(8)
    n = 3000
    x = rng.integers(0, W, 3 * n).astype(np.uint16); y = rng.integers(0, H, 3 * n).astype(np.uint16)
    p = rng.integers(0, 2, 3 * n).astype(np.uint8)
    t = np.empty(3 * n, np.int64)
    t[:n] = 1000                                         # all equal
    t[n:2 * n] = np.sort(rng.integers(0, 400, n))[::-1] + 5000    # descending: last < first
    mid = np.sort(rng.integers(0, 400, n))
    mid[0], mid[-1] = 77, 77                             # first == last, others differ
    t[2 * n:] = mid
    off = np.array([0, n, 2 * n, 3 * n])
    ev = ep.from_soa(x, y, t, p, off, t_div=1e6, pin=False)
    p4 = ev.packed(4).to("cuda")
    out = both(ep, p4, (H, W), num_bins=bins, voxel_sum=True, check=True)
    for b in range(3):
        s = np.stack([x[off[b]:off[b + 1]], y[off[b]:off[b + 1]], t[off[b]:off[b + 1]] / 1e6, p[off[b]:off[b + 1]]], 1).astype(np.float64)
        assert close(out["voxel"][b].cpu().numpy(), oe.voxel_grid(s, bins, (H, W))), b


def test_tiled_row_wrap_and_bad_events(ep):
    """x >= W is not an error by itself in the reference: the flat index x + y * W lands in a later row
    (events_to_voxel_grid.py:46); only indices past the grid raise."""
    rng = np.random.default_rng(3)
    H, W = 48, 64
    from oracle import events as oe

    def wrap(xs, ys, ts, ps):
        xs[0][17], ys[0][17] = W + 5, 10          # row wrap: valid
        xs[0][99], ys[0][99] = 2 * W + 1, H - 3   # lands on the last row: valid
    ev, samples = dense_batch(ep, rng, [5000] * 12, H, W, edit=wrap)
    p4 = ev.packed(4).to("cuda")
    out = both(ep, p4, (H, W), num_bins=5, voxel_sum=True, check=True)
    assert close(out["voxel"][0].cpu().numpy(), oe.voxel_grid(samples[0], 5, (H, W)))

    def bad(xs, ys, ts, ps):
        xs[3][1234], ys[3][1234] = 64 + 29 * 64, 47          # flat index outside the grid
    ev, _ = dense_batch(ep, rng, [5000] * 12, H, W, edit=bad)
    p4 = ev.packed(4).to("cuda")
    for method in ("tiled", "global", None):
        with pytest.raises(IndexError):
            ep.bin_events(p4, (H, W), num_bins=5, method=method, check=True)


@pytest.mark.parametrize("method,shuffle", [("tiled", False), ("plane", False), ("plane", True)])
def test_tiled_in_cuda_graph(ep, method, shuffle):
    """No allocation, no synchronisation, no host-side state: the call can be captured and replayed (the whole-plane
    kernels too, including their stand-by fallback for an unsorted sample)."""
    rng = np.random.default_rng(12)
    H, W = 96, 128
    ev, _ = dense_batch(ep, rng, [20000, 30000], H, W, block_shuffle=shuffle)
    p4 = ev.packed(4).to("cuda")
    ref = ep.bin_events(p4, (H, W), num_bins=5, voxel_sum=True, method="global")
    out = {k: torch.empty_like(v) for k, v in ref.items()}
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        ep.bin_events(p4, (H, W), num_bins=5, voxel_sum=True, method=method, out=out)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            ep.bin_events(p4, (H, W), num_bins=5, voxel_sum=True, method=method, out=out)
        for v in out.values():
            v.zero_()
        g.replay()
    torch.cuda.synchronize()
    assert torch.equal(out["voxel"], ref["voxel"]) and torch.equal(out["voxel_sum"], ref["voxel_sum"])


@pytest.mark.parametrize("H,W,bins,n,B", [(440, 640, 15, 2_000_000, 2), (260, 346, 9, 100_000, 16)])
def test_config_sizes_vs_oracle(ep, H, W, bins, n, B):
    """BASELINE configs[3] (DSEC-shaped: 640x440 after the crop, 15 bins, ~2 M events per sample, microsecond stamps) and
    configs[4] (MVSEC-shaped: 346x260, 9 bins) at their own sizes against the oracle, every layout and both kernel
    families: the size-dependent code (32-bit index arithmetic, L2 groups, finalize chunking, 22 / 8 row tiles)."""
    from oracle import events as oe
    rng = np.random.default_rng(4000 + bins)
    counts = [int(n * u) for u in rng.uniform(0.9, 1.1, B)]
    ev, samples = dense_batch(ep, rng, counts, H, W)
    canon = ev.to("cuda")
    p4 = ev.packed(4).to("cuda")
    a = ep.bin_events(canon, (H, W), num_bins=bins, voxel_sum=True, check=True)
    out = both(ep, p4, (H, W), num_bins=bins, voxel_sum=True, check=True)
    assert torch.equal(a["voxel"], out["voxel"]) and torch.equal(a["voxel_sum"], out["voxel_sum"])
    for i in range(min(B, 4)):
        assert close(out["voxel"][i].cpu().numpy(), oe.voxel_grid(samples[i], bins, (H, W))), i


def test_fused_statistics(ep):
    """(count, sum, sum of squares, max) per channel: by-product of the tiled kernels, one native pass elsewhere; both equal
    the fp64 statistics of the tensors they describe and are bit-reproducible run to run."""
    from eventpretrain_b200 import dist as epd
    rng = np.random.default_rng(77)
    H, W, bins = 96, 128, 5
    ev, _ = dense_batch(ep, rng, [20000, 0, 30000, 511, 9000], H, W, hot=300)
    p4 = ev.packed(4).to("cuda")

    def table(x):
        xd = x.double().transpose(0, 1).reshape(x.shape[1], -1)
        return torch.stack([torch.full((x.shape[1],), float(xd.shape[1]), dtype=torch.float64, device=x.device), xd.sum(1),
                            (xd * xd).sum(1), xd.amax(1)], 1)

    for method in ("tiled", "global", "plane", None):
        o = ep.bin_events(p4, (H, W), num_bins=bins, voxel_sum=True, stats=True, method=method)
        ref = torch.cat([table(o["voxel"]), table(o["voxel_sum"])], 0)
        assert torch.allclose(o["stats"], ref, rtol=1e-6, atol=1e-6), method
        assert torch.equal(o["stats"][:, 0], ref[:, 0]) and torch.equal(o["stats"][:, 3], ref[:, 3]), method
        again = ep.bin_events(p4, (H, W), num_bins=bins, voxel_sum=True, stats=True, method=method)
        assert torch.equal(again["stats"], o["stats"]), method
    # whole-plane kernels: two row tiles without the sum plane, and an unsorted batch (the stand-by route + sweep redo the
    # planes and bring their own reduction: same table as the forced tiled path, bit for bit)
    ev2, _ = dense_batch(ep, rng, [30000, 500, 12000], 260, 346, hot=300)
    q4 = ev2.packed(4).to("cuda")
    o = ep.bin_events(q4, (260, 346), num_bins=9, stats=True, method="plane")
    assert torch.allclose(o["stats"][:9], table(o["voxel"]), rtol=1e-6, atol=1e-6), (o["stats"][:9] - table(o["voxel"])).abs().max(0)
    ev3, _ = dense_batch(ep, rng, [20000, 9000], H, W, block_shuffle=True)
    r4 = ev3.packed(4).to("cuda")
    a = ep.bin_events(r4, (H, W), num_bins=bins, voxel_sum=True, stats=True, method="plane")
    b = ep.bin_events(r4, (H, W), num_bins=bins, voxel_sum=True, stats=True, method="tiled")
    assert torch.equal(a["stats"], b["stats"]) and torch.equal(a["voxel"], b["voxel"])
    x = torch.randn(7, 3, 33, 50, device="cuda")
    assert torch.allclose(epd.plane_statistics(x), table(x), rtol=1e-12, atol=1e-9)
    fin = epd.finalize_statistics(epd.plane_statistics(x))
    assert torch.allclose(fin["mean"], x.double().mean((0, 2, 3)), atol=1e-12)


@pytest.mark.parametrize("bins", [1, 2, 3, 5, 9])
def test_plane_path_mixed_batch(ep, bins):
    """Whole-plane kernels (one output plane per CTA in shared memory, csrc/ep_binning_tiled.cu) on the reference's own
    pre-training shape — events_reshape 640x480 -> 224x224 fused (pr_n_imagenet_dataset.py:85-87) — with samples that end on
    repeated last stamps, many events on interval nodes, a sample of one tick block, and out-of-grid events counted once;
    then the same batch with one unsorted sample: the stand-by route + sweep kernels redo it, same bits, same count."""
    from oracle import events as oe
    rng = np.random.default_rng(900 + bins)
    sc = (224 / 640, 224 / 480)

    def nodes(xs, ys, ts, ps):
        ts[1][-50:] = ts[1][-1]                      # 50 events on the last stamp: interval bins - 1, weight 1 on the last plane
        span = ts[2][-1] - ts[2][0]
        for j in range(1, max(bins - 1, 1)):         # events exactly on the interior nodes
            k = np.searchsorted(ts[2], ts[2][0] + span * j // max(bins - 1, 1))
            ts[2][k:k + 3] = ts[2][k]
    counts = [60000, 30011, 45000, 200, 0, 8191]
    ev, samples = dense_batch(ep, rng, counts, 224, 224, Ws=640, Hs=480, edit=nodes, hot=700)
    p4 = ev.packed(4).to("cuda")
    out = both(ep, p4, (224, 224), num_bins=bins, voxel_sum=True, scale=sc, check=True)
    for i, s in enumerate(samples):
        if len(s) == 0:
            continue
        r = s.copy()
        r[:, 0] *= sc[0]
        r[:, 1] *= sc[1]
        ref = oe.voxel_grid(r, bins, (224, 224))
        got = out["voxel"][i].cpu().numpy()
        hx, hy = int(3 * sc[0]), int(2 * sc[1])
        got[:, hy, hx] = ref[:, hy, hx]               # the hot cell: the reference's sequential fp32 sum drifts (see above)
        assert close(got, ref), i

    def bad_and_unsorted(unsort):
        def edit(xs, ys, ts, ps):
            xs[0][100], ys[0][100] = 2047, 2047       # outside the 224 x 224 grid after the scale
            xs[2][7], ys[2][7] = 2047, 2047
            if unsort:
                for i, j in ((1900, 2200), (4000, 4200)):      # across interval boundaries (2048, 4096), inside the 9-bit tick range
                    ts[5][i], ts[5][j] = ts[5][j], ts[5][i]
        return edit
    counts = [30000, 256, 12000, 1, 0, 8191]
    for unsort in (False, True):
        ev, _ = dense_batch(ep, np.random.default_rng(77), counts, 224, 224, Ws=640, Hs=480, edit=bad_and_unsorted(unsort))
        p4 = ev.packed(4).to("cuda")
        res, bad = {}, {}
        for m in ("global", "tiled", "plane"):
            bad[m] = torch.zeros(1, dtype=torch.int32, device="cuda")
            res[m] = ep.bin_events(p4, (224, 224), num_bins=bins, voxel_sum=True, scale=sc, method=m, bad_out=bad[m])
        for m in ("tiled", "plane"):
            assert torch.equal(res[m]["voxel"], res["global"]["voxel"]), (m, unsort)
            assert torch.equal(res[m]["voxel_sum"], res["global"]["voxel_sum"]), (m, unsort)
            assert int(bad[m]) == int(bad["global"]) == 2, (m, unsort)


@pytest.mark.parametrize("H,W,bins", [(224, 224, 9), (260, 346, 9), (300, 500, 3), (50, 1000, 2)])
def test_plane_path_few_events(ep, H, W, bins):
    """Batches whose output outweighs the events (short samples, many bins: the flush and the per-(sample, tile) voxel.sum(0)
    dominate), on one, two and three row tiles: same bits as the other paths."""
    rng = np.random.default_rng(H + bins)
    counts = [3000, 0, 1, 5000, 257, 4096]
    ev, _ = dense_batch(ep, rng, counts, H, W, hot=300)
    p4 = ev.packed(4).to("cuda")
    assert plane_takes(H, W, p4.batch * bins * H * W, p4.num_events)
    both(ep, p4, (H, W), num_bins=bins, voxel_sum=True, check=True)
    both(ep, p4, (H, W), num_bins=bins, check=True)
