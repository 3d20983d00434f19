"""Host-side transport layouts (no GPU): RaggedEvents.packed() / compact() encode losslessly and refuse what they cannot hold."""
import numpy as np
import pytest
import torch

import eventpretrain_b200 as ep


def _batch(rng, counts, W=640, H=480, span=50_000, t0=1_700_000_000_000_000):
    xs, ys, ts, ps = [], [], [], []
    for n in counts:
        xs.append(rng.integers(0, W, n)); ys.append(rng.integers(0, H, n)); ps.append(rng.integers(0, 2, n))
        ts.append(np.sort(rng.integers(0, span, n)).astype(np.int64) + t0)
    off = np.cumsum([0] + list(counts))
    return ep.from_soa(np.concatenate(xs).astype(np.uint16), np.concatenate(ys).astype(np.uint16), np.concatenate(ts),
                       np.concatenate(ps).astype(np.uint8), off, t_div=1e6, pin=False)


def test_packed_round_trip():
    rng = np.random.default_rng(0)
    ev = _batch(rng, [3000, 0, 1, 5003, 1024, 2047, 9999, 0, 2])
    pk = ev.packed()
    assert pk.y is None and pk.x.dtype == torch.uint32 and pk.t.dtype == torch.uint8 and pk.p.dtype == torch.uint32
    assert pk.p.numel() == (ev.num_events + 1023) // 1024 and pk.t_base.numel() == ev.batch
    assert pk.nbytes() < 0.4 * ev.nbytes()
    x, y, t, p = pk.unpack_host()
    assert np.array_equal(x, ev.x.numpy()) and np.array_equal(y, ev.y.numpy())
    assert np.array_equal(t, ev.t.numpy()) and np.array_equal(p, ev.p.numpy())
    assert pk.packed() is pk
    with pytest.raises(ValueError):
        ev.packed(4)                                  # ~17 ticks per event on average: 256 events do not fit 9 bits
    assert ev.transport().t is not None and ev.transport().y is None      # ... so the 5 B/event form is the densest that fits


def test_packed4_round_trip():
    rng = np.random.default_rng(2)
    ev = _batch(rng, [3000, 0, 1, 5003, 256, 255, 257, 9999], span=400)    # dense stream: 256 events within 2^9 ticks
    pk = ev.packed(4)
    assert pk.y is None and pk.t is None and pk.x.dtype == torch.uint32 and pk.pack_block() == 256
    assert pk.p.numel() == (ev.num_events + 255) // 256 and pk.nbytes() < 0.32 * ev.nbytes()
    x, y, t, p = pk.unpack_host()
    assert np.array_equal(x, ev.x.numpy()) and np.array_equal(y, ev.y.numpy())
    assert np.array_equal(t, ev.t.numpy()) and np.array_equal(p, ev.p.numpy())
    assert ev.transport().t is None


def test_packed_mildly_unsorted_and_limits():
    rng = np.random.default_rng(1)
    ev = _batch(rng, [5000, 3000])
    t = ev.t.numpy().copy()
    t[100:200] = t[100:200][::-1]                    # local disorder inside a block is fine (offsets count from the block minimum)
    ev2 = ep.from_soa(ev.x.numpy(), ev.y.numpy(), t, ev.p.numpy(), ev.offsets_host, t_div=1e6, pin=False)
    x, y, tt, p = ev2.packed().unpack_host()
    assert np.array_equal(tt, t)
    sparse = _batch(rng, [3000], span=10_000_000)    # 3000 events over 10 s: a 1024-block spans far more than 2^17 ticks
    with pytest.raises(ValueError):
        sparse.packed()
    assert sparse.compact().t.dtype == torch.uint32   # the 8 B/event layout takes it
    wide = ep.from_soa(np.array([2048], np.uint16), np.array([0], np.uint16), np.array([5], np.int64), np.array([1], np.uint8),
                       np.array([0, 1]), pin=False)
    with pytest.raises(ValueError):
        wide.packed()
    with pytest.raises(ValueError):
        ev.shard(1, 2).packed()                       # block offsets are tied to array positions: whole batches only


def _same(a, b):
    for name in ("x", "y", "t", "p", "t_base"):
        u, v = getattr(a, name), getattr(b, name)
        assert (u is None) == (v is None), name
        assert u is None or (u.dtype == v.dtype and torch.equal(u, v)), name


@pytest.mark.parametrize("threads", [1, 3, 0])
def test_native_packer_equals_the_numpy_rule(threads):
    """ep_pack_transport_host (threaded C++) against the numpy statement of the layout rule, word for word: ragged batches with
    empty samples, samples starting exactly on block boundaries, mild disorder, both widths."""
    rng = np.random.default_rng(11)
    for counts, span in (([3000, 0, 1, 5003, 1024, 2047, 9999, 0, 2], 50_000), ([1024, 1024, 0, 512, 512, 70000], 30_000),
                         ([256, 0, 0, 256, 1, 255, 300000], 300), ([5], 3)):
        ev = _batch(rng, counts, span=span)
        t = ev.t.numpy().copy()
        if len(t) > 300:
            t[100:200] = t[100:200][::-1]
        ev = ep.from_soa(ev.x.numpy(), ev.y.numpy(), t, ev.p.numpy(), ev.offsets_host, t_div=1e6, pin=False)
        for nbytes in (4, 5):
            try:
                ref = ev.packed(nbytes, native=False)
            except ValueError:
                with pytest.raises(ValueError):
                    ev.packed(nbytes, native=True, threads=threads)
                continue
            _same(ev.packed(nbytes, native=True, threads=threads), ref)


def test_native_packer_refuses_what_the_layout_cannot_hold():
    one = lambda x, p, t: ep.from_soa(np.array([x], np.uint16), np.array([0], np.uint16), np.array(t, np.int64).reshape(-1)[:1],
                                      np.array([p], np.uint8), np.array([0, 1]), pin=False)
    for bad in (one(2048, 1, [5]), one(3, 2, [5])):
        for nbytes in (4, 5):
            with pytest.raises(ValueError):
                bad.packed(nbytes)
    rng = np.random.default_rng(3)
    far = _batch(rng, [4000], span=2_000_000_000_000)       # a block's offset would not fit 32 bits
    with pytest.raises(ValueError):
        far.packed(5)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_native_collate_equals_pack_events(dtype):
    """ep_collate_aos_host: the reference's per-sample (N,4) x,y,t,p arrays -> canonical SoA with integer tick stamps."""
    rng = np.random.default_rng(21)
    samples = []
    for n in (1000, 0, 1, 70000, 65536, 3):
        t_us = np.sort(rng.integers(0, 250_000, n))
        t = t_us / 1e6 if dtype == np.float64 else t_us.astype(np.float32)       # seconds (N-ImageNet) / raw ticks as fp32 (DDD17)
        samples.append(np.stack([rng.integers(0, 640, n), rng.integers(0, 480, n), t, rng.integers(0, 2, n)], 1).astype(dtype))
    scale = 1e6 if dtype == np.float64 else 1.0
    for threads in (1, 0):
        ev = ep.collate_events(samples, scale, pin=False, threads=threads)
        ref = ep.pack_events([np.column_stack([s[:, 0], s[:, 1], np.rint(s[:, 2].astype(np.float64) * scale), s[:, 3]]) for s in samples],
                             t_div=scale, pin=False)
        assert ev.t.dtype == torch.int64 and ev.t_div == scale
        assert torch.equal(ev.x, ref.x) and torch.equal(ev.y, ref.y) and torch.equal(ev.p, ref.p)
        assert np.array_equal(ev.t.numpy(), ref.t.numpy().astype(np.int64)) and np.array_equal(ev.offsets_host, ref.offsets_host)
    with pytest.raises(ValueError):
        ep.collate_events([np.array([[0.5, 1, 0, 1]], dtype)], scale, pin=False)           # fractional coordinate (after erase_and_add_events)
    with pytest.raises(ValueError):
        ep.collate_events([np.array([[1, 1, 0, -1]], dtype)], scale, pin=False)            # polarity -1: generic layout
    with pytest.raises(ValueError):
        ep.collate_events([np.array([[70000, 1, 0, 1]], dtype)], scale, pin=False)


def test_collate_mixed_dtypes_promote_to_float64():
    """A float32 first sample must not narrow the float64 samples behind it: stamps in seconds keep their microseconds."""
    rng = np.random.default_rng(22)
    t_us = np.sort(rng.integers(200_000_000, 200_250_000, 500))              # 200 s into the recording: fp32 would lose the us
    wide = np.stack([rng.integers(0, 640, 500), rng.integers(0, 480, 500), t_us / 1e6, rng.integers(0, 2, 500)], 1)
    narrow = np.array([[3, 4, 0.25, 1], [5, 6, 0.5, 0]], np.float32)
    ev = ep.collate_events([narrow, wide], 1e6, pin=False)
    lo = int(ev.offsets_host[1])
    assert np.array_equal(ev.t.numpy()[lo:], t_us) and np.array_equal(ev.t.numpy()[:lo], [250_000, 500_000])
    both32 = ep.collate_events([narrow, narrow], 1e6, pin=False)
    assert np.array_equal(both32.t.numpy(), [250_000, 500_000] * 2)


def test_native_compact_equals_the_numpy_rule():
    rng = np.random.default_rng(31)
    ev = _batch(rng, [3000, 0, 1, 70000, 2], span=2_000_000_000)
    a, b = ev.compact(native=True, threads=3), ev.compact(native=False)
    _same(a, b)
    assert a.x is ev.x and a.p is None and a.t.dtype == torch.uint32
    sh = ev.shard(1, 2)                                   # offsets[0] > 0: positions of the full arrays
    c, d = sh.compact(native=True), sh.compact(native=False)
    lo, hi = int(sh.offsets_host[0]), int(sh.offsets_host[-1])
    assert torch.equal(c.t[lo:hi], d.t[lo:hi]) and torch.equal(c.t_base, d.t_base)
    far = _batch(rng, [10], span=1 << 40)
    for native in (True, False):
        with pytest.raises(ValueError):
            far.compact(native=native)


def test_host_packers_accept_batches_of_empty_samples():
    ev = ep.from_soa(np.zeros(0, np.uint16), np.zeros(0, np.uint16), np.zeros(0, np.int64), np.zeros(0, np.uint8), np.array([0, 0, 0]),
                     t_div=1e6, pin=False)
    for make in (lambda: ev.packed(5), lambda: ev.packed(4), ev.compact, lambda: ev.packed(5, native=False), lambda: ev.compact(native=False)):
        r = make()
        assert r.batch == 2 and r.num_events == 0 and r.t_base.tolist() == [0, 0]
    c = ep.collate_events([np.zeros((0, 4)), np.zeros((0, 4))], pin=False)
    assert c.batch == 2 and c.num_events == 0 and c.offsets_host.tolist() == [0, 0, 0]


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_native_collate_rejections_inside_vector_groups(dtype):
    """Violations anywhere in a sample (not only in the scalar tail of the 4-wide AVX2 loop) are reported."""
    base = np.stack([np.arange(13) % 7, np.arange(13) % 5, np.arange(13) * 1e-6, np.arange(13) % 2], 1).astype(dtype)
    assert ep.collate_events([base], 1e6, pin=False).num_events == 13
    for row, col, val in ((2, 0, 0.5), (5, 1, -1.0), (6, 0, 65536.0), (1, 3, 2.0), (9, 3, -1.0), (3, 1, np.nan), (12, 0, 1.25)):
        bad = base.copy()
        bad[row, col] = val
        with pytest.raises(ValueError):
            ep.collate_events([base, bad], 1e6, pin=False)


def test_native_packers_fuzz_against_numpy():
    """Randomised small batches around the block boundaries (256 / 1024 events), with empty samples, duplicated stamps and local
    disorder: the native packers and the numpy rule agree on every word, or both refuse."""
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=60, deadline=None, derandomize=True, database=None)
    @given(st.lists(st.sampled_from([0, 1, 2, 255, 256, 257, 511, 513, 1023, 1024, 1025, 1500]), min_size=1, max_size=5),
           st.integers(0, 2**31 - 1), st.sampled_from([1, 40, 400, 100_000]), st.booleans())
    def run(counts, seed, span, disorder):
        rng = np.random.default_rng(seed)
        ev = _batch(rng, counts, span=span, t0=int(rng.integers(0, 1 << 40)))
        t = ev.t.numpy().copy()
        if disorder and len(t) > 8:
            i = int(rng.integers(0, len(t) - 4))
            t[i:i + 4] = t[i:i + 4][::-1]
        ev = ep.from_soa(ev.x.numpy(), ev.y.numpy(), t, ev.p.numpy(), ev.offsets_host, t_div=1e6, pin=False)
        for make in (lambda n: ev.packed(4, native=n), lambda n: ev.packed(5, native=n), lambda n: ev.compact(native=n)):
            try:
                ref = make(False)
            except ValueError:
                with pytest.raises(ValueError):
                    make(True)
                continue
            got = make(True)
            _same(got, ref)
            if got.y is None:                                   # packed forms decode back to the stamps they were given
                assert np.array_equal(got.unpack_host()[2], t)

    run()


def test_reshape_axis_multiplier_against_numpy(native_lib):
    """The host decision the whole-plane binning kernels take per call: an axis of the fused events_reshape
    (events_augment.py:22-26: x * scale in fp64, truncated by .long()) is computed as (x * multiplier) >> 32 only where that
    equals the reference's expression for every coordinate of the packed layouts; 0.35 (640 -> 224) has traps and keeps its
    table, 224 / 480 does not."""
    import ctypes
    x = np.arange(2048, dtype=np.float64)
    rng = np.random.default_rng(3)
    scales = [224 / 640, 224 / 480, 224 / 346, 224 / 260, 224 / 240, 224 / 180, 0.4, 0.5, 1.0, 1.5, 1e-6, 0.999999]
    scales += list(rng.uniform(0.01, 0.999, 40))
    seen = set()
    for s in scales:
        m = ctypes.c_uint32(0)
        ok = native_lib.ep_reshape_axis_multiplier_host(ctypes.c_double(s), ctypes.byref(m))
        exact = (x * np.float64(s)).astype(np.int64)                      # the reference's arithmetic
        if ok:
            got = (np.arange(2048, dtype=np.uint64) * np.uint64(m.value)) >> np.uint64(32)
            assert np.array_equal(got.astype(np.int64), exact), s
        else:
            assert m.value == 0
            if s < 1.0:
                mul = int(np.ceil(np.ldexp(np.float64(s), 32)))
                got = (np.arange(2048, dtype=np.uint64) * np.uint64(mul)) >> np.uint64(32)
                assert not np.array_equal(got.astype(np.int64), exact), s      # refused only where it would differ
        seen.add(bool(ok))
    m = ctypes.c_uint32(0)
    assert native_lib.ep_reshape_axis_multiplier_host(ctypes.c_double(224 / 640), ctypes.byref(m)) == 0      # 0.35 * 180 = 62.99999999999999
    assert native_lib.ep_reshape_axis_multiplier_host(ctypes.c_double(224 / 480), ctypes.byref(m)) == 1
    assert seen == {True, False}


def test_collate_transport_one_pass_equals_two_steps(native_lib):
    """ep_collate_transport4_host (rows -> 4 B packed layout in one pass) against ep_collate_aos_host + ep_pack_transport_host:
    the same words, block offsets, bases and offsets bit for bit on time-sorted samples (float64 and float32 rows, empty and
    one-event samples, samples that start and end inside a tick block); batches the one-pass form does not take (unsorted
    rows, sparse stamps, wide coordinates) fall back to the two-step form and its choice of layout."""
    rng = np.random.default_rng(21)

    def sample(n, dtype=np.float64, rate=4.0, W=640, H=480):
        t = np.sort(rng.integers(0, max(int(n / rate), 1), n)).astype(np.float64) / 1e6 + 0.25
        return np.stack([rng.integers(0, W, n), rng.integers(0, H, n), t, rng.integers(0, 2, n)], 1).astype(dtype)

    batch = [sample(n) for n in (5000, 0, 1, 257, 70001, 255, 4096)]
    one = ep.collate_transport(batch, pin=False, threads=3)
    two = ep.collate_events(batch, pin=False, threads=3).packed(4, threads=3)
    assert one.t is None and one.y is None and one.t_base is not None          # the 4 B layout
    for a, b in ((one.x, two.x), (one.p, two.p), (one.t_base, two.t_base), (one.offsets, two.offsets)):
        assert torch.equal(a, b)
    assert np.array_equal(one.offsets_host, two.offsets_host) and one.t_div == two.t_div
    # float32 rows (stamps with few digits so that float32 holds them)
    b32 = [np.stack([rng.integers(0, 346, n), rng.integers(0, 260, n), np.sort(rng.integers(0, n // 3 + 1, n)) / 1e3,
                     rng.integers(0, 2, n)], 1).astype(np.float32) for n in (3000, 700)]
    one, two = ep.collate_transport(b32, ticks_per_unit=1e3, pin=False), ep.collate_events(b32, ticks_per_unit=1e3, pin=False).packed(4)
    assert torch.equal(one.x, two.x) and torch.equal(one.p, two.p) and torch.equal(one.t_base, two.t_base)
    # fallbacks: an unsorted sample (its first row is not its smallest stamp), sparse stamps (5 B layout), x >= 2048 (8 B layout)
    uns = [sample(3000), sample(2000)[::-1].copy()]
    for bad in (uns, [sample(4000, rate=0.01)], [sample(3000, W=4000)]):
        got, ref = ep.collate_transport(bad, pin=False), ep.collate_events(bad, pin=False).transport()
        assert (got.t is None) == (ref.t is None) and (got.y is None) == (ref.y is None)
        assert torch.equal(got.x, ref.x) and torch.equal(got.offsets, ref.offsets)
        x1, y1, t1, p1 = got.unpack_host() if got.y is None else (None,) * 4
        if x1 is not None:
            x2, y2, t2, p2 = ref.unpack_host()
            assert np.array_equal(t1, t2) and np.array_equal(x1, x2) and np.array_equal(p1, p2)
