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
