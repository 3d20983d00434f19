"""SURVEY.md §8 row f4 — on-disk readers: same values as the reference's loaders (formulas restated inline with file:line),
straight into the SoA layout."""
import numpy as np

import eventpretrain_b200 as ep


def test_nimagenet_npz(tmp_path):
    rng = np.random.default_rng(0)
    n = 5000
    ev = np.zeros(n, dtype=[("x", np.uint16), ("y", np.uint16), ("t", np.int64), ("p", np.bool_)])
    ev["x"], ev["y"] = rng.integers(0, 640, n), rng.integers(0, 480, n)
    ev["t"], ev["p"] = np.sort(rng.integers(0, 50_000, n)), rng.integers(0, 2, n).astype(bool)
    path = tmp_path / "n01440764_10026.npz"
    np.savez(path, event_data=ev)
    # the reference's load_events (dataset/pretrain/pr_n_imagenet_dataset.py:45-56)
    d = np.load(path)["event_data"]
    ref = np.vstack([d["x"], d["y"], d["t"], d["p"]]).T.astype(np.float64)
    ref[:, 2] = ref[:, 2] / 1e6
    x, y, t, p = ep.read_nimagenet_npz(str(path))
    assert (x.dtype, y.dtype, t.dtype, p.dtype) == (np.uint16, np.uint16, np.int64, np.uint8)
    assert np.array_equal(x, ref[:, 0]) and np.array_equal(y, ref[:, 1]) and np.array_equal(p, ref[:, 3])
    assert np.array_equal(t.astype(np.float64) / 1e6, ref[:, 2])          # what the kernels compute from (t, t_div = 1e6)
    xs, ys, ts, ps = ep.read_nimagenet_npz(str(path), 100, 1100)
    assert np.array_equal(xs, x[100:1100]) and np.array_equal(ts, t[100:1100])
    batch = ep.pack_soa([(x, y, t, p), (xs, ys, ts, ps)], t_div=1e6, pin=False)
    assert batch.batch == 2 and batch.num_events == n + 1000 and batch.t.dtype.is_floating_point is False
    assert batch.transport().x.dtype.itemsize == 4                          # the collate can ship it packed


def test_ddd17_memmap(tmp_path):
    rng = np.random.default_rng(1)
    n = 4000
    t = np.sort(rng.integers(0, 10**9, n)).astype(np.int64)
    xyp = np.stack([rng.integers(0, 346, n), rng.integers(0, 260, n), rng.integers(0, 2, n)], 1).astype(np.int16)
    tf, xf = tmp_path / "events.dat.t", tmp_path / "events.dat.xyp"
    t.tofile(tf)
    xyp.tofile(xf)
    lo, hi = 500, 2500
    # the reference's extract_events_from_memmap (dataset/finetune_semseg/ft_ddd17_dataset.py:84-102)
    ref = np.concatenate([np.array(t[lo:hi, None], dtype="float32"), np.array(xyp[lo:hi], dtype="float32")], -1)[:, [1, 2, 0, 3]]
    x, y, tt, p = ep.read_ddd17_memmap(str(tf), str(xf), lo, hi)
    assert tt.dtype == np.float32 and np.array_equal(tt, ref[:, 2])
    assert np.array_equal(x, ref[:, 0]) and np.array_equal(y, ref[:, 1]) and np.array_equal(p, ref[:, 3])
