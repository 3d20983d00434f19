"""On-disk event formats -> ragged SoA batches, without the reference's float64 (N,4) detour.

The reference loaders build a float64 / float32 `(N,4)` x,y,t,p array per sample in `Dataset.__getitem__` and hand it to the
per-sample CPU binning.  On the batched path the samples go to the GPU as structure-of-arrays (x,y u16 | t ticks | p u8), so
the readers here produce exactly those arrays, with the timestamp convention (`t_div`, dtype) that reproduces the values the
reference's loader would have produced:

  read_nimagenet_npz   dataset/pretrain/pr_n_imagenet_dataset.py:45-56      t int64 µs, t_div = 1e6  (== float64(t) / 1e6)
  read_ddd17_memmap    dataset/finetune_semseg/ft_ddd17_dataset.py:76-102   t float32 ns (the reference casts the window to
                                                                            float32, so its time arithmetic is fp32: bin with
                                                                            time_f32=True)
  pack_soa             collate of per-sample SoA tuples into one RaggedEvents (pinned when CUDA is available)

File formats that need h5py / hdf5plugin (DSEC, MVSEC) are not covered: those modules are not in this image.
"""
import os

import numpy as np

from .events import from_soa


def read_nimagenet_npz(path, start=None, end=None):
    """One N-ImageNet sample: (x u16, y u16, t int64 µs, p u8), t_div = 1e6.  start/end select the window the reference
    picks with get_random_index (events_augment.py:5-20)."""
    ev = np.load(path)["event_data"]
    sl = slice(start, end)
    x = np.asarray(ev["x"][sl])
    y = np.asarray(ev["y"][sl])
    if x.size and (x.min() < 0 or y.min() < 0 or x.max() > 65535 or y.max() > 65535):
        raise ValueError("coordinates outside uint16")
    p = np.asarray(ev["p"][sl])
    return x.astype(np.uint16), y.astype(np.uint16), np.asarray(ev["t"][sl]).astype(np.int64), (p > 0).astype(np.uint8)


def read_ddd17_memmap(t_file, xyp_file, lo, hi):
    """Events [lo, hi) of a DDD17 recording: (x int16, y int16, t float32, p u8), t_div = 1: the reference's
    extract_events_from_memmap casts stamps (ns) and x,y,p to float32 (:96-97)."""
    n = os.path.getsize(t_file) // 8
    t = np.memmap(t_file, dtype="int64", mode="r", shape=(n,))
    xyp = np.memmap(xyp_file, dtype="int16", mode="r", shape=(n, 3))
    lo = max(int(lo), 0)
    w = np.asarray(xyp[lo:hi])
    return np.ascontiguousarray(w[:, 0]), np.ascontiguousarray(w[:, 1]), np.asarray(t[lo:hi]).astype(np.float32), \
        (w[:, 2] > 0).astype(np.uint8)


def pack_soa(samples, t_div=1.0, pin=True):
    """[(x, y, t, p), ...] -> RaggedEvents on the host (offsets by concatenation; dtypes are kept, so they must agree)."""
    counts = np.array([len(s[0]) for s in samples], np.int64)
    offsets = np.zeros(len(samples) + 1, np.int64)
    np.cumsum(counts, out=offsets[1:])
    cat = lambda i: np.concatenate([np.asarray(s[i]) for s in samples]) if samples else np.zeros(0)
    return from_soa(cat(0), cat(1), cat(2), cat(3), offsets, t_div=t_div, pin=pin)
