"""EventCollator: the DataLoader collate_fn that replaces per-sample CPU binning in the reference's datasets (no GPU needed)."""
import numpy as np
import pytest
import torch
from torch.utils.data import DataLoader, Dataset

import eventpretrain_b200 as ep


class _Windows(Dataset):
    """Stands in for a reference dataset whose __getitem__ returns the raw window (x, y, t [s], p) instead of a voxel grid."""

    def __init__(self, n_items, fractional=False):
        self.n, self.fractional = n_items, fractional

    def __len__(self):
        return self.n

    def __getitem__(self, i):
        rng = np.random.default_rng(100 + i)
        n = int(rng.integers(1, 4000)) if i % 5 else 0
        t = np.sort(rng.integers(0, max(1, n // 4), n)) / 1e6
        ev = np.stack([rng.integers(0, 640, n), rng.integers(0, 480, n), t, rng.integers(0, 2, n)], 1).astype(np.float64).reshape(-1, 4)
        if self.fractional and n:
            ev[0, 0] += 0.25
        return {"events": ev, "label": i % 3, "image_name": f"sample_{i:03d}"}


@pytest.mark.parametrize("workers", [0, 2])
@pytest.mark.parametrize("layout", ["transport", "canonical"])
def test_collator_in_a_dataloader(workers, layout):
    ds = _Windows(10)
    dl = DataLoader(ds, batch_size=4, shuffle=False, num_workers=workers, collate_fn=ep.EventCollator(layout=layout))
    seen = 0
    for batch in dl:
        ev = batch["events"]
        B = ev.batch
        assert isinstance(ev, ep.RaggedEvents) and B == len(batch["image_name"]) == batch["label"].numel()
        want = [ds[seen + b]["events"] for b in range(B)]
        assert ev.offsets_host.tolist() == np.cumsum([0] + [len(w) for w in want]).tolist() and ev.t_div == 1e6
        if ev.y is None:                                   # a packed transport layout: decode it
            x, y, t, p = ev.unpack_host()
        elif ev.t_base is not None:                        # compact
            rel = ev.t.numpy().astype(np.int64)
            x, y, p = ev.x.numpy(), ev.y.numpy(), (rel >> 31).astype(np.uint8)
            t = (rel & 0x7fffffff) + np.repeat(ev.t_base.numpy(), np.diff(ev.offsets_host))
        else:
            x, y, t, p = (a.numpy() for a in (ev.x, ev.y, ev.t, ev.p))
        cat = np.concatenate(want, 0)
        assert np.array_equal(x, cat[:, 0]) and np.array_equal(y, cat[:, 1]) and np.array_equal(p, cat[:, 3])
        assert np.array_equal(t, np.rint(cat[:, 2] * 1e6).astype(np.int64))
        assert batch["label"].tolist() == [(seen + b) % 3 for b in range(B)]
        seen += B
    assert seen == 10


def test_collator_falls_back_to_the_generic_layout():
    dl = DataLoader(_Windows(4, fractional=True), batch_size=4, collate_fn=ep.EventCollator())
    ev = next(iter(dl))["events"]
    assert ev.x.dtype == torch.float64 and ev.t_base is None and ev.batch == 4       # sub-pixel coordinates survive
