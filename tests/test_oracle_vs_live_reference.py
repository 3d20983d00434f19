"""The CPU oracle against the UNMODIFIED reference executed in place (only where /root/reference is mounted: the build
container; skipped on the GPU box, where the committed golden vectors of tests/golden/ stand in for it).

Randomised inputs beyond the committed fixtures: random grid sizes, bin counts, event counts, float64 / float32 rows,
second / microsecond stamps, polarity conventions {0,1} and {-1,1}, duplicated stamps, a sample with deltaT == 0."""
import os
import sys
from types import SimpleNamespace

import numpy as np
import pytest

REF = os.environ.get("EP_REFERENCE_ROOT", "/root/reference")
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "dataset", "dataset_utils")),
                                reason="reference tree not mounted (GPU box): golden vectors cover parity there")


@pytest.fixture(scope="module")
def ref():
    sys.path.insert(0, REF)
    try:
        from dataset.dataset_utils import events_to_image as eti
        from dataset.dataset_utils import events_to_voxel_grid as etv
    finally:
        sys.path.remove(REF)
    return SimpleNamespace(voxel=etv.events_to_voxel_grid, ecdp=eti.events_to_image_ecdp, mem=eti.events_to_image_mem,
                           hot=eti.remove_hot_pixel_mem, evrep=eti.events_to_EvRep)


def _events(rng, n, H, W, dtype, unit, pm1, dup):
    t = np.sort(rng.integers(0, max(2, n // 3) if dup else 400_000, n)).astype(np.float64) * unit
    p = rng.integers(0, 2, n).astype(np.float64)
    if pm1:
        p = p * 2 - 1
    return np.stack([rng.integers(0, W, n), rng.integers(0, H, n), t, p], 1).astype(dtype)


@pytest.mark.parametrize("seed", range(12))
def test_voxel_count_hotpixel(ref, seed):
    from oracle import events as oe
    rng = np.random.default_rng(1000 + seed)
    H, W = int(rng.integers(3, 70)), int(rng.integers(3, 90))
    bins = int(rng.integers(1, 17))
    n = int(rng.integers(1, 30_000))
    dtype = np.float32 if seed % 4 == 3 else np.float64
    unit = 1.0 if dtype == np.float32 or seed % 2 else 1e-6              # microsecond integers, or seconds
    ev = _events(rng, n, H, W, dtype, unit, pm1=seed % 3 == 2, dup=seed % 5 == 4)
    if seed == 7:
        ev[:, 2] = ev[0, 2]                                               # deltaT == 0 -> 1.0 (events_to_voxel_grid.py:24-25)
    args = SimpleNamespace(num_bins=bins)
    want = ref.voxel(args, ev.copy(), (H, W)).numpy()
    got = oe.voxel_grid(ev.copy(), bins, (H, W))
    assert got.dtype == np.float32 and np.array_equal(got, want), np.abs(got - want).max()
    assert np.array_equal(oe.count_frame(ev.copy(), (H, W), 2), ref.ecdp(args, ev.copy(), (H, W)).numpy())
    mem = ref.mem(args, ev.copy(), (H, W)).numpy()
    assert np.array_equal(oe.count_frame(ev.copy(), (H, W), 3), mem)
    import torch
    hot = ref.hot(torch.from_numpy(mem.copy()) / 255).numpy()
    assert np.array_equal(oe.remove_hot_pixel_mem(mem / np.float32(255)), hot)


@pytest.mark.parametrize("seed", range(6))
def test_evrep(ref, seed):
    from oracle import events as oe
    rng = np.random.default_rng(2000 + seed)
    H, W = int(rng.integers(4, 40)), int(rng.integers(4, 50))
    n = int(rng.integers(2, 8000))
    xs, ys = rng.integers(0, W, n).astype(np.int16), rng.integers(0, H, n).astype(np.int16)
    ts = np.sort(rng.uniform(0, 5e4 if seed % 2 else 0.05, n))
    if seed == 3:
        ts = rng.permutation(ts)                                          # unsorted input: lexsort decides the order
    ps = rng.integers(0, 2, n).astype(np.float64)
    want = ref.evrep(xs.copy(), ys.copy(), ts.copy(), ps.copy(), resolution=(W, H))
    got = oe.evrep(xs, ys, ts, ps, (W, H))
    assert got.shape == want.shape and np.array_equal(got, want, equal_nan=True)
