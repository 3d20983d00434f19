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


@pytest.fixture(scope="module")
def ref_models():
    """The reference's model-side modules (timm is absent from the image: the 3-symbol stub of tests/golden/make_golden.py)."""
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    import make_golden
    saved = list(sys.path)
    try:
        return make_golden._import_reference()
    finally:
        sys.path[:] = saved


@pytest.mark.parametrize("seed", range(8))
def test_masking_and_target(ref_models, seed):
    """random_masking (random / density / anti-density), frame2emb and the norm_pix target of reconstruct_loss."""
    import torch
    from oracle import stage3_np as s3
    rng = np.random.default_rng(3000 + seed)
    p = [16, 8, 32][seed % 3]
    g = int(rng.integers(2, 8))
    B, C, S = int(rng.integers(1, 5)), int(rng.integers(1, 6)), g * p
    L = g * g
    ratio = [0.75, 0.5, 0.9][seed % 3]
    strategy = ["random", "density", "anti-density"][seed % 3 if seed < 6 else 0]
    x = torch.from_numpy(rng.standard_normal((B, C, S, S)).astype(np.float32))
    fake = SimpleNamespace(num_patches=L, mask_ratio=ratio, patch_size=p, args=SimpleNamespace(masking_strategy=strategy))
    torch.manual_seed(seed)
    ik, m, ir = ref_models.vit.ViT.random_masking(fake, x)                # model/backbone/vit.py:66-105, unbound
    if strategy == "random":
        torch.manual_seed(seed)
        noise = torch.rand(B, L).numpy()
    else:
        noise = s3.patch_density(x.numpy(), p) * (1.0 if strategy == "density" else -1.0)
        want = torch.nn.AvgPool2d(p, p)(abs(torch.sum(x, dim=1))).flatten(1).numpy() * (1.0 if strategy == "density" else -1.0)
        assert np.array_equal(noise, want)
    keep = s3.len_keep(L, ratio)
    gk, gm, gr = s3.mask_from_noise(noise, keep)
    clean = ~s3.rows_with_ties(noise)                                     # unstable argsort: parity is defined on tie-free rows
    assert np.array_equal(gk[clean], ik.numpy()[clean]) and np.array_equal(gm[clean], m.numpy()[clean])
    assert np.array_equal(gr[clean], ir.numpy()[clean])
    frame = torch.from_numpy(rng.standard_normal((B, 1, S, S)).astype(np.float32))
    emb = ref_models.reshape.frame2emb(p, frame)
    assert np.array_equal(s3.target_normpix(frame.numpy(), p, False), emb.numpy())
    tgt = (emb - emb.mean(dim=-1, keepdim=True)) / (emb.var(dim=-1, keepdim=True) + 1.0e-6) ** .5    # pr_hub_model.py:129-131
    got = s3.target_normpix(frame.numpy(), p, True)
    assert np.all(np.abs(got - tgt.numpy()) <= 1e-5 * np.abs(tgt.numpy()) + 1e-6)
