"""View augmentation (row f1): the host-side RNG mirror is checked on CPU against the reference's golden outputs
(using torch's own F.interpolate as the resampler); the fused CUDA kernel is checked on the GPU."""
from types import SimpleNamespace

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from synth import hash_uniform

MODES = ["nearest", "bilinear", "bicubic"]


def _inputs(name, c):
    if name.startswith("evg_nocrop"):
        return SimpleNamespace(crop_min=0.999, num_bins=5, input_size=8), torch.from_numpy(hash_uniform((5, 3, 3), 4200)), (8, 8), "nearest", 31
    if name.startswith("evg_"):
        bins = int(c["bins"])
        args = SimpleNamespace(crop_min=0.2, num_bins=bins, input_size=32)
        return args, torch.from_numpy(hash_uniform((bins, 60, 80), 4000 + bins)) * 8, (32, 32), MODES[int(c["mode"])], int(c["seed"])
    args = SimpleNamespace(crop_min=0.2, num_bins=5, input_size=32)
    return args, torch.from_numpy(hash_uniform((1, 60, 80), 4100)), (32, 32), "bicubic", int(c["seed"])


def _torch_apply(x, ch, size, mode):
    v = x[:, ch.crop_y: ch.crop_y + ch.crop_h, ch.crop_x: ch.crop_x + ch.crop_w]
    v = F.interpolate(v[None], size, mode=mode)[0]
    if ch.hflip:
        v = torch.flip(v, dims=[2])
    if ch.time_flip:
        v = torch.flip(v, dims=[0])
    return -v if ch.negate else v


def test_rng_mirror_matches_reference(golden_views):
    import eventpretrain_b200 as ep
    seen = set()
    for name, c in golden_views.items():
        args, x, size, mode, seed = _inputs(name, c)
        if name.startswith("frame_"):
            ch = ep.draw_frame_choice(args, x.shape, seed, bool(c["tflip"]))
        else:
            ch = ep.draw_evg_choice(args, x.shape, seed)
            assert ch.time_flip == bool(c["flag"]), name
        seen.add((ch.hflip, ch.time_flip, ch.negate, ch.crop_w == x.shape[2]))
        got = _torch_apply(x, ch, size, mode)
        assert np.array_equal(got.numpy(), c["out"]), name     # same crop / flips => torch reproduces the reference bit for bit
    assert len(seen) >= 5                                       # flips, no-flips, sign and the no-crop fallback all occur


@pytest.mark.gpu
def test_fused_kernel_vs_reference(golden_views, native_lib):
    import eventpretrain_b200 as ep
    for name, c in golden_views.items():
        args, x, size, mode, seed = _inputs(name, c)
        if name.startswith("frame_"):
            got = ep.frame_augment(args, x, seed=seed, time_flip_flag=bool(c["tflip"]))
        else:
            got, flag = ep.evg_augment(args, x, size, mode=mode, seed=seed)
            assert flag == bool(c["flag"])
        assert not got.is_cuda and tuple(got.shape) == c["out"].shape
        if mode == "nearest":
            assert np.array_equal(got.numpy(), c["out"]), name
        else:
            assert np.all(np.abs(got.numpy() - c["out"]) <= 1e-5 * np.abs(c["out"]) + 2e-6), (name, np.abs(got.numpy() - c["out"]).max())


@pytest.mark.gpu
def test_batched_views_and_shared_seed(native_lib):
    """A voxel grid and its sub_frame augmented with the same seed get the same crop / flip (the reference relies on
    this, pr_ef_imagenet_dataset.py:187-206); the batched call equals per-sample calls."""
    import eventpretrain_b200 as ep
    args = SimpleNamespace(crop_min=0.2, num_bins=5, input_size=224)
    grids = torch.from_numpy(hash_uniform((4, 5, 120, 160), 5)).cuda()
    choices = [ep.draw_evg_choice(args, grids.shape, seed=100 + i) for i in range(4)]
    fchoices = [ep.draw_frame_choice(args, grids.shape, seed=100 + i, time_flip_flag=choices[i].time_flip) for i in range(4)]
    for a, b in zip(choices, fchoices):
        assert (a.crop_x, a.crop_y, a.crop_w, a.crop_h, a.hflip) == (b.crop_x, b.crop_y, b.crop_w, b.crop_h, b.hflip)
        assert b.negate == a.time_flip
    out = ep.apply_views(grids, choices, (224, 224), "bilinear")
    for i in range(4):
        ref = _torch_apply(grids[i].cpu(), choices[i], (224, 224), "bilinear")
        assert torch.allclose(out[i].cpu(), ref, rtol=1e-5, atol=2e-6)
    # a crop box that leaves the frame, or a choice list that does not match the batch, is refused on the host
    bad = ep.ViewChoice(100, 0, 100, 50, False, False, False)
    with pytest.raises(ValueError):
        ep.apply_views(grids, [bad] * 4, (224, 224))
    with pytest.raises(ValueError):
        ep.apply_views(grids, choices[:3], (224, 224))


@pytest.mark.gpu
@pytest.mark.parametrize("shape,size", [((3, 9, 260, 346), (224, 224)), ((2, 5, 480, 640), (224, 224)), ((2, 3, 37, 53), (64, 80)),
                                        ((1, 2, 1200, 1600), (96, 96))])
def test_bilinear_staged_rows(native_lib, shape, size):
    """The bilinear kernel that stages a tile's source rows in shared memory (the MVSEC pair: org grid + 224x224 copy,
    ft_mvsec_dataset.py:229-239): full frames, odd crop offsets, flips, down- and up-scaling, and a frame whose rows exceed
    the staging buffer (direct loads inside the same kernel), against F.interpolate of the same crop."""
    import eventpretrain_b200 as ep
    B, C, H, W = shape
    x = torch.from_numpy(hash_uniform(shape, 11)).cuda()
    rng = np.random.default_rng(H)
    choices = [ep.ViewChoice(0, 0, W, H, False, False, False)]
    for i in range(1, B):
        cw, ch = int(rng.integers(W // 3, W)), int(rng.integers(H // 3, H))
        choices.append(ep.ViewChoice(int(rng.integers(0, W - cw + 1)) | 1 if W - cw > 1 else 0, int(rng.integers(0, H - ch + 1)), cw, ch,
                                     bool(i & 1), bool(i & 2), bool(i & 1)))
    choices = [c if c.crop_x + c.crop_w <= W else ep.ViewChoice(W - c.crop_w, c.crop_y, c.crop_w, c.crop_h, c.hflip, c.time_flip, c.negate)
               for c in choices]
    out = ep.apply_views(x, choices, size, "bilinear")
    for i in range(B):
        ref = _torch_apply(x[i].cpu(), choices[i], size, "bilinear")
        assert torch.allclose(out[i].cpu(), ref, rtol=1e-5, atol=2e-6), i


def test_event_stream_augmentation_rng_mirror(golden_stream_aug):
    """Row f2: erase_and_add_events / add_noise_events drop-ins reproduce the reference's seeded output bit for bit."""
    import eventpretrain_b200 as ep
    args = SimpleNamespace(num_bins=5)
    for name, c in golden_stream_aug.items():
        ev, sd = c["events"].copy(), int(c["seed"])
        if name.startswith("add_noise"):
            np.random.seed(sd)
            got = ep.add_noise_events(args, ev, (48, 64))
        else:
            size = (12, 16) if name.endswith("tiny") else (48, 64)
            got = ep.events_augment(args, ev, size=size, seed=sd)
        assert np.array_equal(got, c["aug"]), name


@pytest.mark.gpu
def test_augmented_streams_bin_like_reference(golden_stream_aug, native_lib):
    import eventpretrain_b200 as ep
    args = SimpleNamespace(num_bins=5)
    for name in ("erase_add_s1", "erase_add_s2"):
        c = golden_stream_aug[name]
        aug = ep.events_augment(args, c["events"].copy(), size=(48, 64), seed=int(c["seed"]))
        v = ep.events_to_voxel_grid(args, aug, (48, 64)).numpy()
        assert np.all(np.abs(v - c["voxel"]) <= 1e-5 * np.abs(c["voxel"]) + 1e-6)
        assert np.array_equal(ep.events_to_image_ecdp(args, aug, (48, 64)).numpy(), c["ecdp"])
