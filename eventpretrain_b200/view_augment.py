"""View augmentation: drop-ins for dataset/augmentation/view_augment.py.

The random choices are drawn on the host from the global numpy RNG in exactly the reference's order (so a run
seeded like the reference produces the same crops and flips, and a voxel grid and its sub_frame augmented with the
same seed stay aligned); the tensor work — crop, F.interpolate, flips, sign — is one fused CUDA kernel.
"""
import math
from dataclasses import dataclass

import numpy as np
import torch

from . import _lib
from ._runtime import contiguous_f32, lib, require_cuda, stream_ptr


@dataclass
class ViewChoice:
    crop_x: int
    crop_y: int
    crop_w: int
    crop_h: int
    hflip: bool = False
    time_flip: bool = False
    negate: bool = False


def draw_crop(sensor_h, sensor_w, scale=(0.8, 1.0), ratio=(3 / 4, 4 / 3)):
    """view_crop's box (view_augment.py:9-33): up to 10 attempts; (0, 0, W, H) when none fits."""
    area = sensor_w * sensor_h
    for _ in range(10):
        target_area = np.random.uniform(*scale) * area
        aspect = np.random.uniform(sensor_w / sensor_h * ratio[0], sensor_w / sensor_h * ratio[1])
        cw = int(round(math.sqrt(target_area * aspect)))
        ch = int(round(math.sqrt(target_area / aspect)))
        if np.random.randint(0, 10) < 5:
            cw, ch = ch, cw
        if cw < sensor_w and ch < sensor_h:
            x0 = np.random.randint(0, sensor_w - cw)
            y0 = np.random.randint(0, sensor_h - ch)
            return x0, y0, cw, ch
    return 0, 0, sensor_w, sensor_h


def draw_evg_choice(args, shape, seed=None):
    """The RNG draws of evg_augment (view_augment.py:65-77): crop, horizontal flip, time flip."""
    if seed is not None:
        np.random.seed(seed)
    h, w = shape[-2], shape[-1]
    x0, y0, cw, ch = draw_crop(h, w, scale=(args.crop_min, 1))
    hflip = bool(np.random.random() < 0.5)
    tflip = bool(np.random.random() < 0.5)
    negate = tflip and args.num_bins in (5, 6)        # evg_time_flip (:54-55)
    return ViewChoice(x0, y0, cw, ch, hflip, tflip, negate)


def draw_frame_choice(args, shape, seed=None, time_flip_flag=False):
    """The RNG draws of frame_augment (view_augment.py:79-89): crop, horizontal flip; sign from the paired grid."""
    if seed is not None:
        np.random.seed(seed)
    h, w = shape[-2], shape[-1]
    x0, y0, cw, ch = draw_crop(h, w, scale=(args.crop_min, 1))
    hflip = bool(np.random.random() < 0.5)
    return ViewChoice(x0, y0, cw, ch, hflip, False, bool(time_flip_flag))


def prepare_views(choices, height, width, device):
    """One ViewChoice per sample -> the device array of ep_view_params that ep_view_augment reads (validated on the host:
    the kernel indexes the crop box directly).  Reusable: a fixed set of views (e.g. the full-frame resize of the MVSEC
    pair, ft_mvsec_dataset.py:229-239) is prepared once and passed to apply_views in place of the list."""
    B = len(choices)
    arr = np.zeros((B, 8), np.int32)
    arr[:, :7] = [(c.crop_x, c.crop_y, c.crop_w, c.crop_h, c.hflip, c.time_flip, c.negate) for c in choices]
    ok = ((arr[:, 2] > 0) & (arr[:, 3] > 0) & (arr[:, 0] >= 0) & (arr[:, 0] + arr[:, 2] <= width) & (arr[:, 1] >= 0)
          & (arr[:, 1] + arr[:, 3] <= height))
    if not ok.all():
        i = int(np.flatnonzero(~ok)[0])
        raise ValueError(f"apply_views: crop box {tuple(int(v) for v in arr[i, :4])} of sample {i} leaves the {height}x{width} frame")
    return torch.from_numpy(arr).to(device, non_blocking=True)


def apply_views(x, choices, size, mode="nearest"):
    """x (B,C,H,W) CUDA f32 + one ViewChoice per sample (or the array prepare_views made of them) ->
    (B,C,size[0],size[1]) in one launch (ep_view_augment)."""
    require_cuda(x)
    x = contiguous_f32(x, "x")
    B, C, H, W = x.shape
    if len(choices) != B:
        raise ValueError(f"apply_views: {len(choices)} view choices for a batch of {B}")
    prm = choices if isinstance(choices, torch.Tensor) else prepare_views(choices, H, W, x.device)
    if prm.dtype != torch.int32 or tuple(prm.shape) != (B, 8) or prm.device != x.device:
        raise ValueError("apply_views: prepared views must come from prepare_views() for this batch and device")
    out = torch.empty((B, C, int(size[0]), int(size[1])), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        rc = lib().ep_view_augment(stream_ptr(x.device), x.data_ptr(), B, C, H, W, prm.data_ptr(), int(size[0]), int(size[1]),
                                   {"nearest": 0, "bilinear": 1, "bicubic": 2}[mode], out.data_ptr())
    _lib.check(rc, "ep_view_augment")
    return out


def evg_augment(args, events_voxel_grid, size, mode="nearest", seed=None):
    """Drop-in for evg_augment(args, events_voxel_grid, size, mode, seed) -> (grid, time_flip_flag)
    (view_augment.py:65-77).  Accepts a CPU or CUDA (C,H,W) tensor and returns one on the same device."""
    choice = draw_evg_choice(args, events_voxel_grid.shape, seed)
    dev = events_voxel_grid.device
    from . import config
    x = events_voxel_grid.to(config.device()) if not events_voxel_grid.is_cuda else events_voxel_grid
    out = apply_views(x.unsqueeze(0).float(), [choice], size, mode)[0]
    return (out if dev.type == "cuda" else out.cpu()), choice.time_flip


def frame_augment(args, frame, seed=None, time_flip_flag=False):
    """Drop-in for frame_augment(args, frame, seed, time_flip_flag) (view_augment.py:79-89): bicubic to input_size."""
    choice = draw_frame_choice(args, frame.shape, seed, time_flip_flag)
    dev = frame.device
    from . import config
    x = frame.to(config.device()) if not frame.is_cuda else frame
    out = apply_views(x.unsqueeze(0).float(), [choice], (args.input_size, args.input_size), "bicubic")[0]
    return out if dev.type == "cuda" else out.cpu()
