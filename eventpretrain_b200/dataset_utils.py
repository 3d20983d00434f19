"""Drop-in replacements for dataset/dataset_utils/{events_to_voxel_grid,events_to_image}.py.

Same names, argument order and meaning, return types and error behaviour as the reference, so the
reference's dataset classes can import these instead:

    from eventpretrain_b200.dataset_utils import events_to_voxel_grid, events_to_image_ecdp, ...

`events` is the reference's numpy (N,4) x,y,t,p array (float64 or float32); results come back as CPU
tensors like the reference's, unless `eventpretrain_b200.config.output_device = "cuda"` keeps them on
the GPU for the batched path.  All arithmetic runs in the CUDA kernels behind the C ABI
(include/eventpretrain_b200.h); there is no CPU fallback.
"""
import numpy as np
import torch

from . import config
from .events import bin_events_aos, evrep, from_soa, mem_hotpixel


def _upload(events, is_txyp):
    if isinstance(events, torch.Tensor):
        t = events
    else:
        events = np.asarray(events)
        if events.dtype not in (np.float64, np.float32):
            events = events.astype(np.float64)
        t = torch.from_numpy(np.ascontiguousarray(events))
    assert t.dim() == 2 and t.shape[1] == 4          # events_to_voxel_grid.py:9, events_to_image.py:11
    if is_txyp:                                      # t,x,y,p column order -> x,y,t,p
        t = t[:, [1, 2, 0, 3]]
    return t.to(config.device(), non_blocking=True).contiguous()


def _ret(t):
    return t if config.output_device == "cuda" else t.cpu()


def events_to_voxel_grid(args, events, size, is_txyp=False):
    """dataset/dataset_utils/events_to_voxel_grid.py:4-61 — reads args.num_bins only."""
    out = bin_events_aos(_upload(events, is_txyp), (size[0], size[1]), num_bins=args.num_bins)
    return _ret(out["voxel"])


def events_to_image_ecdp(args, events, size, is_txyp=False):
    """dataset/dataset_utils/events_to_image.py:6-32 — (2,H,W) [pos, neg] counts as float32."""
    out = bin_events_aos(_upload(events, is_txyp), (size[0], size[1]), count_channels=2)
    return _ret(out["count"])


def events_to_image_mem(args, events, size, is_txyp=False):
    """dataset/dataset_utils/events_to_image.py:35-62 — (3,H,W) [pos, 0, neg]."""
    out = bin_events_aos(_upload(events, is_txyp), (size[0], size[1]), count_channels=3)
    return _ret(out["count"])


def remove_hot_pixel_mem(hist, num_stds=10):
    """dataset/dataset_utils/events_to_image.py:65-75 — in place, returns hist."""
    if hist.is_cuda:
        return mem_hotpixel(hist, num_stds=num_stds)
    dev = hist.to(config.device()).contiguous()
    mem_hotpixel(dev, num_stds=num_stds)
    hist.copy_(dev)
    return hist


def events_to_EvRep(event_xs, event_ys, event_timestamps, event_polarities, resolution=(320, 240)):
    """dataset/dataset_utils/events_to_image.py:77-125 — resolution is (W, H); returns numpy (3,H,W) float64."""
    width, height = resolution
    n = len(event_xs)
    if n == 0:
        raise IndexError("index 0 is out of bounds for axis 0 with size 0")   # sorted_timestamps[0], :110
    ev = from_soa(np.asarray(event_xs), np.asarray(event_ys), np.asarray(event_timestamps, np.float64),
                  np.asarray(event_polarities, np.float64), np.array([0, n], np.int64)).to(config.device())
    out = evrep(ev, (height, width), check=True)[0]
    return out if config.output_device == "cuda" else out.cpu().numpy()
