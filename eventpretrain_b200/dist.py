"""Multi-GPU plumbing: shard a batch by sample, one process per GPU, no data-path collective.

Every stage of the path is per-sample, so ranks bin / mask / gather their own samples
independently — identical to the reference's DistributedSampler split (main_pretrain.py:215-220).  The
only collective is one small all-reduce of normalisation statistics (count, sum, sum of squares: SUM;
max: MAX), the analogue of misc.all_reduce_mean (utils/misc.py:406-414), issued on a side stream so it
never gates the binning kernels.  torch.distributed is the transport: NCCL over NVLink on GPUs, gloo in
the CPU tests.
"""
import datetime
import os

import torch
import torch.distributed as dist


def init_from_env(backend=None):
    """torchrun-style init (RANK / LOCAL_RANK / WORLD_SIZE / MASTER_ADDR / MASTER_PORT); returns (rank, world, local)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group(backend, device_id=torch.device("cuda", local), timeout=datetime.timedelta(seconds=180))
        else:
            dist.init_process_group(backend, timeout=datetime.timedelta(seconds=180))
    elif torch.cuda.is_available():
        torch.cuda.set_device(local)
    return rank, world, local


def shard_range(batch, rank, world):
    """Contiguous sample range of `rank`: [rank*B/G, (rank+1)*B/G)."""
    return (batch * rank) // world, (batch * (rank + 1)) // world


def balance_by_events(counts, world):
    """Greedy longest-first assignment of samples to ranks by event count (ragged batches); returns a list
    of sorted sample-index lists, one per rank.  Deterministic: ties go to the lowest rank."""
    order = sorted(range(len(counts)), key=lambda i: (-int(counts[i]), i))
    load = [0] * world
    parts = [[] for _ in range(world)]
    for i in order:
        r = min(range(world), key=lambda k: (load[k], k))
        parts[r].append(i)
        load[r] += int(counts[i])
    return [sorted(p) for p in parts]


def plane_statistics(x):
    """Per-channel (count, sum, sum of squares, max) of a CUDA (B,C,H,W) float32 tensor as fp64, shape (C,4): one native pass
    with a fixed reduction order (ep_plane_statistics).  ep.bin_events(..., stats=True) returns the same table for the voxel
    grid as a by-product of the binning kernels; this entry is for tensors produced elsewhere (count frames, targets)."""
    import ctypes  # noqa: F401
    from ._runtime import lib, require_cuda, stream_ptr, workspace
    from . import _lib
    require_cuda(x)
    if x.dim() != 4 or x.dtype != torch.float32 or not x.is_contiguous():
        raise ValueError("plane_statistics needs a contiguous float32 (B,C,H,W) tensor")
    B, C, H, W = x.shape
    L = lib()
    out = torch.empty((C, 4), dtype=torch.float64, device=x.device)
    ws = workspace(L.ep_plane_statistics_workspace_bytes(C), x.device, "stats")
    with torch.cuda.device(x.device):
        rc = L.ep_plane_statistics(stream_ptr(x.device), x.data_ptr(), B, C, H, W, out.data_ptr(), ws.data_ptr(), ws.numel())
    _lib.check(rc, "ep_plane_statistics")
    return out


def allreduce_statistics(stats, group=None, async_op=False):
    """SUM over ranks for columns 0..2, MAX for column 3; in place.  Returns (stats, works)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return stats, []
    sums = stats[:, :3].contiguous()
    mx = stats[:, 3].contiguous()
    w1 = dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=group, async_op=async_op)
    w2 = dist.all_reduce(mx, op=dist.ReduceOp.MAX, group=group, async_op=async_op)

    def finish():
        stats[:, :3] = sums
        stats[:, 3] = mx
        return stats

    if async_op:
        return finish, [w1, w2]
    return finish(), []


def finalize_statistics(stats):
    """(C,4) reduced statistics -> dict(mean, std, max) per channel (population std)."""
    n, s, q, mx = stats[:, 0], stats[:, 1], stats[:, 2], stats[:, 3]
    mean = s / n
    var = (q / n - mean * mean).clamp_min(0)
    return {"mean": mean, "std": var.sqrt(), "max": mx}
