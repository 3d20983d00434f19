"""Device plumbing shared by the operators: streams, scratch buffers, argument checks."""
import torch

from . import _lib

_workspaces = {}


def require_cuda(*tensors):
    if not torch.cuda.is_available():
        raise RuntimeError("eventpretrain_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise ValueError("expected a CUDA tensor")


def stream_ptr(device=None):
    return torch.cuda.current_stream(device).cuda_stream


def workspace(nbytes, device, tag="default"):
    """Grow-only scratch buffer per (device, stream, tag); stream-ordered reuse is safe."""
    key = (torch.device(device).index, torch.cuda.current_stream(device).cuda_stream, tag)
    buf = _workspaces.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)
        _workspaces[key] = buf
    return buf


def ptr(t):
    return 0 if t is None else t.data_ptr()


def contiguous_f32(t, name):
    if t.dtype != torch.float32:
        raise TypeError(f"{name} must be float32, got {t.dtype}")
    return t.contiguous()


def lib():
    return _lib.load()
