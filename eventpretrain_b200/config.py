"""Process-wide settings of the drop-in layer."""
import torch

output_device = "cpu"     # "cpu": drop-ins return CPU tensors like the reference; "cuda": keep results on the GPU
_device = None


def device():
    """CUDA device used by the drop-in functions (one process per GPU: LOCAL_RANK's device by default)."""
    if not torch.cuda.is_available():
        raise RuntimeError("eventpretrain_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device()) if _device is None else _device


def set_device(dev):
    global _device
    _device = torch.device(dev)
