"""Patchify, the patchified diff-map target and its loss tail.

Drop-ins for utils/reshape.py:15-22 (frame2emb) and the target half of PrHubModel.reconstruct_loss
(model/pretrain/pr_hub_model.py:125-141); plus the pre-embed patch gather that lets PatchEmbed run on
visible patches only.
"""
import torch

from . import _lib
from ._runtime import contiguous_f32, lib, ptr, require_cuda, stream_ptr

ORDER = {"cpq": _lib.EP_ORDER_CPQ, "pqc": _lib.EP_ORDER_PQC}


def patchify_gather(x, patch, ids_keep=None, order="cpq"):
    """(B,C,H,W) -> (B,K,C*p*p): patches ids_keep[b,k] (all L patches in order when None).
    order 'cpq' = Conv2d(k=s=p) operand order (vit_block.py:44-68); 'pqc' = frame2emb order (reshape.py:19)."""
    require_cuda(x)
    x = contiguous_f32(x, "x")
    B, C, H, W = x.shape
    L = (H // patch) * (W // patch)
    K = L if ids_keep is None else ids_keep.shape[1]
    out = torch.empty((B, K, C * patch * patch), dtype=torch.float32, device=x.device)
    ids = None if ids_keep is None else ids_keep.contiguous()
    with torch.cuda.device(x.device):
        rc = lib().ep_patchify_gather(stream_ptr(x.device), x.data_ptr(), ptr(ids), B, C, H, W, patch, K, ORDER[order],
                                      out.data_ptr())
    _lib.check(rc, "ep_patchify_gather")
    return out


def frame2emb(patch_size, frame):
    """utils/reshape.py:15-22 — (b,c,H,W) -> (b, l, p*p*c), element order (ph, pw, c); square frames like the
    reference (it reads frame.shape[2] for both axes)."""
    return patchify_gather(frame, patch_size, None, "pqc")


def target_normpix(frame, patch_size, norm_pix_loss=True, eps=1.0e-6):
    """frame2emb + per-patch (x-mean)/sqrt(var+eps) in one pass   (pr_hub_model.py:126-131)."""
    require_cuda(frame)
    frame = contiguous_f32(frame, "frame")
    B, C, H, W = frame.shape
    L = (H // patch_size) * (W // patch_size)
    out = torch.empty((B, L, C * patch_size * patch_size), dtype=torch.float32, device=frame.device)
    with torch.cuda.device(frame.device):
        rc = lib().ep_patchify_normpix(stream_ptr(frame.device), frame.data_ptr(), B, C, H, W, patch_size,
                                       int(bool(norm_pix_loss)), float(eps), out.data_ptr())
    _lib.check(rc, "ep_patchify_normpix")
    return out


def target_patch_loss(pred, frame, patch_size, norm_pix_loss=True, eps=1.0e-6, mask=None):
    """Per-patch mean((pred - target)^2), target built on the fly and never written   (pr_hub_model.py:126-134).
    mask (B,L), 1 = removed patch: only those patches are read, normalised and compared (the loss discards the rest, :139);
    the others get 0.  No autograd: use for evaluation / monitoring; training keeps target_normpix + torch ops for the gradient."""
    require_cuda(pred, frame)
    frame = contiguous_f32(frame, "frame")
    pred = contiguous_f32(pred, "pred")
    B, C, H, W = frame.shape
    L = (H // patch_size) * (W // patch_size)
    if tuple(pred.shape) != (B, L, C * patch_size * patch_size):
        raise ValueError("pred must be (B, L, p*p*c)")
    out = torch.empty((B, L), dtype=torch.float32, device=frame.device)
    with torch.cuda.device(frame.device):
        if mask is None:
            rc = lib().ep_target_patch_loss(stream_ptr(frame.device), frame.data_ptr(), pred.data_ptr(), B, C, H, W,
                                            patch_size, int(bool(norm_pix_loss)), float(eps), out.data_ptr())
        else:
            mask = contiguous_f32(mask, "mask")
            if tuple(mask.shape) != (B, L):
                raise ValueError("mask must be (B, L)")
            rc = lib().ep_target_patch_loss_masked(stream_ptr(frame.device), frame.data_ptr(), pred.data_ptr(), mask.data_ptr(), B, C, H, W,
                                                   patch_size, int(bool(norm_pix_loss)), float(eps), out.data_ptr())
    _lib.check(rc, "ep_target_patch_loss")
    return out


def reconstruct_loss(self, reconstruct_pred, sub_frame, mask):
    """Drop-in for PrHubModel.reconstruct_loss (pr_hub_model.py:125-141): reads self.patch_size,
    self.norm_pix_loss, self.mask_ratio.  The target is built by one fused kernel; the subtraction and
    reductions stay in torch so gradients flow to reconstruct_pred."""
    target = target_normpix(sub_frame.float(), self.patch_size, self.norm_pix_loss)
    loss = ((reconstruct_pred - target) ** 2).mean(dim=-1)
    if self.mask_ratio == 0:
        return loss.mean()
    return (mask * loss).sum() / mask.sum()


def diffmap_frames(f0, f1, mode="linear", eps=1.0e-3, negate=None):
    """Frame-side temporal intensity-difference target g(f1) - g(f0), g = identity | log(.+eps); negate[b]
    flips the sign (time reversal, view_augment.py:60-63).  The reference ships no generator for its
    pre-computed sub_frame files, so this formula is this package's own."""
    require_cuda(f0, f1)
    f0 = contiguous_f32(f0, "f0")
    f1 = contiguous_f32(f1, "f1")
    out = torch.empty_like(f0)
    B = f0.shape[0] if f0.dim() == 4 else 1
    neg = None
    if negate is not None:
        neg = torch.as_tensor(negate, device=f0.device).to(torch.uint8).contiguous()
    with torch.cuda.device(f0.device):
        rc = lib().ep_diffmap_frames(stream_ptr(f0.device), f0.data_ptr(), f1.data_ptr(), out.data_ptr(), f0.numel(),
                                     f0.numel() // B, 1 if mode == "log" else 0, float(eps), ptr(neg))
    _lib.check(rc, "ep_diffmap_frames")
    return out
