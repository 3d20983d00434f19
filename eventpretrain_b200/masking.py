"""Stage 3: random / density masking, ConvViT block masks, Swin apply_mask, token gathers.

Drop-ins for the first lines of backbone.forward(x, mask=True):
  model/backbone/vit.py:66-115, convvit.py:85-159, swin.py:113-179, model/pretrain/pr_rec_decoder.py:56-62
"""
import torch

from . import _lib
from ._runtime import contiguous_f32, lib, ptr, require_cuda, stream_ptr


def len_keep_of(L, mask_ratio):
    return int(L * (1 - mask_ratio))     # vit.py:75


def mask_from_noise(noise, len_keep):
    """noise (B,L) f32 -> ids_keep (B,len_keep) i64, mask (B,L) f32 (1 = removed), ids_restore (B,L) i64."""
    require_cuda(noise)
    noise = contiguous_f32(noise, "noise")
    B, L = noise.shape
    dev = noise.device
    ids_keep = torch.empty((B, len_keep), dtype=torch.int64, device=dev)
    mask = torch.empty((B, L), dtype=torch.float32, device=dev)
    ids_restore = torch.empty((B, L), dtype=torch.int64, device=dev)
    with torch.cuda.device(dev):
        rc = lib().ep_mask_from_noise(stream_ptr(dev), noise.data_ptr(), B, L, len_keep, ptr(ids_keep), mask.data_ptr(),
                                      ids_restore.data_ptr())
    _lib.check(rc, "ep_mask_from_noise")
    return ids_keep, mask, ids_restore


def patch_density(x, patch, sign=1.0):
    """AvgPool2d(p,p)(abs(x.sum(1))).flatten(1) * sign   (vit.py:80-83)."""
    require_cuda(x)
    x = contiguous_f32(x, "x")
    B, C, H, W = x.shape
    out = torch.empty((B, (H // patch) * (W // patch)), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        rc = lib().ep_patch_density(stream_ptr(x.device), x.data_ptr(), B, C, H, W, patch, float(sign), out.data_ptr())
    _lib.check(rc, "ep_patch_density")
    return out


def random_masking(self, x, x_org=None, mask_ratio=None):
    """Drop-in for backbone.random_masking (vit.py:66-105; Swin's variant takes (x, x_org, mask_ratio),
    swin.py:113-152).  Reads self.num_patches, self.mask_ratio, self.patch_size, self.args.masking_strategy.
    `torch.rand` stays the noise source so masks follow the caller's torch RNG stream exactly."""
    B = x.shape[0]
    L = self.num_patches
    ratio = self.mask_ratio if mask_ratio is None else mask_ratio
    keep = len_keep_of(L, ratio)
    strategy = self.args.masking_strategy
    if strategy == "random":
        noise = torch.rand(B, L, device=x.device)
    elif strategy in ("density", "anti-density"):
        src, p = (x_org, 32) if x_org is not None else (x, self.patch_size)   # swin.py:129 hard-codes 32
        noise = patch_density(src, p, 1.0 if strategy == "density" else -1.0)
    else:
        raise ValueError
    return mask_from_noise(noise, keep)


def gather_tokens(tokens, ids_keep, pos_embed=None):
    """(tokens + pos_embed) gathered by ids_keep: (B,L,D) -> (B,K,D)   (vit.py:113-115)."""
    require_cuda(tokens, ids_keep)
    tokens = contiguous_f32(tokens, "tokens")
    B, L, D = tokens.shape
    K = ids_keep.shape[1]
    pos = None
    if pos_embed is not None:
        pos = contiguous_f32(pos_embed.reshape(-1, D), "pos_embed")
        if pos.shape[0] != L:
            raise ValueError("pos_embed must be (L, D) or (1, L, D)")
    out = torch.empty((B, K, D), dtype=torch.float32, device=tokens.device)
    ids = ids_keep.contiguous()
    with torch.cuda.device(tokens.device):
        rc = lib().ep_gather_tokens(stream_ptr(tokens.device), tokens.data_ptr(), ptr(pos), ids.data_ptr(), B, L, K, D,
                                    out.data_ptr())
    _lib.check(rc, "ep_gather_tokens")
    return out


def unshuffle_tokens(emb, mask_token, ids_restore, pos_embed=None):
    """cat([emb, mask_token...]) gathered by ids_restore, + pos_embed   (pr_rec_decoder.py:56-62)."""
    require_cuda(emb, ids_restore)
    emb = contiguous_f32(emb, "emb")
    B, K, D = emb.shape
    L = ids_restore.shape[1]
    mt = contiguous_f32(mask_token.reshape(-1), "mask_token")
    pos = None if pos_embed is None else contiguous_f32(pos_embed.reshape(-1, D), "pos_embed")
    out = torch.empty((B, L, D), dtype=torch.float32, device=emb.device)
    with torch.cuda.device(emb.device):
        rc = lib().ep_unshuffle_tokens(stream_ptr(emb.device), emb.data_ptr(), mt.data_ptr(), ptr(pos),
                                       ids_restore.contiguous().data_ptr(), B, L, K, D, out.data_ptr())
    _lib.check(rc, "ep_unshuffle_tokens")
    return out


def block_mask_expand(mask, grid, rep, invert=True):
    """(B, grid*grid) -> (B,1,grid*rep,grid*rep); invert=True gives the `1 - mask` ConvBlock receives
    (convvit.py:129-133,142-146)."""
    require_cuda(mask)
    mask = contiguous_f32(mask, "mask")
    B = mask.shape[0]
    out = torch.empty((B, 1, grid * rep, grid * rep), dtype=torch.float32, device=mask.device)
    with torch.cuda.device(mask.device):
        rc = lib().ep_block_mask_expand(stream_ptr(mask.device), mask.data_ptr(), B, grid, rep, int(invert), out.data_ptr())
    _lib.check(rc, "ep_block_mask_expand")
    return out


def convvit_keep_masks(mask):
    """The two keep-masks of ConvViT.forward(mask=True): (B,1,56,56) and (B,1,28,28) for a 14x14 mask."""
    g = int(round(mask.shape[1] ** 0.5))
    return block_mask_expand(mask, g, 4), block_mask_expand(mask, g, 2)


def swin_apply_mask(x, mask, patches_resolution, n_vis=None):
    """Drop-in for SwinTransformer.apply_mask(x, mask, patches_resolution)   (swin.py:154-179).

    mask: (B', Mh*Mw) bool or float, row 0 is used for the whole batch.  Returns (x_vis (B,n_vis,C),
    coords (1,n_vis,2) int64, vis_mask (1,N) bool).  n_vis (tokens kept) is data dependent; pass it when
    known on the host (len_keep * rep^2 for masks from random_masking) to avoid a device sync.
    """
    require_cuda(x, mask)
    x = contiguous_f32(x, "x")
    B, N, C = x.shape
    H, W = patches_resolution
    M = mask.shape[1]
    up = N // M
    assert up * M == N
    rep = int(up ** 0.5)
    Mh, Mw = H // rep, W // rep
    row = mask[:1].to(torch.float32).contiguous()
    dev = x.device
    sync = n_vis is None
    cap = N if sync else int(n_vis)
    coords = torch.empty((cap, 2), dtype=torch.int64, device=dev)
    vis = torch.empty((N,), dtype=torch.uint8, device=dev)
    count = torch.empty((1,), dtype=torch.int32, device=dev)
    x_vis = torch.empty((B, cap, C), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        rc = lib().ep_swin_apply_mask(stream_ptr(dev), x.data_ptr(), row.data_ptr(), B, Mh, Mw, rep, C, cap,
                                      x_vis.data_ptr(), coords.data_ptr(), vis.data_ptr(), count.data_ptr())
    _lib.check(rc, "ep_swin_apply_mask")
    if sync:
        nv = int(count.item())
        if nv != cap:
            # rows were packed with stride nv by the kernel; reinterpret the prefix
            x_vis = x_vis.reshape(-1)[: B * nv * C].reshape(B, nv, C)
            coords = coords[:nv]
    return x_vis, coords.unsqueeze(0), vis.bool().unsqueeze(0)
