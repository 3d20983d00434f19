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


# ---- token permutations with a gradient ---------------------------------------------------------------------------------------
# The reference trains through these index ops under torch.cuda.amp.autocast (pr_trainer.py:26-36), so every drop-in below is a
# torch.autograd.Function: forward and backward are the library's kernels (fp32 on the device); half / bfloat16 activations are
# widened on the way in and the result is returned in the activation's dtype (a permutation is exact in either direction, the
# positional add is then done in fp32 and rounded once, where eager autocast would add in the narrow type).

def _f32(t):
    return t.contiguous() if t.dtype == torch.float32 else t.float().contiguous()


def _act_dtype(*ts):
    for t in ts:
        if t is not None and t.dtype in (torch.float16, torch.bfloat16):
            return t.dtype
    return torch.float32


def _check_float(t, name):
    if t.dtype not in (torch.float32, torch.float16, torch.bfloat16):
        raise TypeError(f"{name} must be float32, float16 or bfloat16, got {t.dtype}")


def _scatter_add_tokens(grad, ids, shared, L):
    """grad (B,K,D) f32 -> (B,L,D): rows added at ids (B,K) or the batch-shared ids (K,)."""
    B, K, D = grad.shape
    out = torch.empty((B, L, D), dtype=torch.float32, device=grad.device)
    with torch.cuda.device(grad.device):
        rc = lib().ep_scatter_add_tokens(stream_ptr(grad.device), grad.data_ptr(), ids.data_ptr(), int(shared), B, L, K, D, out.data_ptr())
    _lib.check(rc, "ep_scatter_add_tokens")
    return out


def _sum_over_batch(x):
    """(B, ...) f32 -> (...): the sum over the batch in a fixed order."""
    B = x.shape[0]
    n = x[0].numel()
    out = torch.empty(x.shape[1:], dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        rc = lib().ep_sum_over_batch(stream_ptr(x.device), x.data_ptr(), B, n, out.data_ptr())
    _lib.check(rc, "ep_sum_over_batch")
    return out


class _GatherTokens(torch.autograd.Function):
    @staticmethod
    def forward(ctx, tokens, ids, pos):
        dt = _act_dtype(tokens)
        tk = _f32(tokens)
        B, L, D = tk.shape
        ids = ids.contiguous()
        shared = ids.dim() == 1
        K = ids.shape[-1]
        ids_b = ids.unsqueeze(0).expand(B, K).contiguous() if shared else ids          # the forward kernel takes (B,K)
        ps = None if pos is None else _f32(pos.reshape(-1, D))
        if ps is not None and ps.shape[0] != L:
            raise ValueError("pos_embed must be (L, D) or (1, L, D)")
        out = torch.empty((B, K, D), dtype=torch.float32, device=tk.device)
        with torch.cuda.device(tk.device):
            rc = lib().ep_gather_tokens(stream_ptr(tk.device), tk.data_ptr(), ptr(ps), ids_b.data_ptr(), B, L, K, D, out.data_ptr())
        _lib.check(rc, "ep_gather_tokens")
        ctx.save_for_backward(ids)
        ctx.meta = (L, shared, tokens.dtype, None if pos is None else (pos.shape, pos.dtype))
        return out.to(dt)

    @staticmethod
    def backward(ctx, g):
        (ids,) = ctx.saved_tensors
        L, shared, tdt, pmeta = ctx.meta
        need_t, need_p = ctx.needs_input_grad[0], ctx.needs_input_grad[2] and pmeta is not None
        gt = gp = None
        if need_t or need_p:
            gt32 = _scatter_add_tokens(_f32(g), ids, shared, L)
            if need_p:
                gp = _sum_over_batch(gt32).reshape(pmeta[0]).to(pmeta[1])
            if need_t:
                gt = gt32.to(tdt)
        return gt, None, gp


def gather_tokens(tokens, ids_keep, pos_embed=None):
    """(tokens + pos_embed) gathered by ids_keep: (B,L,D) -> (B,K,D)   (vit.py:113-115).  ids_keep (B,K), or (K,) shared by the
    batch (GroupingModule).  Differentiable in tokens and pos_embed."""
    require_cuda(tokens, ids_keep)
    _check_float(tokens, "tokens")
    if pos_embed is not None:
        _check_float(pos_embed, "pos_embed")
    return _GatherTokens.apply(tokens, ids_keep, pos_embed)


class _UnshuffleTokens(torch.autograd.Function):
    @staticmethod
    def forward(ctx, emb, mask_token, ids_restore, pos):
        dt = _act_dtype(emb)
        e = _f32(emb)
        B, K, D = e.shape
        L = ids_restore.shape[1]
        mt = _f32(mask_token.reshape(-1))
        ps = None if pos is None else _f32(pos.reshape(-1, D))
        ids = ids_restore.contiguous()
        out = torch.empty((B, L, D), dtype=torch.float32, device=e.device)
        with torch.cuda.device(e.device):
            rc = lib().ep_unshuffle_tokens(stream_ptr(e.device), e.data_ptr(), mt.data_ptr(), ptr(ps), ids.data_ptr(), B, L, K, D,
                                           out.data_ptr())
        _lib.check(rc, "ep_unshuffle_tokens")
        ctx.save_for_backward(ids)
        ctx.meta = (K, emb.dtype, mask_token.shape, mask_token.dtype, None if pos is None else (pos.shape, pos.dtype))
        return out.to(dt)

    @staticmethod
    def backward(ctx, g):
        (ids,) = ctx.saved_tensors
        K, edt, mshape, mdt, pmeta = ctx.meta
        g32 = _f32(g)
        B, L, D = g32.shape
        ge = gm = gp = None
        if ctx.needs_input_grad[0] or ctx.needs_input_grad[1]:
            ge32 = torch.empty((B, max(K, 1), D), dtype=torch.float32, device=g.device)
            gm32 = torch.empty((D,), dtype=torch.float32, device=g.device)
            scratch = torch.empty((B, D), dtype=torch.float32, device=g.device)
            with torch.cuda.device(g.device):
                rc = lib().ep_unshuffle_tokens_bwd(stream_ptr(g.device), g32.data_ptr(), ids.data_ptr(), B, L, K, D, ge32.data_ptr(),
                                                   gm32.data_ptr(), scratch.data_ptr())
            _lib.check(rc, "ep_unshuffle_tokens_bwd")
            ge = ge32[:, :K].to(edt)
            gm = gm32.reshape(mshape).to(mdt)
        if pmeta is not None and ctx.needs_input_grad[3]:
            gp = _sum_over_batch(g32).reshape(pmeta[0]).to(pmeta[1])
        return ge, gm, None, gp


def unshuffle_tokens(emb, mask_token, ids_restore, pos_embed=None):
    """cat([emb, mask_token...]) gathered by ids_restore, + pos_embed   (pr_rec_decoder.py:56-62).  Differentiable in emb,
    mask_token and pos_embed."""
    require_cuda(emb, ids_restore)
    _check_float(emb, "emb")
    _check_float(mask_token, "mask_token")
    if mask_token.numel() != emb.shape[-1]:
        raise ValueError("mask_token must hold D values")
    return _UnshuffleTokens.apply(emb, mask_token, ids_restore, pos_embed)


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


class _SwinApplyMask(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, row, H, W, rep, cap, sync):
        dt = _act_dtype(x)
        xf = _f32(x)
        B, N, C = xf.shape
        dev = xf.device
        alloc = torch.empty if sync else torch.zeros          # a caller-supplied n_vis that is too large leaves zeros, never garbage
        coords = alloc((cap, 2), dtype=torch.int64, device=dev)
        vis = torch.empty((N,), dtype=torch.uint8, device=dev)
        count = torch.empty((1,), dtype=torch.int32, device=dev)
        x_vis = alloc((B, cap, C), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            rc = lib().ep_swin_apply_mask(stream_ptr(dev), xf.data_ptr(), row.data_ptr(), B, H // rep, W // rep, rep, C, cap,
                                          x_vis.data_ptr(), coords.data_ptr(), vis.data_ptr(), count.data_ptr())
        _lib.check(rc, "ep_swin_apply_mask")
        if sync:
            nv = int(count.item())
            if nv != cap:
                # rows were packed with stride nv by the kernel; reinterpret the prefix
                x_vis = x_vis.reshape(-1)[: B * nv * C].reshape(B, nv, C)
                coords = coords[:nv]
        ctx.meta = (N, W, x.dtype)
        ctx.save_for_backward(coords)
        ctx.mark_non_differentiable(coords, vis, count)
        return x_vis.to(dt), coords, vis, count

    @staticmethod
    def backward(ctx, g, _gc, _gv, _gn):
        (coords,) = ctx.saved_tensors
        N, W, xdt = ctx.meta
        ids = (coords[:, 0] * W + coords[:, 1]).contiguous()
        return _scatter_add_tokens(_f32(g), ids, True, N).to(xdt), None, None, None, None, None, None


def swin_apply_mask(x, mask, patches_resolution, n_vis=None, check=False):
    """Drop-in for SwinTransformer.apply_mask(x, mask, patches_resolution)   (swin.py:154-179).

    mask: (B', Mh*Mw) bool or float, row 0 is used for the whole batch.  Returns (x_vis (B,n_vis,C),
    coords (1,n_vis,2) int64, vis_mask (1,N) bool).  n_vis (tokens kept) is data dependent; pass it when
    known on the host (len_keep * rep^2 for masks from random_masking) to avoid a device sync.  A caller-supplied n_vis must
    be the true count: one that is too large leaves the rows packed at the true stride followed by zeros (the buffers are
    zero-filled, nothing uninitialised is ever returned), one that is too small truncates — check=True compares it with the
    device count (one sync) and raises ValueError on a mismatch.
    Differentiable in x."""
    require_cuda(x, mask)
    _check_float(x, "x")
    B, N, C = x.shape
    H, W = patches_resolution
    M = mask.shape[1]
    up = N // M
    assert up * M == N
    rep = int(up ** 0.5)
    row = mask[:1].to(torch.float32).contiguous()
    sync = n_vis is None
    cap = N if sync else int(n_vis)
    x_vis, coords, vis, count = _SwinApplyMask.apply(x, row, H, W, rep, cap, sync)
    if not sync and check and int(count.item()) != cap:
        raise ValueError(f"swin_apply_mask: n_vis={cap} but the mask keeps {int(count.item())} tokens")
    return x_vis, coords.unsqueeze(0), vis.bool().unsqueeze(0)


class _SwinScatterDense(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, coords2, G):
        dt = _act_dtype(x)
        xf = _f32(x)
        B, n, C = xf.shape
        out = torch.empty((B, C, G, G), dtype=torch.float32, device=xf.device)
        with torch.cuda.device(xf.device):
            rc = lib().ep_swin_scatter_dense(stream_ptr(xf.device), xf.data_ptr(), coords2.data_ptr(), B, n, G, C, out.data_ptr())
        _lib.check(rc, "ep_swin_scatter_dense")
        ctx.save_for_backward(coords2)
        ctx.meta = (G, x.dtype)
        return out.to(dt)

    @staticmethod
    def backward(ctx, g):
        (coords2,) = ctx.saved_tensors
        G, xdt = ctx.meta
        g32 = _f32(g)
        B, C = g32.shape[:2]
        ids = (coords2[:, 0] * G + coords2[:, 1]).unsqueeze(0).expand(B, -1).contiguous()
        out = torch.empty((B, ids.shape[1], C), dtype=torch.float32, device=g.device)
        with torch.cuda.device(g.device):
            rc = lib().ep_gather_tokens_nchw(stream_ptr(g.device), g32.data_ptr(), ids.data_ptr(), B, G * G, ids.shape[1], C, out.data_ptr())
        _lib.check(rc, "ep_gather_tokens_nchw")
        return out.to(xdt), None, None


def swin_scatter_dense(x, coords, grid):
    """Stage tokens back onto the dense grid, channels first   (swin.py:221-225 and the i == 1, 2 copies):
        _emb = zeros(B, G*G, C); _emb[:, coords[0,:,0]*G + coords[0,:,1], :] = x.float(); _emb.reshape(B,G,G,C).permute(0,3,1,2)
    x (B,n_vis,C), coords (1,n_vis,2) or (n_vis,2) int64 (h,w) -> (B,C,G,G), zero where no token sits.  Differentiable in x."""
    require_cuda(x, coords)
    _check_float(x, "x")
    c2 = coords.reshape(-1, 2).to(torch.int64).contiguous()
    if c2.shape[0] != x.shape[1]:
        raise ValueError("coords must hold one (h, w) pair per token")
    return _SwinScatterDense.apply(x, c2, int(grid))


class _GatherTokensNCHW(torch.autograd.Function):
    @staticmethod
    def forward(ctx, feat, ids):
        dt = _act_dtype(feat)
        f = _f32(feat)
        B, D = f.shape[:2]
        L = f[0, 0].numel()
        ids = ids.contiguous()
        K = ids.shape[1]
        out = torch.empty((B, K, D), dtype=torch.float32, device=f.device)
        with torch.cuda.device(f.device):
            rc = lib().ep_gather_tokens_nchw(stream_ptr(f.device), f.data_ptr(), ids.data_ptr(), B, L, K, D, out.data_ptr())
        _lib.check(rc, "ep_gather_tokens_nchw")
        ctx.save_for_backward(ids)
        ctx.meta = (feat.shape, feat.dtype)
        return out.to(dt)

    @staticmethod
    def backward(ctx, g):
        (ids,) = ctx.saved_tensors
        shape, fdt = ctx.meta
        g32 = _f32(g)
        B, K, D = g32.shape
        L = 1
        for v in shape[2:]:
            L *= v
        out = torch.empty(shape, dtype=torch.float32, device=g.device)
        with torch.cuda.device(g.device):
            rc = lib().ep_scatter_add_tokens_nchw(stream_ptr(g.device), g32.data_ptr(), ids.data_ptr(), B, L, K, D, out.data_ptr())
        _lib.check(rc, "ep_scatter_add_tokens_nchw")
        return out.to(fdt), None


def gather_tokens_nchw(feat, ids_keep):
    """feat.flatten(2).permute(0,2,1) gathered by ids_keep along the tokens   (swin.py:226-228): feat (B,D,Gh,Gw) as the stage
    decoder's convolution leaves it, ids_keep (B,K) -> (B,K,D) without materialising the permuted tensor.  Differentiable."""
    require_cuda(feat, ids_keep)
    _check_float(feat, "feat")
    return _GatherTokensNCHW.apply(feat, ids_keep)


class _ConvVitFuse(torch.autograd.Function):
    @staticmethod
    def forward(ctx, feat1, feat2, ids, emb3):
        dt = _act_dtype(feat1, feat2, emb3)
        f1, f2 = _f32(feat1), _f32(feat2)
        e3 = None if emb3 is None else _f32(emb3)
        B, D = f1.shape[:2]
        L = f1[0, 0].numel()
        ids = ids.contiguous()
        K = ids.shape[1]
        out = torch.empty((B, K, D), dtype=torch.float32, device=f1.device)
        with torch.cuda.device(f1.device):
            rc = lib().ep_gather_sum_nchw(stream_ptr(f1.device), f1.data_ptr(), f2.data_ptr(), ptr(e3), ids.data_ptr(), B, L, K, D,
                                          out.data_ptr())
        _lib.check(rc, "ep_gather_sum_nchw")
        ctx.save_for_backward(ids)
        ctx.meta = (feat1.shape, feat1.dtype, feat2.dtype, None if emb3 is None else emb3.dtype)
        return out.to(dt)

    @staticmethod
    def backward(ctx, g):
        (ids,) = ctx.saved_tensors
        shape, d1, d2, d3 = ctx.meta
        g32 = _f32(g)
        B, K, D = g32.shape
        gf = None
        if ctx.needs_input_grad[0] or ctx.needs_input_grad[1]:
            L = 1
            for v in shape[2:]:
                L *= v
            gf = torch.empty(shape, dtype=torch.float32, device=g.device)
            with torch.cuda.device(g.device):
                rc = lib().ep_scatter_add_tokens_nchw(stream_ptr(g.device), g32.data_ptr(), ids.data_ptr(), B, L, K, D, gf.data_ptr())
            _lib.check(rc, "ep_scatter_add_tokens_nchw")
        return (gf.to(d1) if ctx.needs_input_grad[0] else None, gf.to(d2) if ctx.needs_input_grad[1] else None, None,
                g32.to(d3) if (d3 is not None and ctx.needs_input_grad[3]) else None)


def convvit_fuse_stages(feat1, feat2, ids_keep, emb_stage3=None):
    """emb_stage1 + emb_stage2 + emb_stage3 of ConvViT.forward(mask=True) (convvit.py:137-140, 151-154, 166-167) from the two
    stage decoders' raw outputs feat1, feat2 (B,D,14,14), the per-sample ids_keep (B,K) and the transformer stage's tokens
    (B,K,D): one launch instead of two permuted copies, two gathers and two adds.  Differentiable in all three."""
    require_cuda(feat1, feat2, ids_keep)
    _check_float(feat1, "feat1")
    _check_float(feat2, "feat2")
    if feat1.shape != feat2.shape:
        raise ValueError("the two stage features must have the same shape")
    return _ConvVitFuse.apply(feat1, feat2, ids_keep, emb_stage3)
