#!/usr/bin/env python
"""Generate the golden (known-answer) fixtures under tests/golden/ by EXECUTING the
unmodified reference (EventPretrain, mounted read-only at /root/reference).

The reference ships no tests or golden vectors of its own, so parity is
pinned by running its functions here on seeded synthetic inputs and committing
input + output pairs.  /root/reference does not exist on the GPU box, so nothing at test
time imports it: tests read only the .npz files this script writes.

    python tests/golden/make_golden.py            # rewrites tests/golden/*.npz

Reference entry points exercised (file:line relative to /root/reference):
  dataset/dataset_utils/events_to_voxel_grid.py:4-61    events_to_voxel_grid
  dataset/dataset_utils/events_to_image.py:6-32         events_to_image_ecdp
  dataset/dataset_utils/events_to_image.py:35-62        events_to_image_mem
  dataset/dataset_utils/events_to_image.py:65-75        remove_hot_pixel_mem
  dataset/dataset_utils/events_to_image.py:77-125       events_to_EvRep
  dataset/augmentation/events_augment.py:22-26          events_reshape
  dataset/augmentation/view_augment.py:60-63            frame_time_flip
  model/backbone/vit.py:66-130                          ViT.random_masking / forward(mask=True)
  model/backbone/convvit.py:126-171                     ConvViT.forward(mask=True)
  model/backbone/swin.py:113-179                        SwinTransformer.random_masking / apply_mask
  utils/reshape.py:15-22                                frame2emb
  model/pretrain/pr_hub_model.py:125-141                PrHubModel.reconstruct_loss
  model/pretrain/pr_rec_decoder.py:53-62                PrRecDecoder.forward (un-shuffle)
  utils/pos_embed.py:40-55                              get_2d_sincos_pos_embed
"""
import os
import sys
import types
from types import SimpleNamespace

import numpy as np
import torch
import torch.nn as nn

REF = os.environ.get("EP_REFERENCE_ROOT", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from synth import hash_uniform  # noqa: E402


def _install_stubs():
    """timm is absent from the image; the path uses it only for DropPath/to_2tuple/trunc_normal_."""
    timm = types.ModuleType("timm")
    tm = types.ModuleType("timm.models")
    tl = types.ModuleType("timm.models.layers")

    class DropPath(nn.Identity):
        def __init__(self, *a, **k):
            super().__init__()

    tl.DropPath = DropPath
    tl.to_2tuple = lambda x: x if isinstance(x, tuple) else (x, x)
    tl.trunc_normal_ = nn.init.trunc_normal_
    timm.models = tm
    tm.layers = tl
    sys.modules.update({"timm": timm, "timm.models": tm, "timm.models.layers": tl})


def _import_reference():
    if not os.path.isdir(REF):
        raise SystemExit(f"reference not found at {REF}; fixtures can only be regenerated where it is mounted")
    sys.path.insert(0, REF)
    _install_stubs()
    from dataset.dataset_utils.events_to_voxel_grid import events_to_voxel_grid
    from dataset.dataset_utils import events_to_image as eti
    from dataset.augmentation.events_augment import events_reshape
    from dataset.augmentation import view_augment
    from model.backbone import vit, convvit, swin
    from model.pretrain import pr_hub_model, pr_rec_decoder
    from utils import reshape, pos_embed
    return SimpleNamespace(voxel=events_to_voxel_grid, eti=eti, events_reshape=events_reshape,
                           view_augment=view_augment, vit=vit, convvit=convvit, swin=swin,
                           hub=pr_hub_model, dec=pr_rec_decoder, reshape=reshape, pos_embed=pos_embed)


# --------------------------------------------------------------------------------------
# synthetic streams
# --------------------------------------------------------------------------------------
def uniform_stream(rng, n, w, h, t_span=0.3, pol=(0, 1), dtype=np.float64, t0=0.0):
    x = rng.integers(0, w, n)
    y = rng.integers(0, h, n)
    t = np.sort(rng.uniform(t0, t0 + t_span, n))
    p = rng.choice(np.asarray(pol), n)
    return np.stack([x, y, t, p], 1).astype(dtype)


def stage1_cases(ref):
    rng = np.random.default_rng(1001)
    cases = {}

    def add(name, ev, size, bins, reshape=None, **extra):
        cases[name] = dict(events=ev, size=size, bins=bins, reshape=reshape, **extra)

    add("c1_small", uniform_stream(rng, 20000, 240, 180), (180, 240), 5)
    add("single_event", np.array([[3, 2, 0.5, 1]], np.float64), (8, 8), 5)
    add("two_events", np.array([[0, 0, 0.0, 1], [7, 7, 1.0, 0]], np.float64), (8, 8), 5)
    ev = uniform_stream(rng, 100, 16, 12)
    ev[:, 2] = 0.125
    add("same_timestamp", ev, (12, 16), 5)
    # events exactly on the bin nodes t0 + j*dT/(B-1)
    nodes = np.repeat(np.linspace(0.0, 1.0, 5), 40)
    ev = uniform_stream(rng, nodes.size, 16, 12)
    ev[:, 2] = nodes
    add("bin_boundary", ev, (12, 16), 5)
    add("pm1_polarity", uniform_stream(rng, 3000, 32, 24, pol=(-1, 1)), (24, 32), 5)
    add("all_positive", uniform_stream(rng, 1000, 32, 24, pol=(1,)), (24, 32), 5)
    add("all_negative", uniform_stream(rng, 1000, 32, 24, pol=(0,)), (24, 32), 5)
    # 640x480 -> 224x224 fp64 scale trap (x = 180, 340, 360 land one below the exact floor)
    ev = uniform_stream(rng, 6000, 640, 480)
    ev[:640, 0] = np.arange(640)
    ev[640:1120, 1] = np.arange(480)
    add("reshape_trap", ev, (224, 224), 5, reshape=(640, 480, 224, 224))
    ev = uniform_stream(rng, 4000, 346, 260)
    add("reshape_mvsec", ev, (224, 224), 9, reshape=(346, 260, 224, 224))
    # fractional coordinates as produced by erase_and_add_events (jitter then clip)
    ev = uniform_stream(rng, 5000, 64, 48)
    ev[:, 0] = np.clip(ev[:, 0] + rng.normal(0, 1.5, 5000), 0, 63)
    ev[:, 1] = np.clip(ev[:, 1] + rng.normal(0, 1.5, 5000), 0, 47)
    ev[:, 2] += rng.normal(0, 0.001, 5000)
    ev = ev[ev[:, 2].argsort()]
    add("fractional_xy", ev, (48, 64), 5)
    # float32 events: the reference then does its time arithmetic in fp32 (DDD17 / DVS-Gesture)
    ev = uniform_stream(rng, 5000, 64, 48, t_span=2.5e5, t0=1.0e6, dtype=np.float32)
    add("f32_events", ev, (48, 64), 5)
    # unsorted rows: first/last row are not min/max, so some events fall outside [0, B-1]
    ev = uniform_stream(rng, 4000, 64, 48)
    ev = ev[rng.permutation(4000)]
    add("unsorted_time", ev, (48, 64), 5)
    # DSEC-like: microsecond stamps, 15 bins
    ev = uniform_stream(rng, 8000, 64, 44, t_span=5.0e4, t0=5.2e10)
    ev[:, 2] = np.floor(ev[:, 2])
    add("dsec_15bins", ev, (44, 64), 15)
    add("mvsec_9bins", uniform_stream(rng, 6000, 86, 65, t_span=0.05, t0=1.5e9), (65, 86), 9)
    add("num_bins_1", uniform_stream(rng, 500, 16, 12), (12, 16), 1)
    add("num_bins_2", uniform_stream(rng, 500, 16, 12), (12, 16), 2)
    # x == W wraps into the next row (flat-index semantics of index_add_/bincount; no per-axis check)
    ev = uniform_stream(rng, 300, 16, 11)
    ev[::7, 0] = 16
    add("wrap_x", ev, (12, 16), 5)
    # hot pixel for the MEM frame + remove_hot_pixel_mem
    ev = uniform_stream(rng, 6000, 32, 24)
    ev[1000:3000, 0] = 5
    ev[1000:3000, 1] = 7
    add("hot_pixel", ev, (24, 32), 5)

    out = {}
    for name, c in cases.items():
        ev = c["events"]
        args = SimpleNamespace(num_bins=c["bins"])
        ev_in = ev.copy()
        if c["reshape"] is not None:
            sw, sh, iw, ih = c["reshape"]
            ev_in = ref.events_reshape(ev_in, sw, sh, iw, ih)
        size = c["size"]
        voxel = ref.voxel(args, ev_in.copy(), size).numpy()
        ecdp = ref.eti.events_to_image_ecdp(args, ev_in.copy(), size).numpy()
        mem = ref.eti.events_to_image_mem(args, ev_in.copy(), size)
        mem_hot = ref.eti.remove_hot_pixel_mem((mem / 255).clone()).numpy()
        # per-sample normalisers, as the datasets apply them
        # (pr_n_imagenet_dataset.py:142-143, ft_n_caltech101_dataset.py:93-98)
        e = torch.from_numpy(ecdp)
        ecdp_norm = ((e / (e.amax([1, 2], True) + 1)) - 0.5) * 2
        rec = dict(events=ev, size=np.asarray(size), bins=np.asarray(c["bins"]),
                   reshape=np.asarray(c["reshape"] if c["reshape"] else (0, 0, 0, 0)),
                   voxel=voxel, voxel_sum=torch.from_numpy(voxel).sum(dim=0)[None].numpy(),
                   ecdp=ecdp, mem=mem.numpy(), mem_hot=mem_hot, ecdp_norm=ecdp_norm.numpy())
        mh = torch.from_numpy(mem_hot)
        if mh[0::2].max() != 0:
            m2 = mh.clone()
            m2[0::2] = m2[0::2] * (1.0 / m2[0::2].max())
            rec["mem_norm"] = m2.numpy()
        # EvRep: callers cast x,y to int16 (ft_n_caltech101_dataset.py:79-80); skip fractional / wrap cases
        if name not in ("fractional_xy", "wrap_x") and c["reshape"] is None:
            h, w = size
            xs = ev[:, 0].astype(np.int16)
            ys = ev[:, 1].astype(np.int16)
            rec["evrep"] = ref.eti.events_to_EvRep(xs, ys, ev[:, 2].astype(np.float64), ev[:, 3], (w, h))
            rec["evrep_us"] = ref.eti.events_to_EvRep(xs, ys, ev[:, 2].astype(np.float64) * 1e6, ev[:, 3], (w, h))
        out[name] = rec
    return out


def stage3_cases(ref):
    out = {}
    # ---- random masking (vit.py:66-105; identical copies in convvit.py:85-124, swin.py:113-152)
    for L, ratio, seed in ((196, 0.75, 3000), (196, 0.5, 3001), (49, 0.75, 3002), (196, 0.9, 3003)):
        B = 16
        self_ns = SimpleNamespace(num_patches=L, mask_ratio=ratio, patch_size=16,
                                  args=SimpleNamespace(masking_strategy="random"))
        x = torch.zeros(B, 5, 224, 224)
        torch.manual_seed(seed)
        if L == 49:
            ids_keep, mask, ids_restore = ref.swin.SwinTransformer.random_masking(self_ns, x, x, ratio)
        else:
            ids_keep, mask, ids_restore = ref.vit.ViT.random_masking(self_ns, x)
        torch.manual_seed(seed)
        noise = torch.rand(B, L)
        out[f"mask_random_L{L}_r{int(ratio * 100)}"] = dict(
            noise=noise.numpy(), L=np.asarray(L), ratio=np.asarray(ratio), ids_keep=ids_keep.numpy(),
            mask=mask.numpy(), ids_restore=ids_restore.numpy())

    # ---- density / anti-density noise (vit.py:79-89, swin.py:126-136): tie-free float input, regenerated
    # by the tests from tests/golden/synth.py (seed 3100) instead of being stored
    x = torch.from_numpy(hash_uniform((4, 5, 224, 224), 3100))
    for strat in ("density", "anti-density"):
        for L, p in ((196, 16), (49, 32)):
            self_ns = SimpleNamespace(num_patches=L, mask_ratio=0.75, patch_size=p,
                                      args=SimpleNamespace(masking_strategy=strat))
            if L == 49:
                ik, m, ir = ref.swin.SwinTransformer.random_masking(self_ns, x, x, 0.75)
            else:
                ik, m, ir = ref.vit.ViT.random_masking(self_ns, x)
            with torch.no_grad():
                dens = nn.AvgPool2d(p, p)(abs(torch.sum(x, dim=1))).flatten(1)
            out[f"mask_{strat}_L{L}"] = dict(ids_keep=ik.numpy(), mask=m.numpy(), ids_restore=ir.numpy(),
                                             p=np.asarray(p), L=np.asarray(L), density=dens.numpy(),
                                             seed=np.asarray(3100), shape=np.asarray(x.shape))

    # ---- ViT forward(mask=True): tokens before / after the visible gather (vit.py:107-115)
    args = SimpleNamespace(masking_strategy="random", use_feature_fusion=True, phase="pretrain", pr_phase="rec")
    torch.manual_seed(3200)
    model = ref.vit.vit_small_patch16(args=args, num_bins=5, mask_ratio=0.75)
    model.eval()
    cap = {}
    model.patch_embed.register_forward_hook(lambda m, i, o: cap.__setitem__("embed", o.detach().clone()))
    model.vit_block[0].register_forward_pre_hook(lambda m, i: cap.__setitem__("gathered", i[0].detach().clone()))
    xin = torch.randn(1, 5, 224, 224)
    torch.manual_seed(3201)
    with torch.no_grad():
        _, _, _, mask, ids_restore = model(xin, mask=True)
    torch.manual_seed(3201)
    noise = torch.rand(1, 196)
    tokens = cap["embed"].flatten(2).permute(0, 2, 1).contiguous()
    out["vit_gather"] = dict(tokens=tokens.numpy(), pos_embed=model.pos_embed.detach().numpy()[0],
                             noise=noise.numpy(), mask=mask.numpy(), ids_restore=ids_restore.numpy(),
                             gathered=cap["gathered"].numpy(),
                             pos_embed_ref=ref.pos_embed.get_2d_sincos_pos_embed(384, 14).astype(np.float32))

    # ---- ConvViT forward(mask=True): the two expanded block masks (convvit.py:129-130,142-143)
    torch.manual_seed(3300)
    cmodel = ref.convvit.__dict__["convvit_small_patch16"](args=args, num_bins=5, mask_ratio=0.75)
    cmodel.eval()
    cap2 = {}
    cmodel.conv_block1[0].register_forward_pre_hook(lambda m, i: cap2.__setitem__("m1", i[1].detach().clone()))
    cmodel.conv_block2[0].register_forward_pre_hook(lambda m, i: cap2.__setitem__("m2", i[1].detach().clone()))
    cxin = torch.randn(3, 5, 224, 224)
    torch.manual_seed(3301)
    with torch.no_grad():
        _, _, _, cmask, cids_restore = cmodel(cxin, mask=True)
    torch.manual_seed(3301)
    cnoise = torch.rand(3, 196)
    out["convvit_masks"] = dict(noise=cnoise.numpy(), mask=cmask.numpy(), ids_restore=cids_restore.numpy(),
                                keep_mask_56=cap2["m1"].numpy(), keep_mask_28=cap2["m2"].numpy())

    # ---- Swin apply_mask (swin.py:154-179): batch-shared mask[:1], boolean compaction + coords
    xs = torch.from_numpy(hash_uniform((3, 3136, 8), 3400))
    torch.manual_seed(3401)
    self_ns = SimpleNamespace(num_patches=49, args=SimpleNamespace(masking_strategy="random"))
    ik, m, ir = ref.swin.SwinTransformer.random_masking(self_ns, xs, xs, 0.75)
    x_vis, coords, vis_mask = ref.swin.SwinTransformer.apply_mask(None, xs, m.bool(), (56, 56))
    out["swin_apply_mask"] = dict(seed=np.asarray(3400), shape=np.asarray(xs.shape), mask=m.numpy(),
                                  ids_keep=ik.numpy(), ids_restore=ir.numpy(), x_vis=x_vis.numpy(),
                                  coords=coords.numpy(), vis_mask=vis_mask.numpy())
    # 14x14 mask on a 56x56 grid (up_ratio 16)
    g = torch.Generator().manual_seed(3402)
    m196 = (torch.rand(2, 196, generator=g) < 0.6)
    xv2, c2, v2 = ref.swin.SwinTransformer.apply_mask(None, xs[:2], m196, (56, 56))
    out["swin_apply_mask_196"] = dict(mask=m196.numpy(), x_vis=xv2.numpy(), coords=c2.numpy(), vis_mask=v2.numpy())
    # mask already at token resolution (up_ratio 1)
    m49 = (torch.rand(2, 49, generator=g) < 0.5)
    x49 = torch.from_numpy(hash_uniform((2, 49, 8), 3403))
    xv3, c3, v3 = ref.swin.SwinTransformer.apply_mask(None, x49, m49, (7, 7))
    out["swin_apply_mask_49"] = dict(mask=m49.numpy(), x_vis=xv3.numpy(), coords=c3.numpy(), vis_mask=v3.numpy())

    # ---- target: frame2emb + norm_pix + masked MSE (utils/reshape.py:15-22, pr_hub_model.py:125-141)
    # frame / pred are regenerated by the tests from synth.hash_uniform (seeds below)
    g = torch.Generator().manual_seed(3500)
    for p, L in ((16, 196), (32, 49)):
        frame = torch.from_numpy(hash_uniform((4, 1, 224, 224), 3500 + p)).clone()
        frame[1] *= 1e-4          # near-constant patches exercise the eps
        frame[2, :, :p, :p] = 0.25  # exactly constant patch: var == 0
        pred = torch.from_numpy(hash_uniform((4, L, p * p), 3600 + p)) * 4
        maskf = (torch.rand(4, L, generator=g) < 0.75).float()
        emb = ref.reshape.frame2emb(p, frame)
        rec = dict(mask=maskf.numpy(), emb0=emb[:1].numpy(), p=np.asarray(p), L=np.asarray(L))
        for norm in (True, False):
            ns = SimpleNamespace(patch_size=p, norm_pix_loss=norm, mask_ratio=0.75)
            rec[f"loss_norm{int(norm)}"] = ref.hub.PrHubModel.reconstruct_loss(ns, pred, frame, maskf).numpy()
        ns0 = SimpleNamespace(patch_size=p, norm_pix_loss=True, mask_ratio=0)
        rec["loss_nomask"] = ref.hub.PrHubModel.reconstruct_loss(ns0, pred, frame, maskf).numpy()
        # per-patch loss with a one-hot mask pins the normalised target patch by patch
        per_patch = torch.zeros(4, L)
        ns1 = SimpleNamespace(patch_size=p, norm_pix_loss=True, mask_ratio=0.75)
        for b in range(4):
            for l in range(0, L, 7):
                oh = torch.zeros(4, L)
                oh[b, l] = 1
                per_patch[b, l] = ref.hub.PrHubModel.reconstruct_loss(ns1, pred, frame, oh)
        rec["per_patch_loss_stride7"] = per_patch.numpy()
        out[f"target_p{p}"] = rec
    # multi-channel patchify order (p,q,c)
    frame = torch.from_numpy(hash_uniform((2, 3, 64, 64), 3700))
    out["frame2emb_c3"] = dict(emb=ref.reshape.frame2emb(8, frame).numpy(), p=np.asarray(8))

    # ---- decoder un-shuffle (pr_rec_decoder.py:53-62): mask-token fill + gather(ids_restore) + pos_embed
    torch.manual_seed(3600)
    dec = ref.dec.pretrain_rec_decoder_small_patch16()
    dec.eval()
    nn.init.normal_(dec.mask_token, std=0.02)
    cap3 = {}
    dec.patch_embed.register_forward_hook(lambda m, i, o: cap3.__setitem__("emb", o.detach().clone()))
    dec.vit_block[0].register_forward_pre_hook(lambda m, i: cap3.__setitem__("x", i[0].detach().clone()))
    emb_lh = torch.randn(2, 49, 384)
    ids_r = torch.stack([torch.randperm(196) for _ in range(2)])
    with torch.no_grad():
        dec(emb_lh, ids_r)
    out["decoder_unshuffle"] = dict(emb=cap3["emb"].numpy(), ids_restore=ids_r.numpy(), mask_token=dec.mask_token.detach().numpy()[0, 0],
                                    pos_embed=dec.pos_embed.detach().numpy()[0], x=cap3["x"].numpy())

    # ---- sign convention of the diff target under time reversal (view_augment.py:60-63)
    f = torch.from_numpy(hash_uniform((1, 8, 8), 3800))
    out["frame_time_flip"] = dict(frame=f.numpy(), flipped=ref.view_augment.frame_time_flip(f.clone()).numpy())
    return out


def view_cases(ref):
    """evg_augment / frame_augment (dataset/augmentation/view_augment.py:65-89) with seeds: crop box, resize mode,
    horizontal flip and time flip all come from the reference's own RNG sequence."""
    out = {}
    va = ref.view_augment
    for bins, mode, seeds in ((5, "nearest", (1, 2, 3, 4, 5, 6)), (5, "bilinear", (7, 8, 9)), (15, "bilinear", (10, 11)),
                              (2, "bicubic", (12,))):
        args = SimpleNamespace(crop_min=0.2, num_bins=bins, input_size=32)
        grid = torch.from_numpy(hash_uniform((bins, 60, 80), 4000 + bins)) * 8
        for sd in seeds:
            o, flag = va.evg_augment(args, grid.clone(), (32, 32), mode=mode, seed=sd)
            out[f"evg_{mode}_b{bins}_s{sd}"] = dict(out=o.numpy(), flag=np.asarray(flag), bins=np.asarray(bins), seed=np.asarray(sd),
                                                  mode=np.asarray(["nearest", "bilinear", "bicubic"].index(mode)))
    args = SimpleNamespace(crop_min=0.2, num_bins=5, input_size=32)
    frame = torch.from_numpy(hash_uniform((1, 60, 80), 4100))
    for sd, tf in ((21, False), (22, True), (23, True), (24, False)):
        o = va.frame_augment(args, frame.clone(), seed=sd, time_flip_flag=tf)
        out[f"frame_s{sd}"] = dict(out=o.numpy(), seed=np.asarray(sd), tflip=np.asarray(tf))
    # a grid and its sub_frame augmented with the same seed share the crop (pr_ef_imagenet_dataset.py:187-206)
    # no-crop fallback: crop_min so large that 10 attempts fail for a tiny frame
    args = SimpleNamespace(crop_min=0.999, num_bins=5, input_size=8)
    tiny = torch.from_numpy(hash_uniform((5, 3, 3), 4200))
    o, flag = va.evg_augment(args, tiny.clone(), (8, 8), mode="nearest", seed=31)
    out["evg_nocrop"] = dict(out=o.numpy(), flag=np.asarray(flag))
    return out


def stream_aug_cases(ref):
    """events_augment / add_noise_events (dataset/augmentation/events_augment.py:28-86) with seeds, then binned by the
    reference: pins the RNG order of the host-side mirror and the binning of the fractional coordinates it produces."""
    import importlib
    ea = importlib.import_module("dataset.augmentation.events_augment")
    rng = np.random.default_rng(6001)
    ev = uniform_stream(rng, 5000, 64, 48)
    args = SimpleNamespace(num_bins=5)
    out = {}
    for sd in (1, 2):
        aug = ea.events_augment(args, ev.copy(), size=(48, 64), seed=sd)
        out[f"erase_add_s{sd}"] = dict(events=ev, seed=np.asarray(sd), aug=aug, voxel=ref.voxel(args, aug.copy(), (48, 64)).numpy(),
                                       ecdp=ref.eti.events_to_image_ecdp(args, aug.copy(), (48, 64)).numpy())
    np.random.seed(3)
    noisy = ea.add_noise_events(args, ev.copy(), (48, 64))
    out["add_noise_s3"] = dict(events=ev, seed=np.asarray(3), aug=noisy)
    tiny = uniform_stream(rng, 50, 16, 12)
    out["erase_add_tiny"] = dict(events=tiny, seed=np.asarray(4), aug=ea.events_augment(args, tiny.copy(), size=(12, 16), seed=4))
    return out


def swin_grouping_cases(ref):
    """model/sub_module/swin_block.py:280-464 and :196-203 — knapsack / group_windows on random window fills, full
    GroupingModule plans and PatchMerging's token order for masks of the Swin front-end.  The reference sorts tokens by
    window id with torch.argsort's default; it is run here with stable=True (the order CUDA's radix sort gives on the
    device the models run on), which fixes the order of tokens inside a window."""
    from model.sub_module import swin_block as sb
    out = {}
    rng = np.random.default_rng(5100)
    for c in range(24):
        n, gs = int(rng.integers(1, 70)), int(rng.integers(1, 50))
        wt = rng.integers(1, gs + 1, n).astype(np.int64)
        fills, groups = sb.group_windows(gs, [int(v) for v in wt])
        out[f"gw{c:02d}"] = dict(group_size=np.asarray(gs), wt=wt, fills=np.asarray(fills, np.int64),
                                 first=np.cumsum([0] + [len(g) for g in groups]).astype(np.int64),
                                 idx=np.asarray([i for g in groups for i in g], np.int64))
    orig = torch.argsort
    torch.argsort = lambda x, *a, **k: orig(x, *a, **{**k, "stable": True})
    try:
        gen = torch.Generator().manual_seed(5200)
        for name, res, keep, ws, shift in (("g56_24_s0", 56, 24, 7, 0), ("g56_24_s3", 56, 24, 7, 3), ("g28_24_s3", 28, 24, 7, 3),
                                           ("g56_12_s0", 56, 12, 7, 0), ("m14_24_s0", 14, 24, 7, 0)):
            m = torch.zeros(49)
            m[torch.randperm(49, generator=gen)[:49 - keep]] = 1
            up = res // 7
            mm = m.reshape(7, 7)[:, None, :, None].expand(7, up, 7, up).reshape(-1).bool()
            ii, jj = torch.meshgrid(torch.arange(res), torch.arange(res), indexing="ij")
            coords = torch.stack([ii, jj], -1).reshape(1, -1, 2)[:, ~mm]
            gm = sb.GroupingModule(ws, shift)
            attn_mask, rel_pos_idx = gm.prepare(coords.clone(), coords.shape[1])
            rec = dict(res=np.asarray(res), window=np.asarray(ws), shift=np.asarray(shift), coords=coords.numpy(),
                       grouping=np.asarray(gm._mode == "grouping"), attn_mask=attn_mask.numpy(), rel_pos_idx=rel_pos_idx.numpy())
            if gm._mode == "grouping":
                rec.update(idx_shuffle=gm.idx_shuffle.numpy(), idx_unshuffle=gm.idx_unshuffle.numpy(), group_size=np.asarray(gm.group_size))
            # PatchMerging's order (swin_block.py:196-203) for the same visibility mask
            vis = (~mm).reshape(1, -1)
            mask = vis.reshape(res // 2, 2, res // 2, 2).permute(0, 2, 1, 3).reshape(-1)
            cg = sb.get_coordinates(res, res).reshape(2, -1).permute(1, 0)
            cg = cg.reshape(res // 2, 2, res // 2, 2, 2).permute(0, 2, 1, 3, 4).reshape(-1, 2)
            cl = cg[mask].reshape(-1, 2)
            cl = cl[:, 0] * res + cl[:, 1]
            rec.update(vis=vis.numpy(), merge_order=torch.argsort(torch.argsort(cl)).numpy())
            out[name] = rec
    finally:
        torch.argsort = orig
    return out


def main():
    ref = _import_reference()
    torch.set_num_threads(1)
    s1 = stage1_cases(ref)
    flat = {f"{case}/{k}": v for case, rec in s1.items() for k, v in rec.items()}
    np.savez_compressed(os.path.join(HERE, "stage1_events.npz"), **flat)
    s3 = stage3_cases(ref)
    flat = {f"{case}/{k}": v for case, rec in s3.items() for k, v in rec.items()}
    np.savez_compressed(os.path.join(HERE, "stage3_mask_patch.npz"), **flat)
    vc = view_cases(ref)
    flat = {f"{case}/{k}": v for case, rec in vc.items() for k, v in rec.items()}
    np.savez_compressed(os.path.join(HERE, "views.npz"), **flat)
    sa = stream_aug_cases(ref)
    flat = {f"{case}/{k}": v for case, rec in sa.items() for k, v in rec.items()}
    np.savez_compressed(os.path.join(HERE, "stream_aug.npz"), **flat)
    sg = swin_grouping_cases(ref)
    flat = {f"{case}/{k}": v for case, rec in sg.items() for k, v in rec.items()}
    np.savez_compressed(os.path.join(HERE, "swin_grouping.npz"), **flat)
    for f in ("stage1_events.npz", "stage3_mask_patch.npz", "views.npz", "stream_aug.npz", "swin_grouping.npz"):
        print(f, os.path.getsize(os.path.join(HERE, f)) // 1024, "KiB")
    print("torch", torch.__version__, "numpy", np.__version__)


if __name__ == "__main__":
    main()
