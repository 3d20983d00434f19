"""The token permutations carry a gradient (the reference trains through them under autocast, pr_trainer.py:26-36):
forward values against the torch ops of the reference's forward, gradients against torch.autograd of those same ops,
float16 / bfloat16 activations accepted.  Swin stage-output consumers (swin.py:206-238) against golden vectors produced by
executing the reference's statements (tests/golden/make_golden_swin_consumers.py)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ep(native_lib):
    import eventpretrain_b200 as ep
    return ep


@pytest.fixture(autouse=True)
def exact_fp32_convolutions():
    """The stage decoders are nn.Conv2d: keep cuDNN off TF32 so the comparison with the reference's CPU run is about the
    permutation kernels, not about the convolution's precision mode."""
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32 = old


def cu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def leaf(*shape, dtype=torch.float32, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return torch.randn(*shape, device="cuda", generator=g).to(dtype).requires_grad_(True)


def grads(out, inputs, seed=99):
    g = torch.Generator(device="cuda").manual_seed(seed)
    w = torch.randn(out.shape, device="cuda", generator=g).to(out.dtype)
    return torch.autograd.grad((out * w).sum(), inputs)


def test_gather_tokens_grad_matches_torch(ep):
    B, L, K, D = 3, 196, 49, 64
    tokens, pos = leaf(B, L, D, seed=1), leaf(1, L, D, seed=2)
    ids = torch.argsort(torch.rand(B, L, device="cuda"), dim=1)[:, :K]
    out = ep.gather_tokens(tokens, ids, pos)
    ref = torch.gather(tokens + pos, 1, ids.unsqueeze(-1).repeat(1, 1, D))                 # vit.py:113-115
    assert torch.equal(out, ref) and out.grad_fn is not None
    g_t, g_p = grads(out, (tokens, pos))
    r_t, r_p = grads(ref, (tokens, pos))
    assert torch.equal(g_t, r_t)
    torch.testing.assert_close(g_p, r_p, rtol=1e-6, atol=1e-6)
    # without pos_embed, and with only the tokens needing a gradient
    out2 = ep.gather_tokens(tokens, ids)
    assert torch.equal(grads(out2, (tokens,))[0], grads(torch.gather(tokens, 1, ids.unsqueeze(-1).repeat(1, 1, D)), (tokens,))[0])


def test_gather_tokens_shared_ids_with_duplicates(ep):
    """GroupingModule.group(): batch-shared index_select whose padded slots repeat token 0 (swin_block.py:416,452-457)."""
    B, N, D = 2, 100, 32
    x = leaf(B, N, D, seed=3)
    idx = torch.cat([torch.randperm(N, device="cuda"), torch.zeros(28, dtype=torch.int64, device="cuda")])
    out = ep.gather_tokens(x, idx)
    ref = torch.index_select(x, 1, idx)
    assert torch.equal(out, ref)
    # token 0 receives 29 gradient rows: fp32 atomics in either implementation, the order differs
    torch.testing.assert_close(grads(out, (x,))[0], grads(ref, (x,))[0], rtol=1e-5, atol=1e-5)


def test_grouping_module_carries_gradient(ep, golden_swin_grouping):
    c = golden_swin_grouping["g56_24_s3"]
    coords = cu(c["coords"])
    gm = ep.GroupingModule(int(c["window"]), int(c["shift"]))
    gm.prepare(coords, coords.shape[1])
    x = leaf(2, coords.shape[1], 96, seed=4)
    y = gm.merge(gm.group(x) * 2.0)
    assert torch.equal(y, x * 2.0)
    (gx,) = grads(y, (x,))
    (rx,) = grads(x * 2.0, (x,))
    torch.testing.assert_close(gx, rx, rtol=1e-6, atol=1e-6)


def test_unshuffle_tokens_grad_matches_torch(ep):
    B, L, K, D = 3, 196, 49, 128
    emb, mt, pos = leaf(B, K, D, seed=5), leaf(1, 1, D, seed=6), leaf(1, L, D, seed=7)
    ids_restore = torch.argsort(torch.argsort(torch.rand(B, L, device="cuda"), dim=1), dim=1)
    out = ep.unshuffle_tokens(emb, mt, ids_restore, pos)
    cat = torch.cat([emb, mt.repeat(B, L - K, 1)], dim=1)                                      # pr_rec_decoder.py:56-62
    ref = torch.gather(cat, 1, ids_restore.unsqueeze(-1).repeat(1, 1, D)) + pos
    assert torch.equal(out, ref)
    got, want = grads(out, (emb, mt, pos)), grads(ref, (emb, mt, pos))
    assert torch.equal(got[0], want[0])
    torch.testing.assert_close(got[1], want[1], rtol=1e-5, atol=1e-5)        # a sum over B * (L - K) rows: order differs
    torch.testing.assert_close(got[2], want[2], rtol=1e-6, atol=1e-6)


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
def test_half_activations(ep, dtype):
    """Under autocast the activations are fp16 (pr_trainer.py:26): the drop-ins take them, answer in the same dtype and
    route the gradient back in it."""
    B, L, K, D = 2, 196, 49, 64
    tokens = leaf(B, L, D, dtype=dtype, seed=8)
    ids = torch.argsort(torch.rand(B, L, device="cuda"), dim=1)[:, :K]
    out = ep.gather_tokens(tokens, ids)
    assert out.dtype == dtype and torch.equal(out, torch.gather(tokens, 1, ids.unsqueeze(-1).repeat(1, 1, D)))
    (g,) = grads(out, (tokens,))
    assert g.dtype == dtype and torch.equal(g, grads(torch.gather(tokens, 1, ids.unsqueeze(-1).repeat(1, 1, D)), (tokens,))[0])
    emb, mt = leaf(B, K, D, dtype=dtype, seed=9), leaf(1, 1, D, seed=10)
    ids_restore = torch.argsort(torch.argsort(torch.rand(B, L, device="cuda"), dim=1), dim=1)
    o2 = ep.unshuffle_tokens(emb, mt, ids_restore)
    assert o2.dtype == dtype
    ge, gm = grads(o2, (emb, mt))
    assert ge.dtype == dtype and gm.dtype == torch.float32 and torch.isfinite(gm).all()
    with torch.autocast("cuda", dtype=dtype):
        x = leaf(B, 3136, 96, seed=11)
        h = (x * 1.0).to(dtype)
        mask = torch.zeros(B, 49, device="cuda")
        mask[:, ::4] = 1
        xv, co, vm = ep.swin_apply_mask(h, mask.bool(), (56, 56))
        assert xv.dtype == dtype
        (gx,) = torch.autograd.grad(xv.float().sum(), (x,))
        assert torch.equal(gx.reshape(B, 56, 56, 96)[0, :, :, 0] != 0, vm.reshape(56, 56))


def test_swin_apply_mask_grad_and_n_vis_contract(ep, golden_stage3):
    c = golden_stage3["swin_apply_mask"]
    B, N, C = (int(v) for v in c["shape"])
    x = leaf(B, N, C, seed=12)
    mask = cu(c["mask"])
    xv, co, vm = ep.swin_apply_mask(x, mask, (56, 56))
    ref = x[:, vm[0]]                                                         # swin.py:171-176: boolean compaction, batch-shared
    assert torch.equal(xv, ref)
    ref_grad = grads(ref, (x,))[0]
    assert torch.equal(grads(xv, (x,))[0], ref_grad)
    n_vis = int(vm.sum())
    xv2, co2, _ = ep.swin_apply_mask(x, mask, (56, 56), n_vis=n_vis, check=True)
    assert torch.equal(xv2, ref) and torch.equal(co2, co)
    assert torch.equal(grads(xv2, (x,))[0], ref_grad)
    with pytest.raises(ValueError):
        ep.swin_apply_mask(x, mask, (56, 56), n_vis=n_vis - 16, check=True)
    xv3, co3, _ = ep.swin_apply_mask(x, mask, (56, 56), n_vis=n_vis + 16)     # too large: packed rows, then zeros, never garbage
    flat = xv3.reshape(-1)
    assert torch.equal(flat[: B * n_vis * C].reshape(B, n_vis, C), ref) and not flat[B * n_vis * C:].any() and not co3[:, n_vis:].any()


@pytest.mark.parametrize("stage", [1, 2, 3])
def test_swin_stage_consumers_vs_reference(ep, golden_swin_consumers, stage):
    """swin.py:212-238: tokens back onto the dense grid, stage decoder, gather by the per-sample ids_keep."""
    c = golden_swin_consumers[f"s{stage}"]
    x, coords, ids_keep, G = cu(c["x"]), cu(c["coords"]), cu(c["ids_keep"]), int(c["grid"])
    dense = ep.swin_scatter_dense(x, coords, G)
    assert np.array_equal(dense.cpu().numpy(), c["dense"])
    feat = F.conv2d(dense, cu(c["weight"]), cu(c["bias"]), stride=c["weight"].shape[-1])       # the reference's nn.Conv2d
    emb = ep.gather_tokens_nchw(feat, ids_keep)
    want = torch.gather(feat.flatten(2).permute(0, 2, 1), 1, ids_keep.unsqueeze(-1).repeat(1, 1, feat.shape[1]))
    assert torch.equal(emb, want)
    np.testing.assert_allclose(emb.cpu().numpy(), c["emb_stage"], rtol=1e-4, atol=1e-5)       # cuDNN vs CPU convolution


def test_swin_stage_consumers_grad(ep, golden_swin_consumers):
    c = golden_swin_consumers["s1"]
    coords, ids_keep, G = cu(c["coords"]), cu(c["ids_keep"]), int(c["grid"])
    x = leaf(*c["x"].shape, seed=13)
    w = cu(c["weight"]).requires_grad_(True)

    def ours():
        feat = F.conv2d(ep.swin_scatter_dense(x, coords, G), w, stride=8)
        return ep.gather_tokens_nchw(feat, ids_keep)

    def theirs():
        e = torch.zeros((x.shape[0], G * G, x.shape[-1]), device="cuda").index_copy(1, coords[0, :, 0] * G + coords[0, :, 1], x)
        e = e.reshape(x.shape[0], G, G, -1).permute(0, 3, 1, 2)
        feat = F.conv2d(e, w, stride=8).flatten(2).permute(0, 2, 1)
        return torch.gather(feat, 1, ids_keep.unsqueeze(-1).repeat(1, 1, feat.shape[-1]))

    a, b = ours(), theirs()
    torch.testing.assert_close(a, b, rtol=1e-5, atol=1e-6)
    ga, gb = grads(a, (x, w)), grads(b, (x, w))
    torch.testing.assert_close(ga[0], gb[0], rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(ga[1], gb[1], rtol=1e-4, atol=1e-5)


def test_convvit_fused_stage_sum(ep):
    """convvit.py:137-140, 151-154, 166-167: emb_stage1 + emb_stage2 + emb_stage3 from the raw stage features."""
    B, D, K = 3, 384, 49
    f1, f2, e3 = leaf(B, D, 14, 14, seed=21), leaf(B, D, 14, 14, seed=22), leaf(B, K, D, seed=23)
    ids = torch.argsort(torch.rand(B, 196, device="cuda"), dim=1)[:, :K]

    def ref_stage(f):
        e = f.flatten(2).permute(0, 2, 1)
        return torch.gather(e, dim=1, index=ids.unsqueeze(-1).repeat(1, 1, e.shape[-1]))

    got = ep.convvit_fuse_stages(f1, f2, ids, e3)
    want = ref_stage(f1) + ref_stage(f2) + e3
    assert torch.equal(got, want)
    for a, b in zip(grads(got, (f1, f2, e3)), grads(want, (f1, f2, e3))):
        assert torch.equal(a, b)
    assert torch.equal(ep.convvit_fuse_stages(f1, f2, ids), ref_stage(f1) + ref_stage(f2))
