"""Stages 2-3 on the GPU (through the C ABI) against the reference's golden vectors and the numpy oracle.
Index / mask / gather / patchify outputs are bit-exact; the normalised target is within 1e-5 relative."""
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from synth import hash_uniform
from test_oracle_golden import target_inputs

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ep(native_lib):
    import eventpretrain_b200 as ep
    assert torch.cuda.is_available()
    return ep


def cu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.mark.parametrize("name", ["mask_random_L196_r75", "mask_random_L196_r50", "mask_random_L49_r75", "mask_random_L196_r90"])
def test_mask_from_noise(ep, golden_stage3, name):
    c = golden_stage3[name]
    keep = ep.len_keep_of(int(c["L"]), float(c["ratio"]))
    ik, m, ir = ep.mask_from_noise(cu(c["noise"]), keep)
    assert ik.dtype == torch.int64 and ir.dtype == torch.int64 and m.dtype == torch.float32
    assert np.array_equal(ik.cpu().numpy(), c["ids_keep"])
    assert np.array_equal(m.cpu().numpy(), c["mask"])
    assert np.array_equal(ir.cpu().numpy(), c["ids_restore"])


def test_mask_ties_are_stable(ep):
    from oracle import stage3_np as s3
    noise = np.zeros((3, 196), np.float32)
    noise[1, ::3] = 1.0
    noise[2] = np.repeat(np.arange(49, dtype=np.float32), 4)[::-1]
    ik, m, ir = ep.mask_from_noise(cu(noise), 49)
    ok, om, orr = s3.mask_from_noise(noise, 49)
    assert np.array_equal(ik.cpu().numpy(), ok) and np.array_equal(m.cpu().numpy(), om) and np.array_equal(ir.cpu().numpy(), orr)


def test_random_noise_rows_with_ties_are_detected(ep):
    """torch.rand fp32 has 2^24 values, so about one row in a thousand of a (B,196) draw holds a tie, and bit-exact
    mask parity is only well defined on tie-free rows.  The harness finds the tied rows of a real draw: on those the contract is
    torch.argsort(stable=True); on tie-free rows any argsort (the reference calls the default, unstable one) gives the same."""
    torch.manual_seed(3000)
    noise = torch.rand(16384, 196, device="cuda")
    srt = torch.sort(noise, dim=1).values
    tied = (srt[:, 1:] == srt[:, :-1]).any(dim=1)
    n_tied = int(tied.sum())
    assert 1 <= n_tied < 200, n_tied                       # expectation ~ 16384 * 196^2 / 2 / 2^24 = 19
    ik, m, ir = ep.mask_from_noise(noise, 49)
    stable = torch.argsort(noise, dim=1, stable=True)
    assert torch.equal(ik, stable[:, :49]) and torch.equal(ir, torch.argsort(stable, dim=1, stable=True))
    assert torch.equal(m, (ir >= 49).float())
    plain = torch.argsort(noise, dim=1)                    # what the reference executes (vit.py:92)
    assert torch.equal(ik[~tied], plain[~tied][:, :49])
    # a tied row whose tie straddles nothing still orders the tied pair by index
    r = int(torch.nonzero(tied)[0])
    row = noise[r]
    order = stable[r]
    eq = torch.nonzero(row[order][1:] == row[order][:-1]).flatten()
    assert all(int(order[i]) < int(order[i + 1]) for i in eq.tolist())


def test_random_masking_dropin_follows_torch_rng(ep, golden_stage3):
    """random_masking(self, x) draws torch.rand(B, L, device=x.device) itself, like the reference."""
    self_ns = SimpleNamespace(num_patches=196, mask_ratio=0.75, patch_size=16, args=SimpleNamespace(masking_strategy="random"))
    x = torch.zeros(4, 5, 224, 224, device="cuda")
    torch.manual_seed(11)
    ik, m, ir = ep.random_masking(self_ns, x)
    torch.manual_seed(11)
    noise = torch.rand(4, 196, device="cuda")
    ref = torch.argsort(noise, dim=1, stable=True)
    assert torch.equal(ik, ref[:, :49]) and torch.equal(ir, torch.argsort(ref, dim=1, stable=True))
    assert torch.equal(m, (ir >= 49).float())


@pytest.mark.parametrize("strategy", ["density", "anti-density"])
@pytest.mark.parametrize("L,p", [(196, 16), (49, 32)])
def test_density_masking(ep, golden_stage3, strategy, L, p):
    c = golden_stage3[f"mask_{strategy}_L{L}"]
    x = cu(hash_uniform(tuple(c["shape"]), int(c["seed"])))
    d = ep.patch_density(x, p)
    assert np.array_equal(d.cpu().numpy(), c["density"])
    self_ns = SimpleNamespace(num_patches=L, mask_ratio=0.75, patch_size=p, args=SimpleNamespace(masking_strategy=strategy))
    ik, m, ir = ep.random_masking(self_ns, x, x, 0.75) if L == 49 else ep.random_masking(self_ns, x)
    assert np.array_equal(ik.cpu().numpy(), c["ids_keep"]) and np.array_equal(m.cpu().numpy(), c["mask"])
    assert np.array_equal(ir.cpu().numpy(), c["ids_restore"])


def test_vit_gather(ep, golden_stage3):
    c = golden_stage3["vit_gather"]
    ik, m, ir = ep.mask_from_noise(cu(c["noise"]), 49)
    assert np.array_equal(m.cpu().numpy(), c["mask"]) and np.array_equal(ir.cpu().numpy(), c["ids_restore"])
    g = ep.gather_tokens(cu(c["tokens"]), ik, cu(c["pos_embed"])[None])
    assert np.array_equal(g.cpu().numpy(), c["gathered"])
    g0 = ep.gather_tokens(cu(c["tokens"]), ik)
    assert torch.equal(g0, torch.gather(cu(c["tokens"]), 1, ik[..., None].repeat(1, 1, 384)))


def test_patchify_gather_commutes_with_patch_embed(ep):
    """Gathering raw patches in Conv2d operand order then applying the conv weight as a matmul equals
    PatchEmbed's conv on the full image followed by the token gather (vit.py:110-115)."""
    torch.manual_seed(0)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    x = torch.randn(3, 5, 224, 224, device="cuda")
    conv = torch.nn.Conv2d(5, 32, 16, 16).cuda()
    ik, _, _ = ep.mask_from_noise(torch.rand(3, 196, device="cuda"), 49)
    patches = ep.patchify_gather(x, 16, ik, "cpq")
    assert tuple(patches.shape) == (3, 49, 5 * 256)
    with torch.no_grad():
        full = conv(x).flatten(2).permute(0, 2, 1)
        ref = torch.gather(full, 1, ik[..., None].repeat(1, 1, 32))
        got = patches @ conv.weight.reshape(32, -1).T + conv.bias
    assert torch.allclose(got, ref, atol=1e-3, rtol=1e-3)
    unf = torch.nn.functional.unfold(x, 16, stride=16).permute(0, 2, 1)      # (B, L, C*p*p) in (c,ph,pw) order
    assert torch.equal(patches, torch.gather(unf, 1, ik[..., None].repeat(1, 1, unf.shape[-1])))


def test_patchify_gather_tma_form_is_identical():
    """EP_PATCH_TMA=1 routes the (c,ph,pw) gather through the TMA box-load / bulk-store kernel (csrc/ep_patch_tma.cu);
    the switch is read once per process, so the check runs in a child process."""
    import os
    import subprocess
    import sys
    code = (
        "import torch, eventpretrain_b200 as ep\n"
        "torch.manual_seed(1)\n"
        "for (B, C, H, W, p, K) in [(3, 5, 224, 224, 16, 49), (2, 1, 64, 96, 8, 10), (5, 3, 32, 32, 4, 64), (130, 2, 64, 64, 16, 5)]:\n"
        "    x = torch.randn(B, C, H, W, device='cuda'); L = (H // p) * (W // p)\n"
        "    ids = torch.stack([torch.randperm(L, device='cuda')[:K] for _ in range(B)]) if K < L else None\n"
        "    unf = torch.nn.functional.unfold(x, p, stride=p).permute(0, 2, 1)\n"
        "    ref = unf if ids is None else torch.gather(unf, 1, ids[..., None].repeat(1, 1, unf.shape[-1]))\n"
        "    assert torch.equal(ep.patchify_gather(x, p, ids, 'cpq'), ref), (B, C, H, W, p, K)\n"
        "print('ok')\n")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, EP_PATCH_TMA="1", PYTHONPATH=root)
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "ok" in r.stdout, r.stderr[-2000:]


def test_convvit_masks(ep, golden_stage3):
    c = golden_stage3["convvit_masks"]
    _, m, _ = ep.mask_from_noise(cu(c["noise"]), 49)
    k56, k28 = ep.convvit_keep_masks(m)
    assert np.array_equal(k56.cpu().numpy(), c["keep_mask_56"]) and np.array_equal(k28.cpu().numpy(), c["keep_mask_28"])


def test_swin_apply_mask(ep, golden_stage3):
    c = golden_stage3["swin_apply_mask"]
    x = cu(hash_uniform(tuple(c["shape"]), int(c["seed"])))
    for n_vis in (None, c["x_vis"].shape[1]):
        xv, co, vm = ep.swin_apply_mask(x, cu(c["mask"]).bool(), (56, 56), n_vis=n_vis)
        assert np.array_equal(xv.cpu().numpy(), c["x_vis"]) and np.array_equal(co.cpu().numpy(), c["coords"])
        assert np.array_equal(vm.cpu().numpy(), c["vis_mask"]) and co.dtype == torch.int64 and vm.dtype == torch.bool
    c2 = golden_stage3["swin_apply_mask_196"]
    xv, co, vm = ep.swin_apply_mask(x[:2].contiguous(), cu(c2["mask"]), (56, 56))
    assert np.array_equal(xv.cpu().numpy(), c2["x_vis"]) and np.array_equal(co.cpu().numpy(), c2["coords"])
    c3 = golden_stage3["swin_apply_mask_49"]
    xv, co, vm = ep.swin_apply_mask(cu(hash_uniform((2, 49, 8), 3403)), cu(c3["mask"]), (7, 7))
    assert np.array_equal(xv.cpu().numpy(), c3["x_vis"]) and np.array_equal(co.cpu().numpy(), c3["coords"])
    assert np.array_equal(vm.cpu().numpy(), c3["vis_mask"])


@pytest.mark.parametrize("p,L", [(16, 196), (32, 49)])
def test_target(ep, golden_stage3, p, L):
    from oracle import stage3_np as s3
    c = golden_stage3[f"target_p{p}"]
    frame, pred = target_inputs(p, L)
    f, pr, mk = cu(frame), cu(pred), cu(c["mask"])
    assert np.array_equal(ep.frame2emb(p, f)[:1].cpu().numpy(), c["emb0"])
    for norm in (True, False):
        ns = SimpleNamespace(patch_size=p, norm_pix_loss=norm, mask_ratio=0.75)
        loss = ep.reconstruct_loss(ns, pr, f, mk)
        np.testing.assert_allclose(loss.item(), c[f"loss_norm{int(norm)}"], rtol=1e-5)
        t = ep.target_normpix(f, p, norm).cpu().numpy()
        o = s3.target_normpix(frame, p, norm)
        assert np.all(np.abs(t - o) <= 1e-5 * np.abs(o) + 1e-6)
        pl = ep.target_patch_loss(pr, f, p, norm).cpu().numpy()
        np.testing.assert_allclose(pl, ((pred.astype(np.float64) - o) ** 2).mean(-1), rtol=1e-5)
        plm = ep.target_patch_loss(pr, f, p, norm, mask=mk).cpu().numpy()                    # masked-only: pr_hub_model.py:139
        assert np.array_equal(plm, np.where(c["mask"] != 0, pl, 0.0))
    ns0 = SimpleNamespace(patch_size=p, norm_pix_loss=True, mask_ratio=0)
    np.testing.assert_allclose(ep.reconstruct_loss(ns0, pr, f, mk).item(), c["loss_nomask"], rtol=1e-5)
    ref = c["per_patch_loss_stride7"]
    sel = ref != 0
    np.testing.assert_allclose(ep.target_patch_loss(pr, f, p, True).cpu().numpy()[sel], ref[sel], rtol=1e-5)
    masked = ep.target_patch_loss(pr, f, p, True, mask=mk)
    np.testing.assert_allclose(((mk * masked).sum() / mk.sum()).item(), c["loss_norm1"], rtol=1e-5)


@pytest.mark.parametrize("C,p,H,W", [(1, 8, 64, 96), (1, 16, 224, 224), (1, 32, 224, 224), (1, 4, 32, 48), (3, 8, 64, 64), (1, 16, 64, 80), (2, 16, 32, 32)])
def test_target_vector_and_staged_kernels_agree_with_torch(ep, C, p, H, W):
    """Single-channel p in {8,16,32} takes the register/float4 kernel, everything else the staged one; both against the
    reference formulation (utils/reshape.py:15-22, pr_hub_model.py:126-131) in torch."""
    torch.manual_seed(7)
    frame = torch.randn(5, C, H, W, device="cuda") * 3 + 1
    gh, gw = H // p, W // p
    emb = torch.einsum("bchpwq->bhwpqc", frame.reshape(5, C, gh, p, gw, p)).reshape(5, gh * gw, p * p * C)
    assert torch.equal(ep.target_normpix(frame, p, False), emb)
    ref = (emb - emb.mean(-1, keepdim=True)) / (emb.var(-1, keepdim=True) + 1e-6) ** .5
    got = ep.target_normpix(frame, p, True)
    assert torch.all((got - ref).abs() <= 1e-5 * ref.abs() + 1e-6)
    pred = torch.randn_like(ref)
    want = ((pred.double() - ((emb.double() - emb.double().mean(-1, keepdim=True)) / (emb.double().var(-1, keepdim=True) + 1e-6) ** .5)) ** 2).mean(-1)
    torch.testing.assert_close(ep.target_patch_loss(pred, frame, p, True).double(), want, rtol=1e-5, atol=0)


def test_frame2emb_multichannel(ep, golden_stage3):
    c = golden_stage3["frame2emb_c3"]
    assert np.array_equal(ep.frame2emb(8, cu(hash_uniform((2, 3, 64, 64), 3700))).cpu().numpy(), c["emb"])


def test_reconstruct_loss_has_gradient(ep):
    pred = torch.randn(2, 196, 256, device="cuda", requires_grad=True)
    frame = torch.randn(2, 1, 224, 224, device="cuda")
    mask = (torch.rand(2, 196, device="cuda") < 0.75).float()
    ns = SimpleNamespace(patch_size=16, norm_pix_loss=True, mask_ratio=0.75)
    ep.reconstruct_loss(ns, pred, frame, mask).backward()
    assert pred.grad is not None and torch.isfinite(pred.grad).all()


def test_decoder_unshuffle(ep, golden_stage3):
    c = golden_stage3["decoder_unshuffle"]
    out = ep.unshuffle_tokens(cu(c["emb"]), cu(c["mask_token"]), cu(c["ids_restore"]), cu(c["pos_embed"]))
    assert np.array_equal(out.cpu().numpy(), c["x"])


def test_diffmap(ep, golden_stage3):
    from oracle import stage3_np as s3
    c = golden_stage3["frame_time_flip"]
    f = cu(c["frame"][None])
    z = torch.zeros_like(f)
    assert np.array_equal(ep.diffmap_frames(z, f, negate=[1]).cpu().numpy()[0], c["flipped"])
    a = hash_uniform((3, 1, 48, 64), 1) + np.float32(0.6)
    b = hash_uniform((3, 1, 48, 64), 2) + np.float32(0.6)
    neg = [0, 1, 0]
    lin = ep.diffmap_frames(cu(a), cu(b), "linear", negate=neg).cpu().numpy()
    for i in range(3):
        assert np.array_equal(lin[i], s3.diffmap_frames(a[i], b[i], "linear", negate=bool(neg[i])))
    lg = ep.diffmap_frames(cu(a), cu(b), "log", eps=1e-3).cpu().numpy()
    np.testing.assert_allclose(lg, s3.diffmap_frames(a, b, "log", 1e-3), rtol=1e-5, atol=1e-6)


def test_masked_input_pipeline_graph(ep):
    """The CUDA-graph pipeline equals the individual calls, replay after replay."""
    pipe = ep.MaskedInputPipeline(8, 5, (224, 224), 16, 0.75, "cuda")
    for it in range(3):
        torch.manual_seed(it)
        pipe.x.copy_(torch.randn(8, 5, 224, 224, device="cuda"))
        pipe.sub_frame.copy_(torch.randn(8, 1, 224, 224, device="cuda"))
        pipe.noise.copy_(torch.rand(8, 196, device="cuda"))
        out = pipe.run(draw_noise=False)
        ik, m, ir = ep.mask_from_noise(pipe.noise, 49)
        assert torch.equal(out["ids_keep"], ik) and torch.equal(out["mask"], m) and torch.equal(out["ids_restore"], ir)
        assert torch.equal(out["visible_patches"], ep.patchify_gather(pipe.x, 16, ik, "cpq"))
        assert torch.equal(out["target"], ep.target_normpix(pipe.sub_frame, 16))
