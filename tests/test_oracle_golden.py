"""The CPU oracle (oracle/) against golden vectors produced by executing the unmodified reference
(tests/golden/make_golden.py).  This is what pins the oracle; everything on the GPU is then checked
against the oracle and against the same golden vectors."""
import numpy as np
import pytest

from conftest import reshaped
from oracle import events as oe
from oracle import stage3_np as s3
from synth import hash_uniform

STAGE1 = ["all_negative", "all_positive", "bin_boundary", "c1_small", "dsec_15bins", "f32_events", "fractional_xy",
          "hot_pixel", "mvsec_9bins", "num_bins_1", "num_bins_2", "pm1_polarity", "reshape_mvsec", "reshape_trap",
          "same_timestamp", "single_event", "two_events", "unsorted_time", "wrap_x"]


@pytest.mark.parametrize("name", STAGE1)
def test_stage1_bit_exact(golden_stage1, name):
    c = golden_stage1[name]
    ev = reshaped(c)
    size, bins = tuple(int(v) for v in c["size"]), int(c["bins"])
    assert np.array_equal(oe.voxel_grid(ev, bins, size), c["voxel"])
    ecdp = oe.count_frame(ev, size, 2)
    mem = oe.count_frame(ev, size, 3)
    assert np.array_equal(ecdp, c["ecdp"])
    assert np.array_equal(mem, c["mem"])
    hot = oe.remove_hot_pixel_mem(mem / np.float32(255))
    assert np.array_equal(hot, c["mem_hot"])
    assert np.array_equal(oe.count_normalise(ecdp), c["ecdp_norm"])
    if "mem_norm" in c:
        assert np.array_equal(oe.mem_normalise(hot), c["mem_norm"])
    if "evrep" in c:
        xs, ys = ev[:, 0].astype(np.int16), ev[:, 1].astype(np.int16)
        t = ev[:, 2].astype(np.float64)
        res = (size[1], size[0])
        assert np.array_equal(oe.evrep(xs, ys, t, ev[:, 3], res), c["evrep"], equal_nan=True)
        assert np.array_equal(oe.evrep(xs, ys, t * 1e6, ev[:, 3], res), c["evrep_us"], equal_nan=True)


def test_stage1_errors():
    ev = np.array([[9, 0, 0.0, 1], [0, 0, 1.0, 0]], np.float64)
    with pytest.raises(IndexError):
        oe.count_frame(ev * [1, 100, 1, 1] + [0, 7, 0, 0], (4, 4), 2)   # y = 7 -> flat index >= H*W
    with pytest.raises(IndexError):
        oe.voxel_grid(np.array([[0, 0, 0.0, 1], [3, 99, 1.0, 1]], np.float64), 5, (4, 4))
    with pytest.raises(ValueError):
        oe.voxel_grid(np.zeros((0, 4)), 5, (4, 4))


def test_batch_driver_matches_single(golden_stage1):
    c = golden_stage1["c1_small"]
    ev = c["events"]
    parts = [ev[:7000], ev[7000:7001], ev[7001:]]
    off = np.cumsum([0] + [len(p) for p in parts])
    out = oe.voxel_grid_batch(ev, off, 5, (180, 240), num_threads=3)
    for b, p in enumerate(parts):
        assert np.array_equal(out[b], oe.voxel_grid(p, 5, (180, 240)))
    cnt = oe.count_frame_batch(ev, off, (180, 240), 2, num_threads=2)
    for b, p in enumerate(parts):
        assert np.array_equal(cnt[b], oe.count_frame(p, (180, 240), 2))


@pytest.mark.parametrize("name", ["mask_random_L196_r75", "mask_random_L196_r50", "mask_random_L49_r75", "mask_random_L196_r90"])
def test_random_masking(golden_stage3, name):
    c = golden_stage3[name]
    assert not s3.rows_with_ties(c["noise"]).any()
    ik, m, ir = s3.mask_from_noise(c["noise"], s3.len_keep(int(c["L"]), float(c["ratio"])))
    assert np.array_equal(ik, c["ids_keep"]) and np.array_equal(m, c["mask"]) and np.array_equal(ir, c["ids_restore"])


@pytest.mark.parametrize("strategy", ["density", "anti-density"])
@pytest.mark.parametrize("L,p", [(196, 16), (49, 32)])
def test_density_masking(golden_stage3, strategy, L, p):
    c = golden_stage3[f"mask_{strategy}_L{L}"]
    x = hash_uniform(tuple(c["shape"]), int(c["seed"]))
    d = s3.patch_density(x, p)
    assert np.array_equal(d, c["density"])          # bit-exact fp32 accumulation order
    noise = d if strategy == "density" else -d
    ik, m, ir = s3.mask_from_noise(noise, s3.len_keep(L, 0.75))
    assert np.array_equal(ik, c["ids_keep"]) and np.array_equal(m, c["mask"]) and np.array_equal(ir, c["ids_restore"])


def test_vit_gather(golden_stage3):
    c = golden_stage3["vit_gather"]
    ik, m, ir = s3.mask_from_noise(c["noise"], 49)
    assert np.array_equal(m, c["mask"]) and np.array_equal(ir, c["ids_restore"])
    assert np.array_equal(s3.gather_tokens(c["tokens"], c["pos_embed"], ik), c["gathered"])


def test_convvit_masks(golden_stage3):
    c = golden_stage3["convvit_masks"]
    _, m, _ = s3.mask_from_noise(c["noise"], 49)
    assert np.array_equal(m, c["mask"])
    assert np.array_equal(s3.block_mask_expand(m, 14, 4), c["keep_mask_56"])
    assert np.array_equal(s3.block_mask_expand(m, 14, 2), c["keep_mask_28"])


def test_swin_apply_mask(golden_stage3):
    c = golden_stage3["swin_apply_mask"]
    x = hash_uniform(tuple(c["shape"]), int(c["seed"]))
    xv, co, vm = s3.swin_apply_mask(x, c["mask"].astype(bool), (56, 56))
    assert np.array_equal(xv, c["x_vis"]) and np.array_equal(co, c["coords"]) and np.array_equal(vm, c["vis_mask"])
    c2 = golden_stage3["swin_apply_mask_196"]
    xv, co, vm = s3.swin_apply_mask(x[:2], c2["mask"], (56, 56))
    assert np.array_equal(xv, c2["x_vis"]) and np.array_equal(co, c2["coords"]) and np.array_equal(vm, c2["vis_mask"])
    c3 = golden_stage3["swin_apply_mask_49"]
    xv, co, vm = s3.swin_apply_mask(hash_uniform((2, 49, 8), 3403), c3["mask"], (7, 7))
    assert np.array_equal(xv, c3["x_vis"]) and np.array_equal(co, c3["coords"]) and np.array_equal(vm, c3["vis_mask"])


def target_inputs(p, L):
    frame = hash_uniform((4, 1, 224, 224), 3500 + p).copy()
    frame[1] *= np.float32(1e-4)
    frame[2, :, :p, :p] = 0.25
    pred = hash_uniform((4, L, p * p), 3600 + p) * np.float32(4)
    return frame, pred


@pytest.mark.parametrize("p,L", [(16, 196), (32, 49)])
def test_target(golden_stage3, p, L):
    c = golden_stage3[f"target_p{p}"]
    frame, pred = target_inputs(p, L)
    assert np.array_equal(s3.patchify(frame, p)[:1], c["emb0"])
    for norm in (1, 0):
        t = s3.target_normpix(frame, p, bool(norm))
        np.testing.assert_allclose(s3.masked_mse(pred, t, c["mask"]), c[f"loss_norm{norm}"], rtol=1e-5)
    t = s3.target_normpix(frame, p, True)
    np.testing.assert_allclose(s3.masked_mse(pred, t, c["mask"], 0), c["loss_nomask"], rtol=1e-5)
    ref = c["per_patch_loss_stride7"]
    sel = ref != 0
    np.testing.assert_allclose(((pred - t) ** 2).mean(-1)[sel], ref[sel], rtol=1e-5)


def test_frame2emb_multichannel(golden_stage3):
    c = golden_stage3["frame2emb_c3"]
    assert np.array_equal(s3.patchify(hash_uniform((2, 3, 64, 64), 3700), 8), c["emb"])


def test_decoder_unshuffle(golden_stage3):
    c = golden_stage3["decoder_unshuffle"]
    out = s3.decoder_unshuffle(c["emb"], c["mask_token"], c["ids_restore"], c["pos_embed"])
    assert np.array_equal(out, c["x"])


def test_time_flip_sign(golden_stage3):
    c = golden_stage3["frame_time_flip"]
    f = hash_uniform((1, 8, 8), 3800)
    assert np.array_equal(c["frame"], f)
    z = np.zeros_like(f)
    assert np.array_equal(s3.diffmap_frames(z, f, negate=True), c["flipped"])
