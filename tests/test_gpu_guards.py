"""compute-sanitizer is closed on this pool, so out-of-bounds global writes of the routed kernels are looked for directly: the
workspace and every output tensor are carved out of larger buffers whose margins hold a byte pattern; after the calls the
margins must be untouched.  Run on awkward shapes (ragged tails, a narrower last tile, empty samples, offsets[0] > 0)."""
import numpy as np
import pytest
import torch

from test_gpu_tiled import dense_batch

pytestmark = pytest.mark.gpu
GUARD = 1 << 20


@pytest.fixture(scope="module")
def ep(native_lib):
    import eventpretrain_b200 as ep
    return ep


def guarded(shape, dtype):
    n = int(np.prod(shape)) * torch.empty((), dtype=dtype).element_size()
    raw = torch.full((n + 2 * GUARD,), 0xA5, dtype=torch.uint8, device="cuda")
    view = raw[GUARD:GUARD + n].view(dtype).reshape(shape)
    return raw, view, n


def margins_intact(raw, n):
    return bool((raw[:GUARD] == 0xA5).all()) and bool((raw[GUARD + n:] == 0xA5).all())


def install_guarded_workspace(ep, tag, dev, nbytes):
    """Pre-seeds the package's scratch cache for (device, stream, tag) with a guarded slice of exactly nbytes."""
    from eventpretrain_b200 import _runtime
    n = (int(nbytes) + 255) // 256 * 256
    raw = torch.full((n + 2 * GUARD,), 0xA5, dtype=torch.uint8, device=dev)
    key = (torch.device(dev).index, torch.cuda.current_stream(dev).cuda_stream, tag)
    _runtime._workspaces[key] = raw[GUARD:GUARD + n]
    return raw, n, key


@pytest.mark.parametrize("H,W,bins", [(480, 640, 5), (65, 87, 9), (33, 50, 1), (224, 224, 5)])
def test_tiled_binning_stays_inside_its_buffers(ep, H, W, bins):
    import ctypes
    from eventpretrain_b200 import _runtime
    from eventpretrain_b200.events import _bin_params
    rng = np.random.default_rng(H)
    ev, _ = dense_batch(ep, rng, [30011, 0, 8193, 1, 70000, 255], H, W, hot=300)
    p4 = ev.packed(4).to("cuda").shard(1, 2)                     # offsets[0] > 0: the record array starts mid-chunk
    B = p4.batch
    L = ep.load_library()
    desc = p4._desc()
    prm = _bin_params((H, W), bins, 0, (1.0, 1.0), False, "tiled")
    need = L.ep_bin_events_workspace_bytes_for(ctypes.byref(desc), ctypes.byref(prm))
    ref = ep.bin_events(p4, (H, W), num_bins=bins, voxel_sum=True, method="global")
    raw_ws, n_ws, key = install_guarded_workspace(ep, "bin", p4.device, need)
    raw_v, vox, n_v = guarded((B, bins, H, W), torch.float32)
    raw_s, vsum, n_s = guarded((B, 1, H, W), torch.float32)
    raw_t, stats, n_t = guarded((bins + 1, 4), torch.float64)
    try:
        for _ in range(2):
            out = ep.bin_events(p4, (H, W), num_bins=bins, voxel_sum=True, method="tiled", stats=True,
                                out={"voxel": vox, "voxel_sum": vsum, "stats": stats}, check=True)
        torch.cuda.synchronize()
    finally:
        _runtime._workspaces.pop(key, None)
    assert torch.equal(out["voxel"], ref["voxel"]) and torch.equal(out["voxel_sum"], ref["voxel_sum"])
    assert margins_intact(raw_ws, n_ws), "the tiled kernels wrote outside their workspace"
    assert margins_intact(raw_v, n_v) and margins_intact(raw_s, n_s) and margins_intact(raw_t, n_t), "an output was overrun"


@pytest.mark.parametrize("H,W", [(100, 131), (260, 346), (7, 9)])
def test_routed_evrep_stays_inside_its_buffers(ep, H, W):
    import ctypes
    from eventpretrain_b200 import _runtime
    rng = np.random.default_rng(W)
    ev, _ = dense_batch(ep, rng, [40001, 0, 3, 65000, 8192], H, W, hot=500)
    p4 = ev.packed(4).to("cuda").shard(1, 2)
    L = ep.load_library()
    desc = p4._desc()
    need = L.ep_evrep_workspace_bytes_for(ctypes.byref(desc), H, W)
    ref = ep.evrep(ev.to("cuda").shard(1, 2), (H, W))
    raw_ws, n_ws, key = install_guarded_workspace(ep, "evrep", p4.device, need)
    raw_o, rep, n_o = guarded((p4.batch, 3, H, W), torch.float64)
    try:
        for _ in range(2):
            out = ep.evrep(p4, (H, W), out=rep, check=True)
        torch.cuda.synchronize()
    finally:
        _runtime._workspaces.pop(key, None)
    assert torch.equal(out, ref)
    assert margins_intact(raw_ws, n_ws), "the routed EvRep wrote outside its workspace"
    assert margins_intact(raw_o, n_o), "the EvRep output was overrun"
