"""CPU-side checks of the boundary: the C-ABI library builds (nvcc cross-compiles), loads, and exports every
symbol include/eventpretrain_b200.h declares; the Python binding lists every one of them; host-side logic."""
import ctypes
import os
import re
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from conftest import ROOT


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "eventpretrain_b200.h")).read()
    return sorted(set(re.findall(r"^EP_API [a-z_ \*0-9]+?\b(ep_[a-z0-9_]+)\(", text, flags=re.M)))


def test_header_declares_entry_points():
    syms = declared_symbols()
    assert len(syms) >= 20 and "ep_bin_events" in syms and "ep_mask_from_noise" in syms


def test_library_exports_every_declared_symbol(native_lib):
    from eventpretrain_b200 import _lib
    raw = ctypes.CDLL(_lib.LIB_PATH)
    for s in declared_symbols():
        assert hasattr(raw, s), f"{s} declared in the header but not exported"
    assert sorted(_lib.SIGNATURES) == declared_symbols()
    assert native_lib.ep_abi_version() == 1
    assert native_lib.ep_status_string(-2) == b"workspace too small"


def test_workspace_queries_need_no_gpu(native_lib):
    from eventpretrain_b200 import _lib
    prm = _lib.BinParams(480, 640, 5, 0, 1.0, 1.0, 0, 0)
    mn = ctypes.c_size_t()
    rec = native_lib.ep_bin_events_workspace_bytes(ctypes.byref(prm), 256, ctypes.byref(mn))
    slot = 5 * 480 * 640 * 8
    assert mn.value >= slot and rec >= mn.value and rec <= 256 * slot + (1 << 20)
    assert native_lib.ep_evrep_workspace_bytes(2, 44, 64, 1000) >= 4 * 2 * 44 * 64 * 4 + 8000
    assert native_lib.ep_bin_events_workspace_bytes(ctypes.byref(_lib.BinParams(0, 0, 5, 0, 1.0, 1.0, 0, 0)), 1, None) == 0


def test_no_cpu_fallback():
    import eventpretrain_b200 as ep
    if torch.cuda.is_available():
        pytest.skip("CPU-only check")
    with pytest.raises(RuntimeError):
        ep.events_to_voxel_grid(SimpleNamespace(num_bins=5), np.zeros((4, 4)), (4, 4))
    with pytest.raises(RuntimeError):
        ep.mask_from_noise(torch.rand(2, 196), 49)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "eventpretrain_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dp, f)).read()
                assert "import oracle" not in src and "from oracle" not in src and "ep_oracle" not in src, f


def test_pack_events_and_shard():
    import eventpretrain_b200 as ep
    rng = np.random.default_rng(0)
    samples = [np.stack([rng.integers(0, 64, n), rng.integers(0, 48, n), np.sort(rng.uniform(0, 1, n)),
                         rng.integers(0, 2, n)], 1).astype(np.float64) for n in (10, 0, 7, 3)]
    ev = ep.pack_events(samples, pin=False)
    assert ev.batch == 4 and ev.num_events == 20 and ev.x.dtype == torch.uint16 and ev.p.dtype == torch.uint8
    assert list(ev.offsets_host) == [0, 10, 10, 17, 20]
    s1 = ev.shard(1, 2)
    assert s1.batch == 2 and list(s1.offsets_host) == [10, 17, 20] and s1.num_events == 10
    with pytest.raises(ValueError):
        ep.pack_events([np.array([[0.5, 1, 0.0, 1]])], pin=False)
    with pytest.raises(ValueError):
        ep.pack_events([np.array([[1, 1, 0.0, -1]])], pin=False)
    jit = ep.pack_events([np.array([[0.5, 1, 0.0, -1]])], canonical=False, pin=False)
    assert jit.x.dtype == torch.float64


def test_host_side_augmentation_params(golden_stage1):
    import eventpretrain_b200 as ep
    args = SimpleNamespace(fix_events_num=100, val_fix_events_num=50)
    ev = np.zeros((1000, 4))
    a = ep.get_random_index(args, ev, True, seed=3)
    np.random.seed(3)
    s = np.random.randint(0, 900)
    assert a == (s, s + 100)
    assert ep.get_random_index(args, ev[:40], False) == (0, 40)
    c = golden_stage1["reshape_trap"]
    e = c["events"].copy()
    out = ep.events_reshape(e, 640, 480, 224, 224)
    assert out is e and np.array_equal(e[:, 0], c["events"][:, 0] * (224 / 640))
    assert ep.reshape_scale(640, 480, 224, 224) == (224 / 640, 224 / 480)
    assert ep.len_keep_of(196, 0.75) == 49 and ep.len_keep_of(49, 0.75) == 12 and ep.len_keep_of(196, 0.9) == 19
