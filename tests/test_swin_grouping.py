"""SURVEY.md §8 row f4 — Swin window grouping: oracle and native code against golden vectors produced by executing the
reference (tests/golden/make_golden.py::swin_grouping_cases), drop-in GroupingModule plans, PatchMerging token order."""
import numpy as np
import pytest
import torch

from oracle import stage3_np as s3


@pytest.fixture(scope="module")
def golden(golden_swin_grouping, native_lib):
    return golden_swin_grouping


@pytest.fixture(scope="module")
def ep(native_lib):
    import eventpretrain_b200 as ep
    return ep


def _groups(c):
    return [int(v) for v in c["fills"]], [[int(v) for v in c["idx"][c["first"][i]:c["first"][i + 1]]] for i in range(len(c["fills"]))]


def test_group_windows_oracle_and_native(ep, golden):
    n = 0
    for name, c in golden.items():
        if not name.startswith("gw"):
            continue
        gs, wt = int(c["group_size"]), [int(v) for v in c["wt"]]
        want = _groups(c)
        assert s3.swin_group_windows(gs, wt) == want, name             # oracle pinned to the reference
        assert ep.group_windows(gs, wt) == want, name                  # native C++ (C ABI, host function)
        assert ep.knapsack(gs, wt) == (want[0][0], want[1][0]), name   # the first group is one knapsack
        n += 1
    assert n == 24
    assert ep.group_windows(5, []) == ([], [])
    with pytest.raises(RuntimeError):
        ep.group_windows(4, [5])            # a window larger than the group: the reference never terminates on it


def test_grouping_module_plans(ep, golden):
    ep.GroupingModule._cache.clear()
    for name, c in golden.items():
        if name.startswith("gw"):
            continue
        coords = torch.from_numpy(c["coords"])
        gm = ep.GroupingModule(int(c["window"]), int(c["shift"]))
        for _ in range(2):          # second call comes from the plan cache
            attn_mask, rel_pos_idx = gm.prepare(coords.clone(), coords.shape[1])
            assert (gm._mode == "grouping") == bool(c["grouping"]), name
            assert np.array_equal(attn_mask.numpy(), c["attn_mask"]), name
            assert np.array_equal(rel_pos_idx.numpy(), c["rel_pos_idx"]), name
            if gm._mode == "grouping":
                assert gm.group_size == int(c["group_size"])
                assert np.array_equal(gm.idx_shuffle.numpy(), c["idx_shuffle"]), name
                assert np.array_equal(gm.idx_unshuffle.numpy(), c["idx_unshuffle"]), name
        res = int(c["res"])
        order = ep.patch_merging_order(torch.from_numpy(c["vis"]), res, res)
        assert np.array_equal(order.numpy(), c["merge_order"]), name
    assert len(ep.GroupingModule._cache) == 5


@pytest.mark.gpu
def test_group_merge_on_device(ep, golden):
    c = golden["g56_24_s3"]
    coords = torch.from_numpy(c["coords"]).cuda()
    gm = ep.GroupingModule(int(c["window"]), int(c["shift"]))
    gm.prepare(coords, coords.shape[1])
    x = torch.randn(3, coords.shape[1], 96, device="cuda")
    g = gm.group(x)
    want = torch.index_select(x, 1, gm.idx_shuffle).reshape(-1, gm.group_size, 96)      # swin_block.py:452-457
    assert torch.equal(g, want)
    back = gm.merge(g)
    assert torch.equal(back, x)                                                          # :459-464
