"""Swin sparse-token plumbing: window grouping off the Python critical path.

Drop-ins for model/sub_module/swin_block.py:
  knapsack(W, wt)                               :280-326   -> native (ep_swin_group_windows_host, host C++)
  group_windows(group_size, num_ele_win)        :329-352   -> native
  GroupingModule(window_size, shift_size, ...)  :355-464   -> same attributes and methods; the plan of a mask is computed once
                                                             and cached (the mask is batch-shared, swin.py:158, and changes
                                                             once per step), group()/merge() are one gather kernel each
  PatchMerging's token re-ordering              :196-203   -> patch_merging_order(mask_prev, H, W)

Tie order: the reference sorts tokens by window id with torch.argsort's default (unstable on the CPU, a stable radix sort on
CUDA where the models run).  Here the sort is stable: tokens of a window keep their row-major order.  Attention inside a
group is permutation-equivariant and merge() undoes the shuffle, so the block's output does not depend on that order.
"""
import ctypes
from collections import OrderedDict

import numpy as np
import torch

from . import _lib
from ._runtime import lib


def group_windows(group_size, num_ele_win):
    """(num_ele_group, grouped_idx) exactly as swin_block.py:329-352 returns them (lists)."""
    wt = np.ascontiguousarray(num_ele_win, np.int32)
    n = int(wt.shape[0])
    neg = np.empty(max(n, 1), np.int32)
    first = np.empty(n + 1, np.int32)
    gidx = np.empty(max(n, 1), np.int32)
    ng = ctypes.c_int(0)
    rc = lib().ep_swin_group_windows_host(int(group_size), wt.ctypes.data, n, neg.ctypes.data, first.ctypes.data, gidx.ctypes.data,
                                          ctypes.addressof(ng))
    _lib.check(rc, "ep_swin_group_windows_host")
    g = ng.value
    return [int(v) for v in neg[:g]], [[int(v) for v in gidx[first[i]:first[i + 1]]] for i in range(g)]


def knapsack(W, wt):
    """(best fill, selected indices in increasing order) as swin_block.py:280-326: one step of group_windows."""
    if len(wt) == 0:
        return 0, []
    big = [w for w in wt if w > W]
    if big:        # items that cannot be taken never are: mask them out for the native call (which rejects them)
        keep = [i for i, w in enumerate(wt) if w <= W]
        res, idx = knapsack(W, [wt[i] for i in keep])
        return res, [keep[i] for i in idx]
    neg, groups = group_windows(W, list(wt))
    return neg[0], groups[0]


def patch_merging_order(mask_prev, H, W):
    """idx_shuffle of PatchMerging.forward (swin_block.py:196-203): position of every visible token (row-major order of the
    H x W grid) once tokens are listed 2x2-block by 2x2-block.  mask_prev: bool (1, H*W) or (H*W,)."""
    m = mask_prev.reshape(H // 2, 2, W // 2, 2).permute(0, 2, 1, 3).reshape(-1)
    ii, jj = torch.meshgrid(torch.arange(H, device=m.device), torch.arange(W, device=m.device), indexing="ij")
    key = (ii * H + jj).reshape(H // 2, 2, W // 2, 2).permute(0, 2, 1, 3).reshape(-1)[m]      # coords[:,0] * H + coords[:,1] (:201)
    return torch.argsort(torch.argsort(key, stable=True), stable=True)


class GroupingModule:
    """Drop-in for swin_block.GroupingModule.  prepare() is host/torch logic with the DP in native code and cached per
    (coords, window, shift); group()/merge() run the library's gather kernel on CUDA tensors."""
    _cache = OrderedDict()
    _cache_max = 64

    def __init__(self, window_size, shift_size, group_size=None):
        self.window_size = window_size
        self.shift_size = shift_size
        assert shift_size >= 0 and shift_size < window_size
        self.group_size = group_size or self.window_size ** 2
        self.attn_mask = None
        self.rel_pos_idx = None

    def _window_ids(self, ch):
        """Window id of every visible token: ch (n, 2) int64 numpy (h, w) -> (n,) int64   (swin_block.py:363-366; the row
        multiplier is the token count, as in the reference)."""
        g = (ch + (self.window_size - self.shift_size) % self.window_size) // self.window_size
        return g[:, 0] * ch.shape[0] + g[:, 1]

    def _tables(self, gid, coords, mask_rel, dev):
        """(attn_mask f32, rel_pos_idx i64), each (nG, GS, GS), from the slots' window ids and coordinates: native host code
        (ep_swin_group_tables_host), a few hundred KB at most, shipped to the device once per cached plan."""
        gid = np.ascontiguousarray(gid, np.int64)
        coords = np.ascontiguousarray(coords, np.int64)
        nG, GS = gid.shape
        attn = np.empty((nG, GS, GS), np.float32)
        rel = np.empty((nG, GS, GS), np.int64)
        rc = lib().ep_swin_group_tables_host(gid.ctypes.data, coords.ctypes.data, nG, GS, int(self.window_size), int(mask_rel),
                                             attn.ctypes.data, rel.ctypes.data)
        _lib.check(rc, "ep_swin_group_tables_host")
        return torch.from_numpy(attn).to(dev), torch.from_numpy(rel).to(dev)

    def _prepare_masking(self, coords):
        # few tokens: one group holding all of them, attention restricted to the windows by the mask alone (:393-399)
        ch = coords[0].detach().cpu().numpy().astype(np.int64)
        self.idx_shuffle = None
        self.idx_unshuffle = None
        return self._tables(self._window_ids(ch)[None], ch[None], False, coords.device)

    def _prepare_grouping(self, coords):
        # The index bookkeeping runs on the host in numpy (the coordinates are a few KB; one small D2H copy instead of a
        # chain of tiny device ops and a .tolist() sync), the DP in native code; only the final index / mask tensors go
        # back to the device, where the two (nG, GS, GS) tables are built.
        dev = coords.device
        ch = coords[0].detach().cpu().numpy().astype(np.int64)                       # (N_vis, 2)
        group_id = self._window_ids(ch)
        idx_merge = np.argsort(group_id, kind="stable")
        gid_sorted = group_id[idx_merge]
        change = np.flatnonzero(np.diff(gid_sorted)) + 1
        starts = np.concatenate([[0], change, [gid_sorted.shape[0]]]).astype(np.int64)
        exact_win_sz = np.diff(starts).tolist()
        self.group_size = min(self.window_size ** 2, max(exact_win_sz))
        num_ele_group, grouped_idx = group_windows(self.group_size, exact_win_sz)     # native DP
        GS, nG = self.group_size, len(num_ele_group)
        src = np.full(nG * GS, -1, np.int64)                  # position in idx_merge of every slot, -1 = padding
        for gi, gidx in enumerate(grouped_idx):
            at = gi * GS
            for w in gidx:
                n = exact_win_sz[w]
                src[at:at + n] = np.arange(starts[w], starts[w] + n)
                at += n
        pad = src < 0
        idx_shuffle = np.where(pad, -1, idx_merge[np.maximum(src, 0)])
        amask = np.where(pad, -1, gid_sorted[np.maximum(src, 0)]).reshape(nG, GS)
        idx_unshuffle = np.argsort(idx_shuffle, kind="stable")[-sum(num_ele_group):]
        idx_shuffle = np.where(pad, 0, idx_shuffle)           # index_select does not permit negative index
        self.idx_shuffle = torch.from_numpy(idx_shuffle).to(dev)
        self.idx_unshuffle = torch.from_numpy(np.ascontiguousarray(idx_unshuffle)).to(dev)
        return self._tables(amask, ch[idx_shuffle].reshape(nG, GS, 2), True, dev)

    def prepare(self, coords, num_tokens):
        key = (self.window_size, self.shift_size, int(num_tokens), str(coords.device), coords[:1].cpu().numpy().tobytes())
        hit = GroupingModule._cache.get(key)
        if hit is not None:
            GroupingModule._cache.move_to_end(key)
            self._mode, self.group_size, self.idx_shuffle, self.idx_unshuffle, attn_mask, rel_pos_idx = hit
            return attn_mask, rel_pos_idx
        if num_tokens <= 2 * self.window_size ** 2:
            self._mode = 'masking'
            out = self._prepare_masking(coords)
        else:
            self._mode = 'grouping'
            out = self._prepare_grouping(coords)
        GroupingModule._cache[key] = (self._mode, self.group_size, self.idx_shuffle, self.idx_unshuffle, out[0], out[1])
        while len(GroupingModule._cache) > GroupingModule._cache_max:
            GroupingModule._cache.popitem(last=False)
        return out

    def _select(self, x, idx):
        from .masking import gather_tokens
        return gather_tokens(x, idx)          # batch-shared indices; carries the gradient (padded slots accumulate)

    def group(self, x):
        if self._mode == 'grouping':
            self.ori_shape = x.shape
            x = self._select(x, self.idx_shuffle)                   # (B, nG*GS, C)
            x = x.reshape(-1, self.group_size, x.shape[-1])         # (B*nG, GS, C)
        return x

    def merge(self, x):
        if self._mode == 'grouping':
            B, N, C = self.ori_shape
            x = self._select(x.reshape(B, -1, C), self.idx_unshuffle)   # (B, N, C)
        return x
