"""numpy restatement of stages 2-3 of the EventPretrain input path — TEST INFRASTRUCTURE ONLY.

Each function cites the reference file:line it follows.  Parity status: pinned against golden
vectors produced by executing the unmodified reference (tests/golden/make_golden.py ->
tests/golden/stage3_mask_patch.npz; checked in tests/test_oracle_golden.py).  Two items have no
reference implementation and are labelled self-oracle: the frame-difference generator of the
diff-map target and the true time surface (F5).
"""
import numpy as np


# ---- masking -------------------------------------------------------------------------------
def len_keep(L, mask_ratio):
    """model/backbone/vit.py:75 — Python float arithmetic, int() truncation."""
    return int(L * (1 - mask_ratio))


def mask_from_noise(noise, keep):
    """model/backbone/vit.py:91-105 (same code in convvit.py:110-124, swin.py:138-152).

    ids_shuffle = argsort(noise); ids_restore = argsort(ids_shuffle); ids_keep = first `keep`;
    mask = 1 where rank >= keep.  Ties: torch.argsort's default is unstable on CPU for L=196, so
    bit-exact parity is defined on tie-free rows; the contract for tied rows is the stable order.
    """
    noise = np.asarray(noise, np.float32)
    ids_shuffle = np.argsort(noise, axis=1, kind="stable")
    ids_restore = np.argsort(ids_shuffle, axis=1, kind="stable")
    ids_keep = ids_shuffle[:, :keep]
    mask = (ids_restore >= keep).astype(np.float32)
    return ids_keep.astype(np.int64), mask, ids_restore.astype(np.int64)


def rows_with_ties(noise):
    s = np.sort(np.asarray(noise), axis=1)
    return (s[:, 1:] == s[:, :-1]).any(axis=1)


def patch_density(x, p):
    """model/backbone/vit.py:80-83 — AvgPool2d(p,p)(abs(sum_c x)).flatten(1), fp32.

    Accumulation order: channels in order; window row-major; then one division by p*p
    (ATen's CPU avg_pool2d sums into a float accumulator and divides once).
    """
    x = np.asarray(x, np.float32)
    B, C, H, W = x.shape
    s = x[:, 0].copy()
    for c in range(1, C):
        s = s + x[:, c]
    s = np.abs(s)
    gh, gw = H // p, W // p
    win = s[:, :gh * p, :gw * p].reshape(B, gh, p, gw, p).transpose(0, 1, 3, 2, 4).reshape(B, gh * gw, p * p)
    acc = np.zeros((B, gh * gw), np.float32)
    for k in range(p * p):
        acc = acc + win[:, :, k]
    return (acc / np.float32(p * p)).astype(np.float32)


# ---- gathers / patchify --------------------------------------------------------------------
def gather_tokens(tokens, pos_embed, ids_keep):
    """model/backbone/vit.py:113-115 — (tokens + pos_embed) gathered along dim 1 by ids_keep."""
    x = np.asarray(tokens, np.float32)
    if pos_embed is not None:
        x = x + np.asarray(pos_embed, np.float32)[None]
    return np.take_along_axis(x, ids_keep[:, :, None], axis=1)


def patchify(frame, p, order="pqc"):
    """utils/reshape.py:15-22 (order 'pqc': einsum 'bchpwq->bhwpqc'); order 'cpq' is the
    Conv2d(k=s=p) weight order used to embed visible patches only (vit_block.py:44-68)."""
    f = np.asarray(frame)
    B, C, H, W = f.shape
    gh, gw = H // p, W // p
    f = f.reshape(B, C, gh, p, gw, p)
    if order == "pqc":
        f = f.transpose(0, 2, 4, 3, 5, 1)
    elif order == "cpq":
        f = f.transpose(0, 2, 4, 1, 3, 5)
    else:
        raise ValueError(order)
    return np.ascontiguousarray(f.reshape(B, gh * gw, -1))


def patchify_gather(x, p, ids_keep, order="cpq"):
    return np.take_along_axis(patchify(x, p, order), ids_keep[:, :, None], axis=1)


def target_normpix(frame, p, norm_pix=True, eps=1.0e-6):
    """model/pretrain/pr_hub_model.py:125-131 — frame2emb, then per-patch (x-mean)/sqrt(var+eps)
    with torch's default unbiased variance."""
    t = patchify(np.asarray(frame, np.float32), p, "pqc")
    if norm_pix:
        mean = t.mean(axis=-1, keepdims=True, dtype=np.float32)
        var = t.var(axis=-1, keepdims=True, ddof=1, dtype=np.float32)
        t = (t - mean) / np.sqrt(var + np.float32(eps), dtype=np.float32)
    return t.astype(np.float32)


def masked_mse(pred, target, mask, mask_ratio=0.75):
    """model/pretrain/pr_hub_model.py:133-139."""
    l = ((np.asarray(pred, np.float32) - target) ** 2).mean(axis=-1, dtype=np.float32)
    if mask_ratio == 0:
        return l.mean(dtype=np.float32)
    return (mask * l).sum(dtype=np.float32) / mask.sum(dtype=np.float32)


# ---- block masks ---------------------------------------------------------------------------
def block_mask_expand(mask, grid, rep, invert=True):
    """model/backbone/convvit.py:129-130,142-143 — (B, grid*grid) mask, each cell repeated rep x rep
    -> (B,1,grid*rep,grid*rep); ConvBlock receives 1 - mask (convvit.py:133,146), i.e. invert=True."""
    m = np.asarray(mask, np.float32).reshape(-1, grid, grid)
    m = np.repeat(np.repeat(m, rep, axis=1), rep, axis=2)[:, None]
    return (1.0 - m).astype(np.float32) if invert else m


def swin_apply_mask(x, mask_bool, res):
    """model/backbone/swin.py:154-179 — only mask[:1] is used (batch-shared); expansion of the
    (Mh,Mw) mask to the token grid; row-major boolean compaction of tokens and (h,w) coords."""
    x = np.asarray(x)
    B, N, C = x.shape
    H, W = res
    m = np.asarray(mask_bool, bool)[:1]
    up = N // m.shape[1]
    assert up * m.shape[1] == N
    r = int(up ** 0.5)
    if up > 1:
        Mh, Mw = H // r, W // r
        m = np.broadcast_to(m.reshape(1, Mh, 1, Mw, 1), (1, Mh, r, Mw, r)).reshape(1, -1)
    vis = ~m
    idx = np.nonzero(vis[0])[0]
    coords = np.stack([idx // W, idx % W], axis=-1)[None].astype(np.int64)
    return x[:, idx, :], coords, vis


def decoder_unshuffle(emb, mask_token, ids_restore, pos_embed):
    """model/pretrain/pr_rec_decoder.py:56-62 — append mask tokens, gather by ids_restore, add pos."""
    emb = np.asarray(emb, np.float32)
    B, K, D = emb.shape
    L = ids_restore.shape[1]
    full = np.concatenate([emb, np.broadcast_to(np.asarray(mask_token, np.float32), (B, L - K, D))], axis=1)
    out = np.take_along_axis(full, ids_restore[:, :, None], axis=1)
    return out + np.asarray(pos_embed, np.float32)[None]


# ---- difference-map target -----------------------------------------------------------------
def voxel_sum(voxel):
    """dataset/pretrain/pr_ef_imagenet_dataset.py:192-193 — event-side integral voxel.sum(0)[None]."""
    return np.asarray(voxel, np.float32).sum(axis=-3, keepdims=True, dtype=np.float32)


def diffmap_frames(f0, f1, mode="linear", eps=1.0e-3, negate=False):
    """SELF-ORACLE (the reference loads pre-computed sub_frame files; generator absent, F6).
    T = g(f1) - g(f0), g = identity or log(. + eps); negated under time reversal
    (dataset/augmentation/view_augment.py:60-63)."""
    f0 = np.asarray(f0, np.float32)
    f1 = np.asarray(f1, np.float32)
    if mode == "log":
        d = np.log(f1 + np.float32(eps)) - np.log(f0 + np.float32(eps))
    else:
        d = f1 - f0
    return (-d if negate else d).astype(np.float32)


def time_surface(xs, ys, ts, ps, size, tau, t_ref=None):
    """SELF-ORACLE (no time-surface routine exists in the reference, F5): per-polarity
    exp(-(t_ref - t_last)/tau) of the most recent event at each pixel; 0 where no event."""
    H, W = size
    last = np.full((2, H, W), -np.inf)
    ts = np.asarray(ts, np.float64)
    t_ref = float(ts[-1]) if t_ref is None else t_ref
    ch = np.where(np.asarray(ps) == 1, 0, 1)
    np.maximum.at(last, (ch, np.asarray(ys, np.int64), np.asarray(xs, np.int64)), ts)
    out = np.exp(-(t_ref - last) / tau)
    out[~np.isfinite(last)] = 0.0
    return out.astype(np.float32)


# ---- Swin sparse-token grouping ---------------------------------------------------
def swin_knapsack(W, wt):
    """model/sub_module/swin_block.py:280-326 — 0/1 knapsack with value = weight; returns (best fill, selected indices in
    increasing order).  Ties in the back-tracking go to "not taken" exactly like the reference (res == K[i-1][w])."""
    n = len(wt)
    K = np.zeros((n + 1, W + 1), np.int64)
    for i in range(1, n + 1):
        for w in range(1, W + 1):
            K[i, w] = K[i - 1, w]
            if wt[i - 1] <= w:
                K[i, w] = max(wt[i - 1] + K[i - 1, w - wt[i - 1]], K[i - 1, w])
    res = int(K[n, W])
    left, w, idx = res, W, []
    for i in range(n, 0, -1):
        if left <= 0:
            break
        if left == K[i - 1, w]:
            continue
        idx.append(i - 1)
        left -= wt[i - 1]
        w -= wt[i - 1]
    return res, idx[::-1]


def swin_group_windows(group_size, num_ele_win):
    """model/sub_module/swin_block.py:329-352 — greedy: one knapsack per group over the windows still ungrouped."""
    wt, ori = list(num_ele_win), list(range(len(num_ele_win)))
    groups, fills = [], []
    while wt:
        res, idx = swin_knapsack(group_size, wt)
        fills.append(res)
        groups.append([ori[i] for i in idx])
        keep = [i for i in range(len(wt)) if i not in idx]
        wt, ori = [wt[i] for i in keep], [ori[i] for i in keep]
    return fills, groups
