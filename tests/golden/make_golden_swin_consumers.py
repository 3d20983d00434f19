"""Golden vectors for the Swin stage-output consumers (model/backbone/swin.py:206-238), produced by EXECUTING the reference's
own statements: the three `if i == 0 / elif i == 1 / elif i == 2` bodies of SwinTransformer.forward are read from the
reference file at generation time, dedented and exec'd with a namespace that supplies x, coords, ids_keep and a `self`
carrying the reference-shaped stage decoders (nn.Conv2d(k = s = 8 / 4 / 2), swin.py:92-94).  Nothing of the reference is
stored in this repository; only inputs and outputs are.

    python tests/golden/make_golden_swin_consumers.py        (needs /root/reference)
"""
import ast
import os
import sys
import textwrap
from types import SimpleNamespace

import numpy as np
import torch
import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg  # noqa: E402


def branch_sources():
    """Source of the bodies of the `i == 0`, `i == 1`, `i == 2` branches inside `if i in self.out_indices:`."""
    path = os.path.join(mg.REF, "model", "backbone", "swin.py")
    src = open(path).read()
    tree = ast.parse(src)
    lines = src.splitlines()
    out = {}
    for node in ast.walk(tree):
        if isinstance(node, ast.If) and isinstance(node.test, ast.Compare) and isinstance(node.test.left, ast.Name) \
                and node.test.left.id == "i" and isinstance(node.test.ops[0], ast.Eq) and isinstance(node.test.comparators[0], ast.Constant):
            k = node.test.comparators[0].value
            if k in (0, 1, 2) and k not in out:
                body = "\n".join(lines[node.body[0].lineno - 1:node.body[-1].end_lineno])
                out[k] = textwrap.dedent(body)
    assert sorted(out) == [0, 1, 2], sorted(out)
    return out


def main():
    mg._import_reference()
    torch.set_num_threads(1)
    torch.manual_seed(7100)
    srcs = branch_sources()
    B, L, K = 2, 49, 12
    noise = torch.rand(B, L)
    ids_keep = torch.argsort(noise, dim=1)[:, :K]
    keep_row = torch.zeros(L, dtype=torch.bool)
    keep_row[ids_keep[0]] = True                      # the mask is batch-shared (swin.py:158): coords from row 0
    flat = {}
    for stage, (G, C, ks) in enumerate([(56, 8, 8), (28, 16, 4), (14, 32, 2)]):
        rep = G // 7
        vis = keep_row.reshape(7, 7).repeat_interleave(rep, 0).repeat_interleave(rep, 1)
        hw = torch.nonzero(vis)                       # (n_vis, 2) row-major, as apply_mask emits them
        perm = torch.randperm(hw.shape[0])            # later stages list tokens in PatchMerging order: any order must work
        coords = hw[perm].unsqueeze(0)
        x = torch.randn(B, hw.shape[0], C)
        dec = nn.Conv2d(C, 24, kernel_size=ks, stride=ks)
        ns = {"torch": torch, "x": x, "coords": coords, "ids_keep": ids_keep,
              "self": SimpleNamespace(stage1_output_decode=dec, stage2_output_decode=dec, stage3_output_decode=dec)}
        with torch.no_grad():
            exec(srcs[stage], ns)
        n = stage + 1
        flat[f"s{n}/x"] = x.numpy()
        flat[f"s{n}/coords"] = coords.numpy()
        flat[f"s{n}/ids_keep"] = ids_keep.numpy()
        flat[f"s{n}/grid"] = np.asarray(G)
        flat[f"s{n}/dense"] = ns[f"_emb_l{n}"].contiguous().numpy()            # (B, C, G, G): what the stage decoder receives
        flat[f"s{n}/weight"] = dec.weight.detach().numpy()
        flat[f"s{n}/bias"] = dec.bias.detach().numpy()
        flat[f"s{n}/emb_stage"] = ns[f"emb_stage{n}"].numpy()                   # (B, K, 24) after the gather by ids_keep
    np.savez_compressed(os.path.join(HERE, "swin_consumers.npz"), **flat)
    print("swin_consumers.npz", os.path.getsize(os.path.join(HERE, "swin_consumers.npz")) // 1024, "KiB")


if __name__ == "__main__":
    main()
