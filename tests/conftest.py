import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    """gpu-marked tests are skipped (not errored) where no CUDA device exists, e.g. a plain `pytest` in the build container."""
    try:
        import torch
        has_cuda = torch.cuda.is_available()
    except Exception:
        has_cuda = False
    if has_cuda:
        return
    skip = pytest.mark.skip(reason="needs a CUDA device (B200); run with -m gpu on the GPU box")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


def _load(name):
    z = np.load(os.path.join(ROOT, "tests", "golden", name))
    cases = {}
    for k in z.files:
        case, field = k.split("/", 1)
        cases.setdefault(case, {})[field] = z[k]
    return cases


@pytest.fixture(scope="session")
def golden_stage1():
    return _load("stage1_events.npz")


@pytest.fixture(scope="session")
def golden_stage3():
    return _load("stage3_mask_patch.npz")


@pytest.fixture(scope="session")
def golden_stream_aug():
    return _load("stream_aug.npz")


@pytest.fixture(scope="session")
def golden_views():
    return _load("views.npz")


@pytest.fixture(scope="session")
def golden_swin_grouping():
    return _load("swin_grouping.npz")


@pytest.fixture(scope="session")
def golden_swin_consumers():
    return _load("swin_consumers.npz")


@pytest.fixture(scope="session")
def native_lib():
    """Builds (if stale) and loads the CUDA C-ABI library; nvcc cross-compiles without a GPU."""
    from eventpretrain_b200 import build, _lib
    build.build()
    return _lib.load()


def reshaped(case):
    """Apply events_reshape the way the golden generator did (in-place fp64 multiply by the Python double)."""
    ev = case["events"].copy()
    rs = case["reshape"]
    if rs[0]:
        ev[:, 0] *= (rs[2] / rs[0])
        ev[:, 1] *= (rs[3] / rs[1])
    return ev
