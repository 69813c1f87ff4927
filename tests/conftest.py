import os
import sys

import numpy as np
import pytest
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(REPO, "tests", "golden")
for p in (REPO, GOLDEN):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden_ops():
    return np.load(os.path.join(GOLDEN, "ops.npz"))


@pytest.fixture(scope="session")
def golden_models():
    return np.load(os.path.join(GOLDEN, "models.npz"))


@pytest.fixture(scope="session")
def golden_divide():
    return np.load(os.path.join(GOLDEN, "divide.npz"), allow_pickle=False)


def mirror_namespace():
    """The classes cases.build_op_module needs, taken from this package's nn mirror."""
    from hyperseg_b200.nn import hyperseg_v0_1, hyperseg_v1_0, meta_conv, meta_patch
    return {
        "HyperPatchNoPadding": hyperseg_v1_0.HyperPatchNoPadding,
        "make_hyper_patch_conv2d_block": hyperseg_v1_0.make_hyper_patch_conv2d_block,
        "HyperPatchInvertedResidual": hyperseg_v1_0.HyperPatchInvertedResidual,
        "HyperPatchConv2d": hyperseg_v1_0.HyperPatchConv2d,
        "MetaPatchConv2d": meta_patch.MetaPatchConv2d,
        "make_meta_patch_conv2d_block": meta_patch.make_meta_patch_conv2d_block,
        "MetaConv2d": meta_conv.MetaConv2d,
        "V01HyperPatchInvertedResidual": hyperseg_v0_1.HyperPatchInvertedResidual,
    }


def rel_err(a, b):
    """max |a-b| / max |b| -- the 'relative to the largest logit' measure used throughout."""
    a = torch.as_tensor(a, dtype=torch.float64)
    b = torch.as_tensor(b, dtype=torch.float64)
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()
