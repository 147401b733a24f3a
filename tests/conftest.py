import importlib
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

PKG = "pytorch_empirical-mvm_b200"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def vsw():
    """the product package (directory name has a '-', hence importlib)"""
    return importlib.import_module(PKG)


@pytest.fixture(scope="session")
def oracle():
    from oracle import swin3d_oracle
    return swin3d_oracle


def rel_l2(a: torch.Tensor, b: torch.Tensor) -> float:
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


def pattern_clip(B, Tn, H, W):
    """the deterministic RNG-free clip of tests/golden/make_golden_mvm.py (values in [-0.5, 0.5), none exactly 0)"""
    n = B * Tn * 3 * H * W
    return (((torch.arange(n, dtype=torch.int64) * 7919) % 1013).float() + 0.5).div(1013.0).sub(0.5).view(B, Tn, 3, H, W)
