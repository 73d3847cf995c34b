"""pytest configuration: registers the ``gpu`` marker and puts the product package on sys.path.

``-m "not gpu"`` tests run on CPU (oracle vs golden vectors, host logic, C-ABI symbol export);
``-m gpu`` tests are the parity tests proper and call the CUDA path through the C-ABI.
"""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "rho-diffusion_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


def load_golden(name):
    import torch

    return torch.load(os.path.join(GOLDEN, name), map_location="cpu", weights_only=False)
