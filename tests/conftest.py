import importlib
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "reference: needs /root/reference (build container only)")


@pytest.fixture(scope="session")
def hn():
    """The product package (directory name has hyphens, so it is imported by path).  Builds the CUDA library with nvcc
    if it is missing or stale (the build is a cross-compile: no GPU needed)."""
    mod = importlib.import_module("nerf-3dtalker-code_b200")
    mod.build_library()
    return mod


@pytest.fixture(scope="session")
def oracle():
    from oracle import headnerf_oracle
    return headnerf_oracle
