"""Shared pytest configuration.

* registers the `gpu` marker (tests that need a real B200),
* puts the product module directory `irl-maxent_b200/` (a directory of
  top-level modules, exactly like the reference's `src/`) and the repo root
  (for `oracle`) on sys.path.
"""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "irl-maxent_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def golden():
    def load(name):
        return np.load(os.path.join(GOLDEN, name + ".npz"))
    return load


def has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.fixture(scope="session", autouse=True)
def _native_library():
    """The tests exercise the in-tree shared library; build it if the tree was checked out without it
    (nvcc cross-compiles for sm_100a without a GPU).  The product itself never auto-builds or falls back."""
    lib = os.path.join(PKG, "lib", "libirlmaxent_b200.so")
    if not os.path.exists(lib):
        import importlib.util
        spec = importlib.util.spec_from_file_location("build_native", os.path.join(PKG, "build_native.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        mod.build()
    yield
