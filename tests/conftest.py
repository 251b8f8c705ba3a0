"""pytest configuration: registers the `gpu` marker and puts the repo root on sys.path.

`-m "not gpu"` runs on CPU only (oracle vs golden vectors, host logic, C-ABI symbol export).
`-m gpu` tests are the parity tests proper; they call the CUDA path through the C ABI and need a B200."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200, sm_100a)")


def pytest_sessionstart(session):
    """the C-ABI library is a build artefact (git-ignored): compile it when a fresh checkout has none"""
    try:
        from qoc_b200 import _lib
        if not os.path.exists(_lib.LIB_PATH):
            _lib.build_library(verbose=True)
    except Exception as exc:                      # the tests that need it will fail loudly
        print("qoc_b200: could not build libqocb200.so:", exc)


def _cuda_available():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _cuda_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this environment")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
