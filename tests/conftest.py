import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests", "golden")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_cuda = torch.cuda.is_available()
    except Exception:  # pragma: no cover
        has_cuda = False
    if has_cuda:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(params=["mixed-radix", "legacy"])
def generic_kernel(request, monkeypatch):
    """The two generic kernel families of csrc: the mixed-radix team kernel (default) and, with SPECINV_GENERIC_MR=0,
    the radix-2^2 CTA-wide kernel (powers of two) / the direct DFT (everything else)."""
    monkeypatch.setenv("SPECINV_GENERIC_MR", "1" if request.param == "mixed-radix" else "0")
    return request.param
