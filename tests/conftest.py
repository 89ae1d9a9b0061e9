import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")
    config.addinivalue_line("markers", "reference: needs the read-only reference tree at /root/reference")


def pytest_collection_modifyitems(config, items):
    have_ref = os.path.isdir("/root/reference/druglib")
    skip_ref = pytest.mark.skip(reason="/root/reference not present (GPU box)")
    have_gpu = False
    try:
        import torch
        have_gpu = torch.cuda.is_available() and os.path.exists(os.path.join(ROOT, "diffbindfr_b200", "libb200dock.so"))
    except Exception:
        pass
    skip_gpu = pytest.mark.skip(reason="needs a CUDA device and the built libb200dock.so (run with -m gpu on the B200 box)")
    for item in items:
        if "reference" in item.keywords and not have_ref:
            item.add_marker(skip_ref)
        if "gpu" in item.keywords and not have_gpu:
            item.add_marker(skip_gpu)
