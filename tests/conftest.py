import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "reference: needs the live reference tree at /root/reference")


def pytest_collection_modifyitems(config, items):
    """`-m gpu` tests need an sm_100 device: on a box without one (this authoring container) a plain `pytest tests`
    skips them instead of failing in the first library call (the library has no CPU path: MHT_E_NODEVICE)."""
    try:
        from pymht_b200 import _lib
        have = _lib.load().mht_device_count() > 0
    except Exception:
        have = False
    if have:
        return
    skip = pytest.mark.skip(reason="no sm_100 (B200) device visible")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def golden(name):
    import numpy as np
    return np.load(os.path.join(GOLDEN, name + ".npz"))
