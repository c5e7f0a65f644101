"""CPU-side checks of the C-ABI boundary: the library builds/loads and exports every symbol that
include/mht_b200.h declares.  No compute calls (no GPU here)."""
import ctypes
import os
import re

import pytest

from conftest import ROOT


def test_library_exports_every_declared_symbol():
    from pymht_b200 import build, _lib
    build.build()
    lib = _lib.load()
    header = open(os.path.join(ROOT, "include", "mht_b200.h")).read()
    declared = set(re.findall(r"\b(mht_[a-z_]+)\s*\(", header))
    assert len(declared) >= 15
    for name in declared:
        assert hasattr(lib, name), "missing export " + name
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)
    assert lib.mht_version() >= 100


def test_struct_layouts_match_header():
    from pymht_b200 import _lib
    assert ctypes.sizeof(_lib.Model) == 16 * 4 + 16 * 4 + 8 * 4 + 4 * 4 + 16
    assert ctypes.sizeof(_lib.ForestConfig) == ctypes.sizeof(_lib.Model) + 4 * 3 + 4 + 8 * 2 + 8 * 4 + 16 + 8
    assert ctypes.sizeof(_lib.ScanInfo) == 8 * 3 + 4 * 6 + 8 * 2 + 8 * 2 + 4 * 6 + 8 + 8 + 4 * 2 + 8 + 4 * 4


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from pymht_b200 import _lib
    from pymht_b200.models import pv
    from pymht_b200.tracker import Tracker
    lib = _lib.load()
    assert lib.mht_device_count() == 0
    with pytest.raises(_lib.MhtError) as e:
        Tracker(pv, 2.5, 1e-4, 1e-9, initiator=None)
    assert e.value.code == _lib.MHT_E_NODEVICE
    # the M-of-N initiator has no CPU path either (its constructor creates the device buffers)
    from pymht_b200.initiators import m_of_n
    with pytest.raises(_lib.MhtError) as e:
        m_of_n.Initiator(2, 3, 20, pv.C_RADAR, pv.R_RADAR(), 25.0)
    assert e.value.code == _lib.MHT_E_NODEVICE
    with pytest.raises(_lib.MhtError) as e:
        Tracker(pv, 2.5, 1e-4, 1e-9)              # default: initiator live, like the reference
    assert e.value.code == _lib.MHT_E_NODEVICE


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "pymht_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("no oracle", ""), f
