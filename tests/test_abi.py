"""CPU-side check that the C-ABI library loads and exports every symbol that
include/nkfb200.h declares (no compute calls without a GPU)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "nkfb200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(nkfb_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from netket_fidelity_b200 import _native
    lib = _native.load_library()
    names = _declared()
    assert len(names) >= 14
    for n in names:
        assert hasattr(lib, n), n
    assert sorted(_native.SYMBOLS) == names
    assert lib.nkfb_version() == 100


def test_struct_layouts_match_header():
    from netket_fidelity_b200 import _native as nv
    assert ctypes.sizeof(nv.RbmT) == 40
    assert ctypes.sizeof(nv.OpT) == 8 + 64 + 48 + 8 + 8
    assert ctypes.sizeof(nv.SampleCfgT) == 8 + 16 + 24 + 16 + 8 + 16 + 8


def test_no_cpu_fallback_without_gpu():
    import pytest
    import torch
    from netket_fidelity_b200 import _native as nv
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(nv.NkfbError):
        nv.Handle(0)
