"""The C-ABI library: builds for sm_100a, loads, and exports every symbol include/dcsb200.h
declares.  No compute calls here (no GPU in the build container)."""
import ctypes
import os
import re
import pytest
from conftest import ROOT


def test_library_exports_every_declared_symbol(built):
    from dcsexplorer_b200 import _capi
    L = _capi.lib()
    hdr = open(os.path.join(ROOT, "include", "dcsb200.h")).read()
    declared = set(re.findall(r"\b(dcsb_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(L, name), name
    assert set(_capi.SYMBOLS) == declared


def test_sass_is_sm100(built):
    import shutil, subprocess
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    out = subprocess.run([cuobjdump, "-lelf", os.path.join(ROOT, "dcsexplorer_b200", "libdcsb200.so")],
                         capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_gain_helpers_match_oracle(built):
    from dcsexplorer_b200 import _capi
    from oracle import orc
    L, O = _capi.lib(), orc.lib()
    for v in range(-2, 300, 7):
        assert L.dcsb_master_multiplier(v) == O.dcso_master_multiplier(v)
    for s in range(-9000, 9000, 37):
        for osv in (0x9301, 0x9400):
            assert L.dcsb_level_multiplier(s, osv, 0xFF, 0) == O.dcso_level_multiplier(s, osv, 0xFF, 0)
            assert L.dcsb_level_multiplier(s, osv, 0x40, 1) == O.dcso_level_multiplier(s, osv, 0x40, 1)


def test_no_cpu_fallback(built):
    """Without a CUDA device the product must fail loudly, not decode on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import dcsexplorer_b200 as dx
    with pytest.raises(dx.DcsbError):
        dx.Context(0)
