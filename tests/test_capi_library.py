"""The C-ABI library builds, loads, and exports exactly what include/sot_b200.h declares.
No compute calls here (no GPU needed)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "sot_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(sot_[a-z_0-9]+)\s*\(", text)))


def test_header_and_binding_agree():
    from sot_b200 import _capi
    assert sorted(_capi.SIGNATURES) == _declared()


def test_library_loads_and_exports_every_symbol():
    from sot_b200 import _capi
    lib = _capi.load()
    assert os.path.samefile(_capi.library_path(), os.path.join(ROOT, "1d-spectral-optimal-transport_b200", "_lib",
                                                               "libsot_b200.so"))
    for name in _declared():
        assert hasattr(lib, name), name
    assert lib.sot_abi_version() == _capi.ABI_VERSION
    assert _capi.launch_count() == 0 or _capi.launch_count() > 0
    # paper shapes fit, with and without gradients, shared and per-frame supports
    assert _capi.max_bins(True, True) >= 4097 and _capi.max_bins(True, False) >= 2049


def test_argument_validation_happens_before_any_launch():
    from sot_b200 import _capi
    lib = _capi.load()
    before = _capi.launch_count()
    prob = _capi.SotProblem(4, 9, 9, 0, 0, 0, 0, 0, 0, 2.0, 0)
    assert lib.sot_forward_device(ctypes.byref(prob), None, None) == -1  # NULL pointers
    assert b"NULL" in lib.sot_last_error()
    prob = _capi.SotProblem(4, 9, 9, 16, 16, 16, 16, 0, 0, 0.5, 0)
    assert lib.sot_forward_device(ctypes.byref(prob), ctypes.c_void_p(16), None) == -3  # p < 1
    assert b"only valid for p>=1" in lib.sot_last_error()
    prob = _capi.SotProblem(4, 9, 9, 16, 16, 16, 16, 0, 0, 2.0, 64)
    assert lib.sot_forward_device(ctypes.byref(prob), ctypes.c_void_p(16), None) == -1  # unknown flag
    prob = _capi.SotProblem(4, 60000, 60000, 16, 16, 16, 16, 0, 0, 2.0, 0)
    assert lib.sot_forward_device(ctypes.byref(prob), ctypes.c_void_p(16), None) == -2  # does not fit
    assert lib.sot_set_tuning(48, 7, 0) == -1
    assert lib.sot_set_tuning(0, 0, 0) == 0
    assert _capi.launch_count() == before


def test_sass_contains_tma_bulk_copies():
    """The hot kernels move rows with cp.async.bulk (SASS UBLKCP), not with per-thread loads."""
    import shutil
    import subprocess
    from sot_b200 import _capi
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not available")
    sass = subprocess.run(["cuobjdump", "-sass", _capi.library_path()], capture_output=True, text=True).stdout
    assert "UBLKCP" in sass and "SYNCS" in sass
