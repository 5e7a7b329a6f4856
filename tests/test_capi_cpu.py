"""CPU checks of the C-ABI boundary: the library builds, loads, and exports every symbol include/yb200.h declares."""
import ctypes
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from yolov7_d2_b200 import build, capi

    if build.find_nvcc() is None and not os.path.exists(capi.LIB_PATH):
        pytest.skip("no nvcc and no prebuilt libyb200.so")
    build.build()
    names = capi.declared_symbols()
    assert len(names) >= 20 and "yb200_conv2d_fwd" in names and "yb200_postprocess_nms" in names
    lib = ctypes.CDLL(capi.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), n
    assert lib.yb200_version() == 100


def test_argument_validation_without_gpu():
    """invalid arguments are rejected on the host before any CUDA call (no GPU needed)"""
    from yolov7_d2_b200 import capi

    L = capi.lib()
    a = capi.Act(0, 1, 8, 8, 16, 16, 0)
    assert L.yb200_conv2d_fwd(ctypes.byref(a), None, ctypes.byref(a), 3, 1, None, None, None) == -1
    assert b"null" in L.yb200_last_error()
    assert L.yb200_simota_workspace(0, 8400) < 0 and L.yb200_nms_workspace(4, 70000) < 0
    assert L.yb200_simota_workspace(64, 8400) > 0 and L.yb200_nms_workspace(64, 8400) > 0


def test_sass_contains_blackwell_tensor_and_tma_instructions():
    import shutil
    import subprocess

    from yolov7_d2_b200 import capi

    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not available")
    sass = subprocess.run(["cuobjdump", "-sass", capi.LIB_PATH], capture_output=True, text=True).stdout
    assert "UTCHMMA" in sass and "UTMALDG" in sass and "LDTM" in sass, "tcgen05 / TMA instructions missing from the sm_100a build"
