"""ctypes binding of libyb200.so -- the only way Python reaches the CUDA kernels.

There is no fallback: if the library is missing or a symbol is absent, importing the hot path fails loudly.
Every wrapper takes torch tensors only to read `data_ptr()` / shapes and the current CUDA stream; the C ABI
itself (include/yb200.h) sees plain pointers and integers.
"""
import ctypes
import os
import re

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libyb200.so")
HEADER_PATH = os.path.join(_HERE, "..", "include", "yb200.h")

c_int = ctypes.c_int
c_i64 = ctypes.c_int64
c_void_p = ctypes.c_void_p
c_float = ctypes.c_float


class Yb200Error(RuntimeError):
    pass


ERR_INVALID, ERR_UNSUPPORTED, ERR_CUDA = -1, -2, -3  # yb200_status (include/yb200.h)


class Act(ctypes.Structure):
    """mirror of `yb200_act` (include/yb200.h)"""

    _fields_ = [
        ("ptr", c_void_p),
        ("n", ctypes.c_int32),
        ("h", ctypes.c_int32),
        ("w", ctypes.c_int32),
        ("c", ctypes.c_int32),
        ("c_pitch", ctypes.c_int32),
        ("c_off", ctypes.c_int32),
    ]


class BnBwdSeg(ctypes.Structure):
    """mirror of `yb200_bnbwd_seg` (include/yb200.h)"""

    _fields_ = [
        ("z", Act),
        ("dx_c_begin", ctypes.c_int32),
        ("scale", c_void_p),
        ("shift", c_void_p),
        ("sum_du", c_void_p),
        ("sum_duz", c_void_p),
    ]


class PackDesc(ctypes.Structure):
    """mirror of `yb200_pack_desc` (include/yb200.h)"""

    _fields_ = [("w_oihw", c_void_p), ("w_fwd", c_void_p), ("w_dgrad", c_void_p), ("cout", ctypes.c_int32), ("cin", ctypes.c_int32),
                ("ksize", ctypes.c_int32), ("cout_pad", ctypes.c_int32), ("cin_pad", ctypes.c_int32), ("reserved", ctypes.c_int32)]


def declared_symbols(header_path=HEADER_PATH):
    """Every function the public header declares (used by the CPU test that checks the exports)."""
    with open(header_path) as fh:
        src = fh.read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(yb200_[a-z0-9_]+)\s*\(", src)))


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise Yb200Error(
                f"{LIB_PATH} not found: build it with `python -m yolov7_d2_b200.build` (needs nvcc). "
                "The hot path has no CPU / PyTorch fallback."
            )
        _lib = ctypes.CDLL(LIB_PATH)
        _lib.yb200_last_error.restype = ctypes.c_char_p
        _lib.yb200_conv2d_wgrad_workspace.restype = c_i64
        _lib.yb200_simota_workspace.restype = c_i64
        _lib.yb200_nms_workspace.restype = c_i64
        _lib.yb200_grad_norm_workspace.restype = c_i64
        for _n in ("yb200_dwconv7_wgrad_workspace", "yb200_layernorm_bwd_workspace", "yb200_colsum_workspace", "yb200_attention_bwd_workspace"):
            getattr(_lib, _n).restype = c_i64
        for name in declared_symbols():
            if not hasattr(_lib, name):
                raise Yb200Error(f"libyb200.so does not export {name} declared in include/yb200.h")
    return _lib


def check(rc, what):
    if rc != 0:
        raise Yb200Error(f"{what} failed ({rc}): {lib().yb200_last_error().decode()}")


def stream_ptr():
    return c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t):
    return c_void_p(t.data_ptr()) if t is not None else c_void_p(0)


def act(t, c_off=0, c=None):
    """View of an NHWC bf16 tensor [N,H,W,Cpitch] (contiguous) restricted to channels [c_off, c_off+c)."""
    assert t.dtype in (torch.bfloat16, torch.float16) and t.dim() == 4 and t.is_contiguous(), (t.dtype, t.shape, t.stride())
    n, h, w, cp = t.shape
    return Act(t.data_ptr(), n, h, w, cp - c_off if c is None else c, cp, c_off)
