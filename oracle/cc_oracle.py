"""TEST INFRASTRUCTURE — ctypes wrapper of oracle/cc_oracle.c (built with gcc via oracle/Makefile)."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libccoracle.so")
_lib = None


def build():
    src = os.path.join(_HERE, "cc_oracle.c")
    if not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "_build/libccoracle.so"])
    return _SO


def _load():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.cc_oracle_label.restype = C.c_int
        _lib.cc_oracle_fill_holes.restype = C.c_int
    return _lib


def connected_components(mask_u8):
    """uint8 ndarray [N,1,H,W] -> (labels, counts) int32 ndarrays of the same shape."""
    m = np.ascontiguousarray(mask_u8, dtype=np.uint8)
    N, _, H, W = m.shape
    labels = np.zeros(m.shape, np.int32)
    counts = np.zeros(m.shape, np.int32)
    rc = _load().cc_oracle_label(m.ctypes.data_as(C.c_void_p), labels.ctypes.data_as(C.c_void_p),
                                 counts.ctypes.data_as(C.c_void_p), N, H, W)
    assert rc == 0
    return labels, counts


def fill_holes(scores, max_area):
    """float32 ndarray [N,H,W] (or [N,1,H,W]) -> copy with small holes filled (misc.py:365-393)."""
    s = np.array(scores, dtype=np.float32, copy=True, order="C")
    H, W = s.shape[-2:]
    N = s.size // (H * W)
    rc = _load().cc_oracle_fill_holes(s.ctypes.data_as(C.c_void_p), N, H, W, int(max_area))
    assert rc == 0
    return s
