"""TEST INFRASTRUCTURE ONLY — ctypes wrapper over oracle/_build/libnms_oracle.so (c/nms_oracle.c)."""
import ctypes
import os

import numpy as np

_lib = None


def lib():
    global _lib
    if _lib is None:
        from . import build
        _lib = ctypes.CDLL(build.build_c())
        _lib.oracle_cpu_nms.restype = ctypes.c_int
        _lib.oracle_cpu_nms.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_float, ctypes.c_int, ctypes.c_void_p]
        _lib.oracle_cpu_soft_nms.restype = ctypes.c_int
        _lib.oracle_cpu_soft_nms.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_float, ctypes.c_float,
                                             ctypes.c_float, ctypes.c_uint]
    return _lib


def cpu_nms(dets, thresh, suppress_on_equal=True):
    dets = np.ascontiguousarray(dets, dtype=np.float32)
    n = dets.shape[0]
    if n == 0:
        return []
    keep = np.empty(n, dtype=np.int32)
    k = lib().oracle_cpu_nms(dets.ctypes.data, n, float(thresh), int(suppress_on_equal), keep.ctypes.data)
    return keep[:k].tolist()


def cpu_soft_nms(boxes, sigma=0.5, Nt=0.3, threshold=0.001, method=0):
    """Returns (boxes_out[N,5], N) on a copy."""
    b = np.array(boxes, dtype=np.float32, copy=True, order='C')
    n = lib().oracle_cpu_soft_nms(b.ctypes.data, b.shape[0], sigma, Nt, threshold, method)
    return b[:n].copy(), n
