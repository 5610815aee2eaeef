"""ORACLE - TEST INFRASTRUCTURE ONLY.  ctypes front-end of oracle/assoc_oracle.cpp
(the CPU restatement of dapalib.extract/connect)."""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "liboracle_assoc.so")
NJ, NL, MAXP = 15, 14, 127
_lib = None


def build(force=False):
    src = os.path.join(HERE, "assoc_oracle.cpp")
    if force or not os.path.exists(SO) or os.path.getmtime(SO) < os.path.getmtime(src):
        subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-o", SO, src, "-lm"])
    return SO


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(SO)
        fp = ctypes.POINTER(ctypes.c_float)
        _lib.oracle_nms.argtypes = [fp, ctypes.c_int, ctypes.c_int, ctypes.c_float, fp]
        _lib.oracle_paf.argtypes = [fp, ctypes.c_int, ctypes.c_int, fp, fp]
        _lib.oracle_group.argtypes = [fp, fp, fp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, fp]
        _lib.oracle_group.restype = ctypes.c_int
        _lib.oracle_connect.argtypes = [fp, fp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, fp, fp, fp]
        _lib.oracle_connect.restype = ctypes.c_int
        _lib.oracle_depth_order.argtypes = [fp, ctypes.c_int, ctypes.POINTER(ctypes.c_int)]
    return _lib


def _p(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))


def extract(hms):
    """hms float32 [43,h,w] -> (peaks [15,128,3] with unused slots zeroed, scores [14,127,127])."""
    hms = np.ascontiguousarray(hms, np.float32)
    _, h, w = hms.shape
    peaks = np.zeros((NJ, MAXP + 1, 3), np.float32)
    scores = np.empty((NL, MAXP, MAXP), np.float32)
    lib().oracle_nms(_p(hms), h, w, 0.2, _p(peaks))
    lib().oracle_paf(_p(hms), h, w, _p(peaks), _p(scores))
    return peaks, scores


def connect(hms, rdepth, root_idx=2, dist_flag=True, return_all=False):
    """dapalib.connect restated: returns float32 [P,15,4] (P may be 0)."""
    hms = np.ascontiguousarray(hms, np.float32)
    rdepth = np.ascontiguousarray(rdepth, np.float32)
    _, h, w = hms.shape
    peaks = np.zeros((NJ, MAXP + 1, 3), np.float32)
    scores = np.empty((NL, MAXP, MAXP), np.float32)
    bodies = np.zeros((MAXP, NJ, 4), np.float32)
    P = lib().oracle_connect(_p(hms), _p(rdepth), h, w, root_idx, int(dist_flag), _p(peaks), _p(scores), _p(bodies))
    if return_all:
        return bodies[:P].copy(), peaks, scores
    return bodies[:P].copy()


def depth_order(depth):
    """association.cpp:144 alone: indices of the (unstable, std::sort) ascending depth sort."""
    d = np.ascontiguousarray(depth, np.float32)
    order = np.zeros(len(d), np.int32)
    lib().oracle_depth_order(_p(d), len(d), order.ctypes.data_as(ctypes.POINTER(ctypes.c_int)))
    return order
