"""ctypes binding of oracle/_ref/libdm_ref.so — the REFERENCE's own estd2::DenseMatcher compiled from
/root/reference by oracle/ref/Makefile (driver: oracle/ref/dm_ref_shim.cpp).  TEST INFRASTRUCTURE ONLY."""
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_PATH = os.path.join(_HERE, "_ref", "libdm_ref.so")
_LIB = None


def available():
    if not os.path.exists(_PATH) and os.path.isdir("/root/reference/covins_backend"):
        subprocess.call(["make", "-C", os.path.join(_HERE, "ref"), "-s"])
    return os.path.exists(_PATH)


def lib():
    global _LIB
    if _LIB is None:
        if not available():
            raise RuntimeError("oracle/_ref/libdm_ref.so is not built (needs /root/reference; make -C oracle/ref)")
        _LIB = C.CDLL(_PATH)
        _LIB.dm_ref_match.restype = C.c_int
    return _LIB


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t)) if a is not None else None


def dense_match(A, skipA, B, skipB, thr=50.0, num_threads=1, num_best=4):
    """estd2::DenseMatcher(num_threads, num_best, false).match(algo) with the byte-array MatchingAlgorithm:
    → (idxA, idxB, distance) ordered by B index.  num_threads = 1 is the canonical deterministic order."""
    A = np.ascontiguousarray(A, np.uint8); B = np.ascontiguousarray(B, np.uint8)
    sA = np.ascontiguousarray(skipA, np.uint8) if skipA is not None else None
    sB = np.ascontiguousarray(skipB, np.uint8) if skipB is not None else None
    nA, nB = len(A), len(B)
    oA = np.empty(max(nB, 1), np.int32); oB = np.empty(max(nB, 1), np.int32); oD = np.empty(max(nB, 1), np.float32)
    n = lib().dm_ref_match(_p(A, C.c_uint8), _p(sA, C.c_uint8), nA, _p(B, C.c_uint8), _p(sB, C.c_uint8), nB,
                           C.c_float(thr), num_threads, num_best, _p(oA, C.c_int32), _p(oB, C.c_int32), _p(oD, C.c_float))
    return oA[:n].copy(), oB[:n].copy(), oD[:n].copy()
