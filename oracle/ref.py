"""ctypes wrapper over oracle/_ref/liblm_ref.so -- the reference's own linemodLevelup.cpp compiled
unmodified against oracle/ref_shim (see oracle/ref_shim/ref_driver.cpp).  TEST INFRASTRUCTURE ONLY.

The library can only be BUILT where /root/reference is mounted (`make -C oracle ref`); the prebuilt
.so travels with the repo snapshot (oracle/_ref/ is git-ignored, not gpurun-ignored)."""
import ctypes
import os
import subprocess

import numpy as np

from . import oracle as _o

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_ref", "liblm_ref.so")
_lib = None


def build():
    if os.path.isdir("/root/reference/linemodLevelup"):
        subprocess.check_call(["make", "-C", _HERE, "ref"], stdout=subprocess.DEVNULL)
    return available()


def available():
    return os.path.exists(_SO)


def lib():
    global _lib
    if _lib is None:
        L = ctypes.CDLL(_SO)
        u8p = ctypes.POINTER(ctypes.c_uint8)
        i32p = ctypes.POINTER(ctypes.c_int32)
        L.lmr_match.argtypes = [ctypes.c_int, ctypes.c_int, i32p, i32p, i32p, ctypes.POINTER(u8p), ctypes.c_int, i32p, i32p,
                                i32p, ctypes.c_float, ctypes.c_int, ctypes.POINTER(_o.MatchRec), ctypes.c_long,
                                ctypes.POINTER(ctypes.c_double)]
        L.lmr_match.restype = ctypes.c_long
        L.lmr_similarity_lut.argtypes = [u8p]
        _lib = L
    return _lib


def similarity_lut():
    out = np.zeros(256, np.uint8)
    lib().lmr_similarity_lut(out.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8)))
    return out


def match(quantized, T, packed, threshold, n_threads=1, cap=None, want_stats=False):
    """Same contract as oracle.match; single-threaded like the reference."""
    L = len(quantized)
    M = len(quantized[0])
    qs = [np.ascontiguousarray(quantized[l][m], np.uint8) for l in range(L) for m in range(M)]
    rows = np.asarray([quantized[l][0].shape[0] for l in range(L)], np.int32)
    cols = np.asarray([quantized[l][0].shape[1] for l in range(L)], np.int32)
    Ts = np.asarray(T, np.int32)
    u8p = ctypes.POINTER(ctypes.c_uint8)
    i32 = lambda a: a.ctypes.data_as(ctypes.POINTER(ctypes.c_int32))
    ptrs = (u8p * len(qs))(*[a.ctypes.data_as(u8p) for a in qs])
    cb, tm, ft = packed["class_begin"], packed["tmeta"], packed["feats"]
    if cap is None:
        cap = 1 << 16
    stats = (ctypes.c_double * 8)()
    while True:
        out = np.zeros(cap, _o.REC_DTYPE)
        n = lib().lmr_match(L, M, i32(Ts), i32(rows), i32(cols), ptrs, len(cb) - 1, i32(cb), i32(tm), i32(ft),
                            ctypes.c_float(threshold), int(n_threads), out.ctypes.data_as(ctypes.POINTER(_o.MatchRec)), cap, stats)
        if n < 0:
            raise RuntimeError("reference raised cv::Exception (CV_Assert)")
        if n <= cap:
            break
        cap = int(n)
    res = out[:n].copy()
    if want_stats:
        return res, {"t_match_us": float(stats[5]), "threads": 1.0}
    return res
