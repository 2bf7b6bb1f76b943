"""ctypes wrapper over oracle/liblm_oracle.so -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module (see the header of oracle/lm_oracle.cpp).  The product path never does.
"""
import ctypes
import os
import subprocess

import numpy as np

# libgomp's workers spin after a parallel region by default and starve the caller's serial code
os.environ.setdefault("OMP_WAIT_POLICY", "passive")

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liblm_oracle.so")


class MatchRec(ctypes.Structure):
    _fields_ = [("x", ctypes.c_int32), ("y", ctypes.c_int32), ("similarity", ctypes.c_float),
                ("class_idx", ctypes.c_int32), ("template_id", ctypes.c_int32)]


REC_DTYPE = np.dtype([("x", "<i4"), ("y", "<i4"), ("similarity", "<f4"), ("class_idx", "<i4"), ("template_id", "<i4")])


def build(force=False):
    src = os.path.join(_HERE, "lm_oracle.cpp")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "liblm_oracle.so"], stdout=subprocess.DEVNULL)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        L = ctypes.CDLL(_SO)
        u8p = ctypes.POINTER(ctypes.c_uint8)
        i32p = ctypes.POINTER(ctypes.c_int32)
        L.lmo_spread.argtypes = [u8p, u8p, ctypes.c_int, ctypes.c_int, ctypes.c_int]
        L.lmo_response_maps.argtypes = [u8p, u8p, ctypes.c_int, ctypes.c_int]
        L.lmo_linearize.argtypes = [u8p, u8p, ctypes.c_int, ctypes.c_int, ctypes.c_int]
        L.lmo_similarity_lut.argtypes = [u8p]
        L.lmo_linear_memories.argtypes = [u8p, ctypes.c_int, ctypes.c_int, ctypes.c_int, u8p]
        L.lmo_linear_memories.restype = ctypes.c_int
        L.lmo_match.argtypes = [ctypes.c_int, ctypes.c_int, i32p, i32p, i32p, ctypes.POINTER(u8p),
                                ctypes.c_int, i32p, i32p, i32p, ctypes.c_float, ctypes.c_int,
                                ctypes.POINTER(MatchRec), ctypes.c_long, ctypes.POINTER(ctypes.c_double)]
        L.lmo_match.restype = ctypes.c_long
        L.lmo_match_presort.argtypes = L.lmo_match.argtypes
        L.lmo_match_presort.restype = ctypes.c_long
        L.lmo_coarse_map.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.POINTER(u8p),
                                     i32p, i32p, ctypes.POINTER(ctypes.c_uint16)]
        L.lmo_coarse_map.restype = ctypes.c_long
        L.lmo_max_threads.restype = ctypes.c_int
        _lib = L
    return _lib


def _u8(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8))


def _i32(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_int32))


def similarity_lut():
    out = np.zeros(256, np.uint8)
    lib().lmo_similarity_lut(_u8(out))
    return out


def spread(q, T):
    q = np.ascontiguousarray(q, np.uint8)
    out = np.empty_like(q)
    lib().lmo_spread(_u8(q), _u8(out), q.shape[0], q.shape[1], T)
    return out


def response_maps(sp):
    sp = np.ascontiguousarray(sp, np.uint8)
    out = np.empty((8,) + sp.shape, np.uint8)
    lib().lmo_response_maps(_u8(sp), _u8(out), sp.shape[0], sp.shape[1])
    return out


def linear_memories(q, T):
    """quantized u8 HxW -> [8, T*T, (W/T)*(H/T)] u8."""
    q = np.ascontiguousarray(q, np.uint8)
    H, W = q.shape
    out = np.empty((8, T * T, (W // T) * (H // T)), np.uint8)
    rc = lib().lmo_linear_memories(_u8(q), H, W, T, _u8(out))
    if rc != 0:
        raise RuntimeError("size assertion (rows%T, cols%T, rows*cols%16)")
    return out


def match(quantized, T, packed, threshold, n_threads=1, cap=None, want_stats=False, presort=False):
    """quantized: list over levels of list over modalities of u8 HxW; packed: TemplateBank.pack() dict.

    Returns a structured array (REC_DTYPE) in the reference's final order (+ stats dict)."""
    L = len(quantized)
    M = len(quantized[0])
    qs = [np.ascontiguousarray(quantized[l][m], np.uint8) for l in range(L) for m in range(M)]
    rows = np.asarray([quantized[l][0].shape[0] for l in range(L)], np.int32)
    cols = np.asarray([quantized[l][0].shape[1] for l in range(L)], np.int32)
    Ts = np.asarray(T, np.int32)
    ptrs = (ctypes.POINTER(ctypes.c_uint8) * len(qs))(*[_u8(a) for a in qs])
    cb, tm, ft = packed["class_begin"], packed["tmeta"], packed["feats"]
    if cap is None:
        cap = 1 << 16
    stats = (ctypes.c_double * 8)()
    while True:
        out = np.zeros(cap, REC_DTYPE)
        fn = lib().lmo_match_presort if presort else lib().lmo_match
        n = fn(L, M, _i32(Ts), _i32(rows), _i32(cols), ptrs, len(cb) - 1, _i32(cb), _i32(tm), _i32(ft),
                            ctypes.c_float(threshold), int(n_threads),
                            out.ctypes.data_as(ctypes.POINTER(MatchRec)), cap, stats)
        if n < 0:
            raise RuntimeError("oracle: reference would raise (code %d)" % n)
        if n <= cap:
            break
        cap = int(n)
    res = out[:n].copy()
    if want_stats:
        keys = ["coarse_byte_adds", "refine_byte_adds", "coarse_candidates", "pre_unique", "t_linmem_us", "t_match_us",
                "t_sort_us", "threads"]
        return res, dict(zip(keys, [float(v) for v in stats]))
    return res


def coarse_map(quantized_low, T, tmeta_low, feats):
    M = len(quantized_low)
    qs = [np.ascontiguousarray(q, np.uint8) for q in quantized_low]
    H, W = qs[0].shape
    ptrs = (ctypes.POINTER(ctypes.c_uint8) * M)(*[_u8(a) for a in qs])
    tm = np.ascontiguousarray(tmeta_low, np.int32)
    ft = np.ascontiguousarray(feats, np.int32)
    out = np.zeros((H // T, W // T), np.uint16)
    P = lib().lmo_coarse_map(M, T, H, W, ptrs, _i32(tm), _i32(ft), out.ctypes.data_as(ctypes.POINTER(ctypes.c_uint16)))
    if H % T or W % T or (H * W) % 16:
        raise RuntimeError("size assertion (rows%T, cols%T, rows*cols%16)")
    return out, int(P)


def max_threads():
    return int(lib().lmo_max_threads())
