"""Builds/loads the TEST-ONLY host compilation of gubernator_b200/csrc/bucket_math.cuh (tests/host_math_harness.cpp)."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SO = os.path.join(HERE, "libhostmath_test.so")
SRC = [os.path.join(HERE, "host_math_harness.cpp"), os.path.join(ROOT, "gubernator_b200", "csrc", "bucket_math.cuh"),
       os.path.join(ROOT, "include", "gubernator_b200.h")]

CLOCK_DTYPE = np.dtype([("now_ms", "<i8"), ("greg_expire", "<i8", (6,)), ("greg_duration", "<i8", (6,))])
BUCKET_DTYPE = np.dtype([("key", "<u8"), ("tag", "<u8"), ("limit", "<i8"), ("duration", "<i8"), ("rem", "<u8"),
                         ("stamp", "<i8"), ("burst", "<i8"), ("expire", "<i8"), ("flags", "<u4"), ("_pad", "<u4")])
F_LEAKY, F_OVER, F_LIVE = 1, 2, 4

_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(SO) or any(os.path.getmtime(s) > os.path.getmtime(SO) for s in SRC):
            subprocess.check_call(["g++", "-O2", "-g", "-std=c++17", "-fPIC", "-shared", "-msse2", "-ffp-contract=off",
                                   "-x", "c++", SRC[0], "-o", SO])
        L = C.CDLL(SO)
        L.hm_table_new.restype = C.c_void_p
        L.hm_table_free.argtypes = [C.c_void_p]
        L.hm_apply_seq.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p]
        L.hm_plan_check.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p]
        L.hm_rank_check.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32]
        L.hm_sizeof_bucket.restype = C.c_size_t
        assert L.hm_sizeof_bucket() == BUCKET_DTYPE.itemsize
        _lib = L
    return _lib


class HostTable:
    def __init__(self):
        self._t = lib().hm_table_new()

    def apply_seq(self, reqs, clk):
        import oracle_py as O
        out = np.zeros(len(reqs), dtype=O.HRESP_DTYPE)
        ctr = np.zeros(3, dtype=np.uint64)
        lib().hm_apply_seq(self._t, reqs.ctypes.data, len(reqs), clk.ctypes.data, out.ctypes.data, ctr.ctypes.data)
        return out, dict(over_limit=int(ctr[0]), cache_hit=int(ctr[1]), cache_miss=int(ctr[2]))

    def __del__(self):
        try:
            lib().hm_table_free(self._t)
        except Exception:
            pass


def plan_check(bucket, req, m, clk, cap=16):
    npieces = C.c_uint32(0)
    covered = C.c_uint32(0)
    rc = lib().hm_plan_check(bucket.ctypes.data, req.ctypes.data, m, clk.ctypes.data, cap, C.byref(npieces), C.byref(covered))
    return rc, npieces.value, covered.value


def rank_check(bucket, req, m, clk, stride=1):
    return lib().hm_rank_check(bucket.ctypes.data, req.ctypes.data, m, clk.ctypes.data, stride)
