"""ctypes access to oracle/_build/liboracle.so -- the CHECKER.  Only tests/, smoke() and bench.py's
cpu_baseline may import this; nothing under stochasticsim_b200/ does."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_SO = os.path.join(ROOT, "oracle", "_build", "liboracle.so")
_lib = None


def oracle():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle")], check=True)
        L = C.CDLL(_SO)
        L.tnc_oracle_count.restype = C.c_int
        L.tnc_oracle_count.argtypes = [C.c_void_p, C.c_size_t, C.POINTER(C.c_int64)]
        L.tnc_oracle_format.restype = C.c_int
        L.tnc_oracle_format.argtypes = [C.POINTER(C.c_int64), C.c_char_p, C.c_size_t]
        L.glibc_rand_fill.restype = None
        L.glibc_rand_fill.argtypes = [C.c_uint, C.c_uint64, C.c_int64, C.POINTER(C.c_int32)]
        L.glibc_walk.restype = C.c_int64
        L.glibc_walk.argtypes = [C.c_uint, C.c_char_p, C.c_int64, C.POINTER(C.c_int32)]
        _lib = L
    return _lib


def tnc_counts(data: bytes) -> np.ndarray:
    out = np.zeros(64, dtype=np.int64)
    buf = C.create_string_buffer(data, len(data)) if len(data) else C.create_string_buffer(1)
    rc = oracle().tnc_oracle_count(C.addressof(buf), len(data), out.ctypes.data_as(C.POINTER(C.c_int64)))
    if rc != 0:
        raise ValueError(f"tnc oracle rejected the input ({rc})")
    return out


def tnc_text(counts64) -> str:
    c = np.ascontiguousarray(counts64, dtype=np.int64)
    buf = C.create_string_buffer(4096)
    w = oracle().tnc_oracle_format(c.ctypes.data_as(C.POINTER(C.c_int64)), buf, 4096)
    return buf.raw[:w].decode()


def glibc_rand(seed: int, skip: int, n: int) -> np.ndarray:
    out = np.zeros(n, dtype=np.int32)
    oracle().glibc_rand_fill(seed & 0xFFFFFFFF, skip, n, out.ctypes.data_as(C.POINTER(C.c_int32)))
    return out
