"""ctypes loader for oracle/_build/liboracle.so (kmv_oracle.c).  TEST INFRASTRUCTURE / CPU BASELINE ONLY."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_build", "liboracle.so")
_lib = None


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE])


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        lib = ctypes.CDLL(LIB_PATH)
        lib.oracle_max_threads.restype = ctypes.c_int
        for name in ("oracle_kmv_f32", "oracle_kmv_f64"):
            fn = getattr(lib, name)
            fn.restype = ctypes.c_int
            fn.argtypes = [ctypes.c_void_p, ctypes.c_long, ctypes.c_void_p, ctypes.c_long, ctypes.c_int, ctypes.c_int,
                           ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int]
        _lib = lib
    return _lib


def max_threads():
    return int(load().oracle_max_threads())


def kmv(Z1, Z2, c, J, K, V, dtype=np.float32, threads=0):
    """out = K(Z1, Z2) @ V on the host cores (natural, un-scaled coordinates)."""
    lib = load()
    Z1 = np.ascontiguousarray(Z1, dtype=dtype)
    Z2 = np.ascontiguousarray(Z2, dtype=dtype)
    V = np.ascontiguousarray(V, dtype=dtype)
    c = np.ascontiguousarray(np.broadcast_to(np.asarray(c, dtype=dtype), (J,)))
    m, n, t = Z1.shape[0], Z2.shape[0], V.shape[1]
    out = np.empty((m, t), dtype=dtype)
    fn = lib.oracle_kmv_f32 if dtype == np.float32 else lib.oracle_kmv_f64
    rc = fn(Z1.ctypes.data, m, Z2.ctypes.data, n, J, K, c.ctypes.data, V.ctypes.data, t, out.ctypes.data, int(threads))
    if rc != 0:
        raise ValueError("oracle_kmv: bad arguments")
    return out
