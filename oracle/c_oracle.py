"""ctypes access to the plain-C oracle (``hypergeom_oracle.c``).

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).
"""
import ctypes

import numpy as np

from . import build as _build

_lib = None


def _load():
    global _lib
    if _lib is None:
        lib = ctypes.CDLL(_build.build())
        i64p = ctypes.POINTER(ctypes.c_int64)
        f64p = ctypes.POINTER(ctypes.c_double)
        for name in ("spade_oracle_hypergeom", "spade_oracle_pvalues"):
            fn = getattr(lib, name)
            fn.argtypes = [i64p, ctypes.c_int64, i64p, i64p, ctypes.c_int64, f64p, f64p]
            fn.restype = None
        _lib = lib
    return _lib


def _call(name, k, M, n, N):
    lib = _load()
    k = np.ascontiguousarray(np.broadcast_arrays(k, n, N)[0], dtype=np.int64).ravel()
    shape = np.broadcast_shapes(np.shape(k), np.shape(n), np.shape(N))
    kk, nn, NN = (np.ascontiguousarray(np.broadcast_to(np.asarray(a, dtype=np.int64), shape)).ravel()
                  for a in (k, n, N))
    out_a = np.empty(kk.shape[0], dtype=np.float64)
    out_b = np.empty(kk.shape[0], dtype=np.float64)
    i64p = ctypes.POINTER(ctypes.c_int64)
    f64p = ctypes.POINTER(ctypes.c_double)
    getattr(lib, name)(kk.ctypes.data_as(i64p), int(M), nn.ctypes.data_as(i64p),
                       NN.ctypes.data_as(i64p), kk.shape[0],
                       out_a.ctypes.data_as(f64p), out_b.ctypes.data_as(f64p))
    return out_a.reshape(shape), out_b.reshape(shape)


def hypergeom_logcdf_logsf(k, M, n, N):
    """scipy semantics: returns ``(logcdf, logsf)`` arrays."""
    return _call("spade_oracle_hypergeom", k, M, n, N)


def pvalues(k, M, n, N):
    """reference semantics (nan guard): returns ``(up, down)`` = (logcdf, logsf)."""
    return _call("spade_oracle_pvalues", k, M, n, N)
