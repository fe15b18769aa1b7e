"""
TEST INFRASTRUCTURE — ctypes front-end of the plain-C oracle (oracle/cvmx_oracle.c).

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline leg may import this.
"""

from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libcvmx_oracle.so")


def build(force=False):
    if force or not os.path.exists(_SO):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        for suf, ct in (("f64", C.c_double), ("f32", C.c_float)):
            getattr(_lib, f"orc_pairwise_sum_{suf}").restype = ct
            getattr(_lib, f"orc_pairwise_sum_{suf}").argtypes = [C.c_void_p, C.c_int64]
            getattr(_lib, f"orc_colsum_{suf}").restype = None
            getattr(_lib, f"orc_colsum_{suf}").argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p]
            getattr(_lib, f"orc_fit_{suf}").restype = C.c_int
            getattr(_lib, f"orc_fit_{suf}").argtypes = (
                [C.c_void_p] * 3 + [C.c_int64] * 3 + [C.c_uint32] + [C.c_void_p] * 6 + [C.c_void_p, C.c_void_p]
            )
            getattr(_lib, f"orc_fold_{suf}").restype = C.c_int
            getattr(_lib, f"orc_fold_{suf}").argtypes = (
                [C.c_void_p] * 3 + [C.c_int64] * 3 + [C.c_uint32, C.c_int64, ct] + [C.c_void_p] * 6
                + [ct, C.c_int64, C.c_void_p, C.c_int64, C.c_uint32] + [C.c_void_p] * 5
            )
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def pairwise_sum(a):
    a = np.ascontiguousarray(a)
    suf = "f64" if a.dtype == np.float64 else "f32"
    return a.dtype.type(getattr(lib(), f"orc_pairwise_sum_{suf}")(_p(a), a.size))


def colsum(A):
    A = np.ascontiguousarray(A)
    suf = "f64" if A.dtype == np.float64 else "f32"
    out = np.zeros((1, A.shape[1]), dtype=A.dtype)
    getattr(lib(), f"orc_colsum_{suf}")(_p(A), A.shape[0], A.shape[1], _p(out))
    return out


class COracle:
    """fit + per-fold evaluation through the C restatement; returns the *full* set of
    statistics the flags allow (mean/std rows), plus sum_w_train, nnz_train and status."""

    def __init__(self, center_X=True, center_Y=True, scale_X=True, scale_Y=True, ddof=1, dtype=np.float64):
        self.flags = int(center_X) | int(center_Y) << 1 | int(scale_X) << 2 | int(scale_Y) << 3
        self.ddof = int(ddof)
        self.dtype = np.dtype(dtype)
        assert self.dtype in (np.dtype(np.float64), np.dtype(np.float32))
        self.suf = "f64" if self.dtype == np.float64 else "f32"
        self.resolution = float(np.finfo(self.dtype).resolution * 10)

    def fit(self, X, Y=None, weights=None):
        dt = self.dtype
        X = np.ascontiguousarray(np.asarray(X, dtype=dt))
        self.X = X.reshape(-1, 1) if X.ndim == 1 else X
        self.N, self.K = self.X.shape
        self.Y = None
        self.M = 0
        if Y is not None:
            Y = np.ascontiguousarray(np.asarray(Y, dtype=dt))
            self.Y = Y.reshape(-1, 1) if Y.ndim == 1 else Y
            self.M = self.Y.shape[1]
        self.w = None if weights is None else np.ascontiguousarray(np.asarray(weights, dtype=dt).reshape(-1))
        K, M = self.K, self.M
        self.XTX = np.zeros((K, K), dt)
        self.XTY = np.zeros((K, M), dt) if self.Y is not None else None
        self.sum_X, self.sum_sq_X = np.zeros((1, K), dt), np.zeros((1, K), dt)
        self.sum_Y, self.sum_sq_Y = np.zeros((1, max(M, 1)), dt), np.zeros((1, max(M, 1)), dt)
        sw = np.zeros(1, dt)
        nnz = np.zeros(1, np.int64)
        rc = getattr(lib(), f"orc_fit_{self.suf}")(
            _p(self.X), _p(self.Y), _p(self.w), self.N, K, M, self.flags,
            _p(self.XTX), _p(self.XTY), _p(self.sum_X), _p(self.sum_Y), _p(self.sum_sq_X), _p(self.sum_sq_Y),
            _p(sw), _p(nnz),
        )
        if rc == 1:
            raise ValueError("Weights must be non-negative.")
        self.sum_w, self.nnz_w = sw[0], int(nnz[0])

    def fold(self, val, want_XTX=True, want_XTY=True):
        dt = self.dtype
        K, M = self.K, self.M
        val = np.ascontiguousarray(np.asarray(val, dtype=np.int64))
        XTX = np.zeros((K, K), dt) if want_XTX else None
        XTY = np.zeros((K, M), dt) if (want_XTY and self.Y is not None) else None
        stats = np.zeros(2 * K + 2 * max(M, 1), dt)
        scal = np.zeros(2, dt)
        status = np.zeros(1, np.int32)
        ct = C.c_double if dt == np.float64 else C.c_float
        rc = getattr(lib(), f"orc_fold_{self.suf}")(
            _p(self.X), _p(self.Y), _p(self.w), self.N, K, M, self.flags, self.ddof, ct(self.resolution),
            _p(self.XTX), _p(self.XTY), _p(self.sum_X), _p(self.sum_Y), _p(self.sum_sq_X), _p(self.sum_sq_Y),
            ct(self.sum_w), self.nnz_w, _p(val), val.size, int(want_XTX) | int(want_XTY) << 1,
            _p(XTX), _p(XTY), _p(stats), _p(scal), _p(status),
        )
        if rc == 3:
            raise IndexError("validation index out of range")
        return dict(
            XTX=XTX, XTY=XTY,
            X_mean=stats[0:K].reshape(1, K), X_std=stats[K:2 * K].reshape(1, K),
            Y_mean=stats[2 * K:2 * K + M].reshape(1, M), Y_std=stats[2 * K + M:2 * K + 2 * M].reshape(1, M),
            sum_w_train=scal[0], nnz_train=scal[1], status=int(status[0]),
        )
