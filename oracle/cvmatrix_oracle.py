"""
TEST INFRASTRUCTURE — NOT PRODUCT CODE.

CPU restatement ("port") of the fold-wise training-matrix path of sm00thix/cvmatrix
(reference v3.2.1).  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import this module;
the shipped package ``cvmatrix_b200`` never does.

Parity status: PINNED.  ``oracle/check_against_reference.py`` runs this restatement
next to the live reference (imported from /root/reference in the build container)
over all 16 flag combinations x {weighted, unweighted} x ddof x dtype and demands
bit-identical outputs; ``tests/golden/make_golden.py`` freezes reference outputs as
fixtures that ``tests/test_oracle.py`` re-checks everywhere (also on the GPU box,
where /root/reference does not exist).

The arithmetic of the reference lives in numpy (>=2.0,<3.0; pyproject.toml:16-18,
uv.lock pins 2.4.6/2.5.0) and its bundled OpenBLAS.  What matters for parity is the
*order* of the reductions, restated explicitly here (``order="explicit"``):

* ``np.sum`` of an (n, 1) / 1-D contiguous array  -> numpy pairwise summation
  (blocks of <=128 with 8 interleaved accumulators; split at n/2 rounded down to a
  multiple of 8)                                    [used at cvmatrix/cvmatrix.py:617, 1225]
* ``np.sum(A, axis=0)`` of a C-contiguous (n, C>=2) -> strictly sequential per column
  in row order; C == 1 degenerates to the pairwise case   [cvmatrix.py:709, 716, 727, 737, 1231-1241]

With ``order="numpy"`` the same quantities are obtained with the same numpy calls the
reference makes (this is the variant timed as the CPU baseline).

Reference map (file:line in /root/reference):
  fit .......................... cvmatrix/cvmatrix.py:1131-1243
  fold weight sum / count ...... cvmatrix/cvmatrix.py:589-630
  fold statistics .............. cvmatrix/cvmatrix.py:632-752, 1012-1129
  fold kernel matrices ......... cvmatrix/cvmatrix.py:898-1010
  which stats each call returns  cvmatrix/cvmatrix.py:563-574, 806-896
  Partitioner .................. cvmatrix/partitioner.py:48-107
"""

from __future__ import annotations

from collections import namedtuple

import numpy as np

ERR_NO_NONZERO = (
    "The number of non-zero weights in the training set must be greater than zero."
)
ERR_DDOF = (
    "The number of non-zero weights in the training set must be greater than `ddof`."
)
ERR_NEG_W = "Weights must be non-negative."
ERR_NOTHING = "At least one of `return_XTX` and `return_XTY` must be True."
ERR_NO_Y = "Response variables `Y` are not provided."

FoldResult = namedtuple(
    "FoldResult", "XTX XTY X_mean X_std Y_mean Y_std sum_w_train nnz_train"
)


# --------------------------------------------------------------------------------------
# Explicit summation orders
# --------------------------------------------------------------------------------------
def pairwise_sum(a):
    """numpy's pairwise summation of a 1-D array, restated (see module docstring).

    Every addition is a single rounded add in ``a.dtype``.
    """
    a = np.ascontiguousarray(a).reshape(-1)
    n = a.shape[0]
    dt = a.dtype.type
    if n == 0:
        return dt(0)
    if n < 8:
        # numpy starts this branch from 0., which leaves every finite sum unchanged
        # except the sign of an all-negative-zero input.
        r = dt(0.0)
        for i in range(0, n):
            r = dt(r + a[i])
        return r
    if n <= 128:
        r = a[0:8].copy()
        stop = n - (n % 8)
        for i in range(8, stop, 8):
            r = r + a[i : i + 8]
        res = dt(
            dt(dt(r[0] + r[1]) + dt(r[2] + r[3])) + dt(dt(r[4] + r[5]) + dt(r[6] + r[7]))
        )
        for i in range(stop, n):
            res = dt(res + a[i])
        return res
    n2 = n // 2
    n2 -= n2 % 8
    return dt(pairwise_sum(a[:n2]) + pairwise_sum(a[n2:]))


def sequential_colsum(A, chunk=4096):
    """Column sums of a 2-D array accumulated strictly in row order (acc = 0; acc += row_i).

    ``np.cumsum`` along axis 0 is sequential by construction, so chaining chunked cumsums
    reproduces the order without a Python loop per row.
    """
    A = np.asarray(A)
    n, c = A.shape
    acc = np.zeros((1, c), dtype=A.dtype)  # numpy seeds add-reductions with +0
    i = 0
    while i < n:
        j = min(n, i + chunk)
        acc = np.cumsum(np.concatenate([acc, A[i:j]], axis=0), axis=0)[-1:].copy()
        i = j
    return acc


def colsum(A, order):
    """``np.sum(A, axis=0, keepdims=True)`` in either restated or native order."""
    if order == "numpy":
        return np.sum(A, axis=0, keepdims=True)
    if A.shape[1] == 1:
        return np.asarray(0.0 + pairwise_sum(A[:, 0]), dtype=A.dtype).reshape(1, 1)
    return sequential_colsum(A)


def total_sum(a, order):
    """``np.sum`` of an (n, 1) array."""
    if order == "numpy":
        return np.sum(a)
    p = pairwise_sum(a.reshape(-1))
    return type(p)(0.0 + p)  # numpy seeds add-reductions with +0 (matters for -0 only)


# --------------------------------------------------------------------------------------
# Partitioner
# --------------------------------------------------------------------------------------
class OraclePartitioner:
    """fold label sequence -> {label: ascending int64 row indices}, first-seen key order."""

    def __init__(self, folds):
        buckets = {}
        for row, label in enumerate(folds):
            if label in buckets:
                buckets[label].append(row)
            else:
                buckets[label] = [row]
        self.folds_dict = {k: np.asarray(v, dtype=int) for k, v in buckets.items()}

    def get_validation_indices(self, fold):
        if fold not in self.folds_dict:
            raise ValueError(f"Fold {fold} not found.")
        return self.folds_dict[fold]


# --------------------------------------------------------------------------------------
# Fold-matrix engine
# --------------------------------------------------------------------------------------
class OracleCVMatrix:
    def __init__(
        self,
        center_X=True,
        center_Y=True,
        scale_X=True,
        scale_Y=True,
        ddof=1,
        dtype=np.float64,
        copy=True,
        order="numpy",
    ):
        assert order in ("numpy", "explicit")
        self.cX, self.cY, self.sX, self.sY = center_X, center_Y, scale_X, scale_Y
        self.ddof = ddof
        self.dtype = dtype.type if isinstance(dtype, np.dtype) else dtype
        self.copy = copy
        self.order = order
        self.resolution = np.finfo(dtype).resolution * 10
        self.X = self.Y = self.w = None

    # -- fit ----------------------------------------------------------------------------
    def _as_matrix(self, a):
        a = np.asarray(a, dtype=self.dtype)
        if self.copy:
            a = a.copy()
        return a.reshape(-1, 1) if a.ndim == 1 else a

    def fit(self, X, Y=None, weights=None):
        cX, cY, sX, sY = self.cX, self.cY, self.sX, self.sY
        self.X = self._as_matrix(X)
        self.N, self.K = self.X.shape
        self.Y = None if Y is None else self._as_matrix(Y)
        self.M = None if Y is None else self.Y.shape[1]
        self.w = None if weights is None else self._as_matrix(weights)
        if self.w is not None and bool(np.any(self.w < 0)):
            raise ValueError(ERR_NEG_W)

        has_Y = self.Y is not None
        if self.w is None:
            self.WX = self.X
            self.WY = self.Y
        else:
            self.WX = self.X * self.w
            self.WY = self.Y * self.w if has_Y and (cX or cY or sY) else None

        self.XTX = self.WX.T @ self.X
        self.XTY = self.WX.T @ self.Y if has_Y else None

        self.sum_w = self.nnz_w = None
        self.sum_X = self.sum_Y = self.sum_sq_X = self.sum_sq_Y = None
        self.sq_X = self.sq_Y = None
        if cX or cY or sX or sY:
            if self.w is None:
                self.sum_w, self.nnz_w = self.N, self.N
            else:
                self.sum_w = total_sum(self.w, self.order)
                self.nnz_w = np.count_nonzero(self.w)
        if cX or cY or sX:
            self.sum_X = colsum(self.WX, self.order)
        if has_Y and (cX or cY or sY):
            self.sum_Y = colsum(self.WY, self.order)
        if sX:
            self.sq_X = self.WX * self.X
            self.sum_sq_X = colsum(self.sq_X, self.order)
        if has_Y and sY:
            self.sq_Y = self.WY * self.Y
            self.sum_sq_Y = colsum(self.sq_Y, self.order)

    # -- one fold -----------------------------------------------------------------------
    def _train_weight_mass(self, val):
        dt = self.dtype
        if self.w is None:
            sw = dt(self.sum_w - val.size)
            return sw, sw
        w_val = self.w[val]
        sw = dt(self.sum_w - total_sum(w_val, self.order))
        nz = dt(self.nnz_w - np.count_nonzero(w_val))
        if nz == 0:
            raise ValueError(ERR_NO_NONZERO)
        return sw, nz

    def _std_row(self, q_train, mean, s_train, sw, div):
        var = (-2 * mean * s_train + sw * (mean * mean) + q_train) / div
        std = np.sqrt(np.maximum(var, 0))
        return np.where(std <= self.resolution, 1, std)

    def _fold_stats(self, val, WXv, WYv, need_Xmean, need_Xstd, need_Ymean, need_Ystd):
        """Returns (X_mean, X_std, Y_mean, Y_std, sw, nz); un-needed entries are None."""
        if not (need_Xmean or need_Xstd or need_Ymean or need_Ystd):
            return None, None, None, None, None, None
        sw, nz = self._train_weight_mass(val)
        Xm = Xs = Ym = Ys = None
        if need_Xmean or need_Xstd:
            sX_train = self.sum_X - colsum(WXv, self.order)
            Xm = sX_train / sw
        if need_Ymean or need_Ystd:
            sY_train = self.sum_Y - colsum(WYv, self.order)
            Ym = sY_train / sw
        if need_Xstd or need_Ystd:
            if nz <= self.ddof:
                raise ValueError(ERR_DDOF)
            div = (nz - self.ddof) * sw / nz
        if need_Xstd:
            qX_train = self.sum_sq_X - colsum(self.sq_X[val], self.order)
            Xs = self._std_row(qX_train, Xm, sX_train, sw, div)
        if need_Ystd:
            qY_train = self.sum_sq_Y - colsum(self.sq_Y[val], self.order)
            Ys = self._std_row(qY_train, Ym, sY_train, sw, div)
        return (
            Xm if need_Xmean else None,
            Xs if need_Xstd else None,
            Ym if need_Ymean else None,
            Ys if need_Ystd else None,
            sw,
            nz,
        )

    @staticmethod
    def _downdate(total, WXv, Bv, mA, mB, sA, sB, sw, center):
        out = total - WXv.T @ Bv
        if center:
            out -= sw * (mA.T @ mB)
        if sA is not None and sB is not None:
            return out / (sA.T @ sB)
        if sA is not None:
            return out / sA.T
        if sB is not None:
            return out / sB
        return out

    def fold(self, val, want_XTX=True, want_XTY=True):
        """Everything ``_training_matrices`` computes for one validation index set."""
        cX, cY, sX, sY = self.cX, self.cY, self.sX, self.sY
        if not want_XTX and not want_XTY:
            raise ValueError(ERR_NOTHING)
        if want_XTY and self.Y is None:
            raise ValueError(ERR_NO_Y)
        val = np.asarray(val)
        WXv = self.WX[val]
        Xv = WXv if self.w is None else self.X[val]
        WYv = Yv = None
        if want_XTY:
            Yv = self.Y[val]
            WYv = Yv if (self.w is None or not (cX or cY or sY)) else self.WY[val]
        Xm, Xs, Ym, Ys, sw, nz = self._fold_stats(
            val,
            WXv,
            WYv,
            need_Xmean=cX or (want_XTY and cY),
            need_Xstd=sX,
            need_Ymean=want_XTY and (cX or cY),
            need_Ystd=want_XTY and sY,
        )
        XTX = XTY = None
        if want_XTX:
            XTX = self._downdate(self.XTX, WXv, Xv, Xm, Xm, Xs, Xs, sw, cX)
        if want_XTY:
            XTY = self._downdate(self.XTY, WXv, Yv, Xm, Ym, Xs, Ys, sw, cX or cY)
        return FoldResult(XTX, XTY, Xm, Xs, Ym, Ys, sw, nz)

    # -- reference-shaped entry points ----------------------------------------------------
    def training_XTX(self, validation_indices):
        r = self.fold(validation_indices, True, False)
        return r.XTX, (r.X_mean, r.X_std, r.Y_mean, r.Y_std)

    def training_XTY(self, validation_indices):
        r = self.fold(validation_indices, False, True)
        return r.XTY, (r.X_mean, r.X_std, r.Y_mean, r.Y_std)

    def training_XTX_XTY(self, validation_indices):
        r = self.fold(validation_indices, True, True)
        return (r.XTX, r.XTY), (r.X_mean, r.X_std, r.Y_mean, r.Y_std)

    def training_statistics(self, validation_indices):
        cX, cY, sX, sY = self.cX, self.cY, self.sX, self.sY
        val = np.asarray(validation_indices)
        has_Y = self.Y is not None
        WXv = self.WX[val]
        WYv = None
        if has_Y:
            WYv = (
                self.Y[val]
                if (self.w is None or not (cX or cY or sY))
                else self.WY[val]
            )
        return self._fold_stats(
            val,
            WXv,
            WYv,
            need_Xmean=cX or sX,
            need_Xstd=sX,
            need_Ymean=(cY or sY) and has_Y,
            need_Ystd=sY and has_Y,
        )[:4]


# --------------------------------------------------------------------------------------
# Seeded inputs of the benchmark configurations (benchmarks/benchmark.py:223-232)
# --------------------------------------------------------------------------------------
def make_inputs(N, K, M, P, dtype=np.float64, seed=42, w_offset=0.0):
    rng = np.random.default_rng(seed)
    X = rng.random((N, K)).astype(dtype, copy=False)
    Y = rng.random((N, M)).astype(dtype, copy=False)
    w = (rng.random(N) + w_offset).astype(dtype, copy=False)
    folds = np.arange(N) % P
    return X, Y, w, folds


def rel_fro(a, b):
    """||a-b||_F / ||b||_F with b the reference."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    den = np.linalg.norm(b)
    num = np.linalg.norm(a - b)
    return float(num / den) if den > 0 else float(num)
