"""
TEST INFRASTRUCTURE.  Second, algorithm-independent oracle: the training-set matrices recomputed from the TRAINING rows
(center, scale, weight, then X^T W X / X^T W Y), i.e. what the fast downdating algorithm must equal mathematically.
Plays the role of the reference's own correctness oracle `NaiveCVMatrix` (tests/naive_cvmatrix.py:21-277, Algorithms
2/4/6 of Engstrøm & Jensen 2025) - restated here from the definitions, float64 throughout.  Agreement is to rounding
only (the reference's tests use atol = 1e-8, tests/test_cvmatrix.py:420-537).
"""

import numpy as np


def naive_training_matrices(X, Y, w, val, center_X=True, center_Y=True, scale_X=True, scale_Y=True, ddof=1):
    """Returns dict(XTX, XTY, X_mean, X_std, Y_mean, Y_std) for the rows NOT in `val`."""
    X = np.asarray(X, dtype=np.float64)
    N = X.shape[0]
    keep = np.ones(N, dtype=bool)
    keep[np.asarray(val)] = False
    Xt = X[keep]
    Yt = None if Y is None else np.asarray(Y, dtype=np.float64)[keep]
    wt = np.ones(Xt.shape[0]) if w is None else np.asarray(w, dtype=np.float64).reshape(-1)[keep]
    sw = wt.sum()
    nnz = np.count_nonzero(wt)
    out = {}

    def stats(A):
        mean = (wt[:, None] * A).sum(axis=0, keepdims=True) / sw
        var = (wt[:, None] * (A - mean) ** 2).sum(axis=0, keepdims=True) / ((nnz - ddof) * sw / nnz)
        std = np.sqrt(np.maximum(var, 0))
        std[std <= 1e-14] = 1.0
        return mean, std

    Xm, Xs = stats(Xt)
    Xc = Xt - Xm if center_X else Xt
    if scale_X:
        Xc = Xc / Xs
    out["X_mean"], out["X_std"] = Xm, Xs
    out["XTX"] = (Xc * wt[:, None]).T @ Xc
    if Yt is not None:
        Ym, Ys = stats(Yt)
        # the reference centres X^T W Y with BOTH means whenever either flag is set (cvmatrix/cvmatrix.py:852-863):
        # sum_i w_i (x_i - mx)(y_i - my) = sum_i w_i x_i y_i - sw mx my holds if at least one side is centred
        Xc2 = Xt - Xm if (center_X or center_Y) else Xt
        Yc = Yt
        if scale_X:
            Xc2 = Xc2 / Xs
        if scale_Y:
            Yc = Yc / Ys
        out["Y_mean"], out["Y_std"] = Ym, Ys
        out["XTY"] = (Xc2 * wt[:, None]).T @ Yc
    return out
