"""
GPU tests (-m gpu) of the fused fit + folds path (cvmx_fit_folds): for a true partition every row is contracted once,
per fold, behind the chunked upload (XtWX = sum of the fold Grams) and training_batch finishes the folds from the kept
Grams.  Results must agree with the two-pass path and with the numpy oracle; anything that is not a partition, or
not the chunk-pipelined upload, silently takes the ordinary path.
"""

import numpy as np
import pytest

from cvmatrix_oracle import OracleCVMatrix, rel_fro

pytestmark = pytest.mark.gpu
STATS = ("X_mean", "X_std", "Y_mean", "Y_std")


def _data(N, K, M, seed):
    rng = np.random.default_rng(seed)
    X, Y, w = rng.random((N, K)), rng.random((N, M)), rng.random(N)
    w[::11] = 0.0
    return X, Y, w


@pytest.mark.parametrize("labels", ["strided", "blocks", "random"])
def test_fused_equals_two_pass_and_oracle(labels):
    from cvmatrix_b200 import CVMatrix, Partitioner

    N, K, M, P = 450_000, 48, 3, 5          # 173 MB: two upload chunks
    X, Y, w = _data(N, K, M, 31)
    folds = {"strided": np.arange(N) % P, "blocks": np.arange(N) // (N // P + 1),
             "random": np.random.default_rng(5).integers(0, P, N)}[labels]
    part = Partitioner(folds)
    a = CVMatrix(copy=False)
    a.fit(X, Y, w, folds=part)
    assert a.folds_cached
    ra = a.training_batch(out="numpy")
    b = CVMatrix(copy=False)
    b.fit(X, Y, w)
    b.set_folds(part)
    assert not b.folds_cached
    rb = b.training_batch(out="numpy")
    assert rel_fro(a.XTX, b.XTX) <= 1e-14 and rel_fro(a.XTY, b.XTY) <= 1e-14
    for k in ("sum_X", "sum_Y", "sum_sq_X", "sum_sq_Y"):
        assert np.array_equal(getattr(a, k), getattr(b, k))
    for f in range(P):
        assert rel_fro(ra["XTX"][f], rb["XTX"][f]) <= 1e-12 and rel_fro(ra["XTY"][f], rb["XTY"][f]) <= 1e-12
        for s in STATS:
            assert np.array_equal(ra[s][f], rb[s][f])
    orc = OracleCVMatrix(copy=False)
    orc.fit(X, Y, w)
    assert rel_fro(a.XTX, orc.XTX) <= 1e-13 and rel_fro(a.XTY, orc.XTY) <= 1e-13
    r = orc.fold(part.get_validation_indices(list(part.folds_dict)[1]))
    # centred XTY alone sits on the cancellation floor of ANY independent summation order at this N (SURVEY.md
    # Appendix B; tests/test_gpu_fullsize.py holds it to 2e-11 at N = 1M): XTX and the joint matrix carry the 1e-12 bar
    assert rel_fro(ra["XTX"][1], r.XTX) <= 1e-12 and rel_fro(ra["XTY"][1], r.XTY) <= 1e-11
    assert rel_fro(np.hstack([ra["XTX"][1], ra["XTY"][1]]), np.hstack([r.XTX, r.XTY])) <= 1e-12
    assert np.array_equal(ra["X_mean"][1], r.X_mean) and np.array_equal(ra["Y_std"][1], r.Y_std)
    # a sub-range of the folds, the per-call API and a new CSR all keep working
    sub = a.training_batch(2, 4, out="numpy")
    assert np.array_equal(sub["XTX"][0], ra["XTX"][2]) and np.array_equal(sub["XTY"][1], ra["XTY"][3])
    (XTX, XTY), _ = a.training_XTX_XTY(part.get_validation_indices(list(part.folds_dict)[0]))
    assert rel_fro(XTX, ra["XTX"][0]) <= 1e-12
    a.set_folds([np.arange(0, N, 7), np.arange(3, N, 9)])
    assert not a.folds_cached
    rc = a.training_batch(out="numpy")
    r7 = orc.fold(np.arange(0, N, 7))
    assert rel_fro(rc["XTX"][0], r7.XTX) <= 1e-12


def test_fused_falls_back_quietly():
    from cvmatrix_b200 import CVMatrix, Partitioner, _lib

    X, Y, w = _data(20_000, 40, 2, 32)      # 6 MB: not the chunked upload
    part = Partitioner(np.arange(20_000) % 4)
    m = CVMatrix()
    m.fit(X, Y, w, folds=part)
    assert not m.folds_cached
    orc = OracleCVMatrix()
    orc.fit(X, Y, w)
    res = m.training_batch(out="numpy")
    r = orc.fold(part.get_validation_indices(2))
    assert rel_fro(res["XTX"][2], r.XTX) <= 1e-12 and np.array_equal(res["X_std"][2], r.X_std)
    # C ABI: overlapping index sets are not a partition -> verified, ordinary path
    N, K = 300_000, 40
    X2 = np.random.default_rng(1).random((N, K))
    off = np.array([0, N // 2, N], dtype=np.int64)
    idx = np.concatenate([np.arange(N // 2), np.arange(N // 2)]).astype(np.int64)
    m2 = CVMatrix(center_Y=False, scale_Y=False)
    p = lambda a: a.ctypes.data  # noqa: E731
    rc = m2._lib.cvmx_fit_folds(m2._h, p(X2), N, K, K, None, 0, 0, None, _lib.HOST, p(off), p(idx), 2, 0)
    assert rc == 0 and m2._lib.cvmx_folds_are_cached(m2._h) == 0


@pytest.mark.parametrize("dtype,weighted,has_Y", [(np.float32, True, True), (np.float64, False, False), (np.float64, True, False)])
def test_fused_variants(dtype, weighted, has_Y):
    """float32, unweighted and X-only fits take the fused path too and agree with the two-pass path."""
    from cvmatrix_b200 import CVMatrix, Partitioner

    N, K, M, P = 300_000, 64, 2, 4
    rng = np.random.default_rng(41)
    X = rng.random((N, K)).astype(dtype)
    Y = rng.random((N, M)).astype(dtype) if has_Y else None
    w = rng.random(N).astype(dtype) if weighted else None
    part = Partitioner(np.arange(N) % P)
    flags = dict(center_X=False, center_Y=False) if dtype == np.float32 else {}   # float32: un-centred
    a = CVMatrix(dtype=dtype, copy=False, **flags)
    a.fit(X, Y, w, folds=part)
    assert a.folds_cached
    b = CVMatrix(dtype=dtype, copy=False, **flags)
    b.fit(X, Y, w)
    b.set_folds(part)
    kw = dict(return_XTY=has_Y, out="numpy")
    ra, rb = a.training_batch(**kw), b.training_batch(**kw)
    tol = 1e-12 if dtype == np.float64 else 1e-5
    assert rel_fro(a.XTX, b.XTX) <= tol
    for f in range(P):
        assert rel_fro(ra["XTX"][f], rb["XTX"][f]) <= tol
        if has_Y:
            assert rel_fro(ra["XTY"][f], rb["XTY"][f]) <= tol
        for s in STATS:
            assert (ra[s] is None) == (rb[s] is None)
            if ra[s] is not None:
                assert np.array_equal(ra[s][f], rb[s][f])
    assert np.array_equal(ra["sum_w_train"], rb["sum_w_train"]) and np.array_equal(ra["status"], rb["status"])
