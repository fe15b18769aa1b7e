"""
Fold statistics evaluated INSIDE the Gram kernel (GramParams::fuse_stats: the helper warps of every diagonal tile continue
numpy's sequential column sums over the staged rows) against the separate statistics pass (CVMX_FUSE_STATS=0) and the
oracle: means, stds and matrices must be identical bit for bit between the two, and the statistics bit-identical to numpy.
"""

import os

import numpy as np
import pytest

from cvmatrix_oracle import OracleCVMatrix, OraclePartitioner, make_inputs, rel_fro

pytestmark = pytest.mark.gpu


def _run(fuse, X, Y, w, folds, flags, dtype):
    from cvmatrix_b200 import CVMatrix, Partitioner

    old = os.environ.get("CVMX_FUSE_STATS")
    os.environ["CVMX_FUSE_STATS"] = "1" if fuse else "0"
    try:
        m = CVMatrix(*flags, dtype=dtype)      # the switch is read when the handle is created
    finally:
        if old is None:
            del os.environ["CVMX_FUSE_STATS"]
        else:
            os.environ["CVMX_FUSE_STATS"] = old
    m.fit(X, Y, w)
    m.set_folds(Partitioner(folds))
    return m, m.training_batch(return_XTY=Y is not None)


CASES = [
    # N, K, M, P, dtype, weighted
    (6000, 300, 7, 40, np.float64, True),     # 3 diagonal tiles, K and K + M not multiples of 128, 150-row folds (tail stage)
    (5000, 100, 0, 25, np.float64, True),     # one tile, no Y
    (4000, 130, 5, 16, np.float64, False),    # unweighted, a 2-column last tile row
    (3000, 256, 6, 10, np.float64, True),     # K a multiple of 128: the Y columns are chained by the tile (0, 2) from its B operand
    (3000, 500, 20, 8, np.float64, True),     # Y straddles the last diagonal block (500..511) and a Y-only block (512..519)
    (2000, 128, 3, 8, np.float64, True),      # one diagonal tile + one Y-only tile
    (3000, 250, 6, 10, np.float32, True),     # float32 model on the DMMA kernel
    (2400, 60, 3, 2400 // 17, np.float64, True),   # 17-row folds: one full stage + one row
]


@pytest.mark.parametrize("N,K,M,P,dtype,weighted", CASES)
def test_fused_statistics_are_bit_identical(N, K, M, P, dtype, weighted):
    X, Y, w, folds = make_inputs(N, K, max(M, 1), P, dtype=dtype, seed=17)
    if M == 0:
        Y = None
    if weighted:
        w[::11] = 0
    else:
        w = None
    flags = (True, True, True, True)
    m0, a = _run(False, X, Y, w, folds, flags, dtype)
    m1, b = _run(True, X, Y, w, folds, flags, dtype)
    for key in ("XTX", "XTY", "X_mean", "X_std", "Y_mean", "Y_std", "sum_w_train", "nnz_train", "status"):
        if a.get(key) is None:
            assert b.get(key) is None, key
            continue
        assert np.array_equal(a[key], b[key], equal_nan=True), key
    orc = OracleCVMatrix(*flags, dtype=dtype)
    orc.fit(X, Y, w)
    part = OraclePartitioner(folds)
    for pos in sorted({0, P // 2, P - 1}):
        r = orc.fold(part.get_validation_indices(list(part.folds_dict)[pos]), want_XTY=Y is not None)
        assert np.array_equal(b["X_mean"][pos], r.X_mean) and np.array_equal(b["X_std"][pos], r.X_std)
        if Y is not None:
            assert np.array_equal(b["Y_mean"][pos], r.Y_mean) and np.array_equal(b["Y_std"][pos], r.Y_std)
        if dtype == np.float64:
            assert rel_fro(b["XTX"][pos], r.XTX) <= 1e-12


@pytest.mark.parametrize("flags", [(True, False, False, False), (False, True, False, True), (False, False, True, False), (True, True, False, False)])
def test_fused_statistics_flag_subsets(flags):
    X, Y, w, folds = make_inputs(3000, 140, 4, 12, seed=19)
    _, a = _run(False, X, Y, w, folds, flags, np.float64)
    _, b = _run(True, X, Y, w, folds, flags, np.float64)
    for key in ("XTX", "XTY", "X_mean", "X_std", "Y_mean", "Y_std"):
        if a.get(key) is None:
            assert b.get(key) is None, key
        else:
            assert np.array_equal(a[key], b[key], equal_nan=True), key


def test_fused_statistics_single_index_set():
    """cvmx_training_indices (one ad-hoc validation set, the reference's per-call API) takes the same path."""
    from cvmatrix_b200 import CVMatrix

    X, Y, w, folds = make_inputs(5000, 200, 5, 5, seed=23)
    orc = OracleCVMatrix()
    orc.fit(X, Y, w)
    m = CVMatrix()
    m.fit(X, Y, w)
    val = np.flatnonzero(folds == 2)[::-1].copy()      # descending order: the chains follow the caller's row order
    (XTX, XTY), stats = m.training_XTX_XTY(val)
    r = orc.fold(val)
    for s, g in zip(stats, (r.X_mean, r.X_std, r.Y_mean, r.Y_std)):
        assert np.array_equal(s, g)
    assert rel_fro(XTX, r.XTX) <= 1e-12 and rel_fro(XTY, r.XTY) <= 1e-12
