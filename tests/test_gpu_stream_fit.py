"""
GPU tests (-m gpu) of the streaming / sharded fit (cvmx_fit_begin / cvmx_fit_rows / cvmx_fit_end): the same fitted
state as CVMatrix.fit - numpy-order sums bit-exact, totals to 1e-13 (the Gram is accumulated block by block) - from
host blocks, device blocks, uneven and out-of-order blocks; folds evaluated on top of it match the oracle.
"""

import numpy as np
import pytest

from cvmatrix_oracle import OracleCVMatrix, make_inputs, rel_fro

pytestmark = pytest.mark.gpu


def _check_fit(m, orc, weighted, has_Y):
    assert rel_fro(m.XTX, orc.XTX) <= 1e-13
    assert np.array_equal(m.sum_X, orc.sum_X) and np.array_equal(m.sum_sq_X, orc.sum_sq_X)
    if has_Y:
        assert rel_fro(m.XTY, orc.XTY) <= 1e-13
        assert np.array_equal(m.sum_Y, orc.sum_Y) and np.array_equal(m.sum_sq_Y, orc.sum_sq_Y)
    if weighted:
        assert m.sum_w == orc.sum_w and m.num_nonzero_w == orc.nnz_w


@pytest.mark.parametrize("weighted,has_Y", [(True, True), (False, True), (True, False)])
def test_streamed_host_blocks_equal_fit(weighted, has_Y):
    from cvmatrix_b200 import CVMatrix

    N, K, M = 50_000, 150, 4
    X, Y, w, folds = make_inputs(N, K, M, 4, seed=21)
    w[::13] = 0.0
    Y = Y if has_Y else None
    w = w if weighted else None
    orc = OracleCVMatrix()
    orc.fit(X, Y, w)
    m = CVMatrix()
    m.fit_begin(N, K, M if has_Y else 0, weighted=weighted, max_block_rows=12_000)
    edges = [0, 12_000, 12_001, 20_000, 31_999, 43_999, N]          # uneven blocks, fed out of order
    blocks = list(zip(edges[:-1], edges[1:]))
    for b0, b1 in blocks[::-1]:
        m.fit_rows(b0, X[b0:b1], None if Y is None else Y[b0:b1], None if w is None else w[b0:b1])
    m.fit_end()
    _check_fit(m, orc, weighted, has_Y)
    val = np.flatnonzero(folds == 2)
    if has_Y:
        (XTX, XTY), stats = m.training_XTX_XTY(val)
        r = orc.fold(val)
        assert rel_fro(XTX, r.XTX) <= 1e-12 and rel_fro(XTY, r.XTY) <= 1e-12
        for s, g in zip(stats, (r.X_mean, r.X_std, r.Y_mean, r.Y_std)):
            assert np.array_equal(s, g)
    else:
        XTX, stats = m.training_XTX(val)
        ref, rstats = orc.training_XTX(val)
        assert rel_fro(XTX, ref) <= 1e-12
        assert np.array_equal(stats[0], rstats[0]) and np.array_equal(stats[1], rstats[1])
        with pytest.raises(ValueError, match="not provided"):
            m.training_XTY(val)


def test_streamed_device_blocks_and_strided_host_blocks():
    import torch

    from cvmatrix_b200 import CVMatrix

    N, K, M = 30_000, 96, 3
    X, Y, w, folds = make_inputs(N, K, M, 3, seed=22)
    orc = OracleCVMatrix()
    orc.fit(X, Y, w)
    Xd, Yd, wd = (torch.from_numpy(a).cuda() for a in (X, Y, w))
    m = CVMatrix()
    m.fit_begin(N, K, M, weighted=True, max_block_rows=8192)
    for b0 in range(0, N, 8192):
        b1 = min(N, b0 + 8192)
        m.fit_rows(b0, Xd[b0:b1], Yd[b0:b1], wd[b0:b1])
    m.fit_end()
    _check_fit(m, orc, True, True)
    # strided host blocks (a column window of a wider matrix) take the 2-D copy path
    wide = np.zeros((N, K + 10))
    wide[:, 3:3 + K] = X
    m2 = CVMatrix()
    m2.fit_begin(N, K, M, weighted=True, max_block_rows=16_000)
    for b0 in range(0, N, 16_000):
        b1 = min(N, b0 + 16_000)
        m2.fit_rows(b0, wide[b0:b1, 3:3 + K], Y[b0:b1], w[b0:b1])
    m2.fit_end()
    _check_fit(m2, orc, True, True)
    with pytest.raises(ValueError):
        m2.fit_rows(0, X[:10], Y[:10], w[:10])      # not filling any more


def test_fit_sharded_upload_single_rank_and_abi_order():
    from cvmatrix_b200 import CVMatrix, _lib
    from cvmatrix_b200.distributed import fit_sharded_upload

    X, Y, w, folds = make_inputs(20_000, 64, 2, 2, seed=23)
    orc = OracleCVMatrix()
    orc.fit(X, Y, w)
    m = CVMatrix()
    fit_sharded_upload(m, X, Y, w, block_rows=6000)
    _check_fit(m, orc, True, True)
    assert m.X is not None and m.WX.shape == X.shape
    m3 = CVMatrix()
    assert m3._lib.cvmx_fit_end(m3._h, 0, 1) == _lib.ERR_INVALID
    assert m3._lib.cvmx_fit_begin(m3._h, 10, 0, 0, 0, 10) == _lib.ERR_INVALID
