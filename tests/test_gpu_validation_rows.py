"""
GPU test (-m gpu) of cvmx_validation_rows / CVMatrix.validation_rows: the held-out rows of a fold, optionally centred
and scaled with the training-set statistics of that fold - bit-identical to numpy's (X[val] - X_mean) / X_std.
"""

import numpy as np
import pytest

from cvmatrix_oracle import make_inputs

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_validation_rows_match_numpy(dtype):
    from cvmatrix_b200 import CVMatrix, Partitioner

    X, Y, w, folds = make_inputs(3000, 37, 3, 4, dtype=dtype, seed=51)
    part = Partitioner(folds)
    m = CVMatrix(dtype=dtype)
    m.fit(X, Y, w)
    m.set_folds(part)
    for f in (0, 3):
        val = part.get_validation_indices(list(part.folds_dict)[f])
        Xv, Yv = m.validation_rows(f)
        assert Xv.dtype == dtype and np.array_equal(Xv, X[val]) and np.array_equal(Yv, Y[val])
        _, stats = m.training_XTX_XTY(val)
        Xc, Yc = m.validation_rows(f, stats)
        assert np.array_equal(Xc, (X[val] - stats[0]) / stats[1]) and np.array_equal(Yc, (Y[val] - stats[2]) / stats[3])
        Xo, Yo = m.validation_rows(f, (stats[0], None, None, stats[3]))          # centre X only, scale Y only
        assert np.array_equal(Xo, X[val] - stats[0]) and np.array_equal(Yo, Y[val] / stats[3])
        Xt, Yt = m.validation_rows(f, stats, out="torch")
        assert Xt.is_cuda and np.array_equal(Xt.cpu().numpy(), Xc) and np.array_equal(Yt.cpu().numpy(), Yc)
    with pytest.raises(ValueError, match="not found"):
        m.validation_rows(4)


def test_validation_rows_without_y_and_abi_errors():
    from cvmatrix_b200 import CVMatrix, _lib

    X, _, w, folds = make_inputs(500, 8, 1, 5, seed=52)
    m = CVMatrix(center_Y=False, scale_Y=False)
    m.fit(X, None, w)
    m.set_folds([np.flatnonzero(folds == k) for k in range(5)] + [np.zeros(0, np.int64)])
    Xv, Yv = m.validation_rows(2)
    assert Yv is None and np.array_equal(Xv, X[folds == 2])
    Xe, _ = m.validation_rows(5)
    assert Xe.shape == (0, 8)
    assert m._lib.cvmx_validation_rows(m._h, 0, None, 1, None, None, _lib.HOST) == _lib.ERR_INVALID   # statistics missing
    assert m._lib.cvmx_validation_rows(m._h, 99, None, 0, None, None, _lib.HOST) == _lib.ERR_INVALID
