"""
Leave-few-out in the streaming form (k_few_operands + k_few_tiles: one operand row per validation row, n + 1 FMAs and a
reciprocal scaling per element) against the exact form (cvmx_set_loo_mode(1): k_small_folds, numpy's operation order with
IEEE division) and the oracle: uneven folds of 0 .. 16 rows in one batch, every flag combination, odd K, no Y, float32.
"""

import numpy as np
import pytest

from cvmatrix_oracle import OracleCVMatrix, make_inputs, rel_fro

pytestmark = pytest.mark.gpu


def _sets(N, sizes, seed):
    rng = np.random.default_rng(seed)
    return [rng.choice(N, size=n, replace=False) if n else np.zeros(0, dtype=np.int64) for n in sizes]


@pytest.mark.parametrize("flags", [(True, True, True, True), (False, False, False, False), (False, True, False, False),
                                   (True, False, True, False), (False, False, True, True)])
def test_streaming_few_rows_match_exact_form_and_oracle(flags):
    from cvmatrix_b200 import CVMatrix

    N, K, M = 900, 141, 6
    X, Y, w, _ = make_inputs(N, K, M, 1, seed=41)
    w[::9] = 0
    sets = _sets(N, [1, 2, 16, 7, 0, 3, 16, 1, 5, 11] * 7, seed=42)     # 70 folds: three launches' worth of 32-fold groups
    orc = OracleCVMatrix(*flags)
    orc.fit(X, Y, w)
    res = {}
    for mode in (0, 1):
        m = CVMatrix(*flags)
        m.set_loo_mode(mode)
        m.fit(X, Y, w)
        m.set_folds(sets)
        res[mode] = m.training_batch()
    a, b = res[0], res[1]
    for key in ("X_mean", "X_std", "Y_mean", "Y_std", "sum_w_train", "nnz_train"):
        if a.get(key) is None:
            assert b.get(key) is None
        else:
            assert np.array_equal(a[key], b[key]), key               # the statistics do not depend on the form
    for f, val in enumerate(sets):
        assert rel_fro(a["XTX"][f], b["XTX"][f]) <= 1e-14 and rel_fro(a["XTY"][f], b["XTY"][f]) <= 1e-13, f
        assert np.array_equal(a["XTX"][f], a["XTX"][f].T), f       # exactly symmetric
        if f % 9 == 0:
            r = orc.fold(val)
            assert rel_fro(a["XTX"][f], r.XTX) <= 1e-12 and rel_fro(a["XTY"][f], r.XTY) <= 1e-12, f


def test_streaming_few_rows_no_y_and_float32():
    from cvmatrix_b200 import CVMatrix

    N, K = 700, 260                                  # three column tiles, the last one 4 columns wide
    X, _, w, _ = make_inputs(N, K, 1, 1, seed=43)
    sets = _sets(N, [4, 9, 2, 16, 3] * 8, seed=44)
    orc = OracleCVMatrix()
    orc.fit(X, None, w)
    m = CVMatrix()
    m.fit(X, None, w)
    m.set_folds(sets)
    out = m.training_batch(return_XTY=False)
    for f in (0, 7, 39):
        r = orc.fold(sets[f], want_XTY=False)
        assert rel_fro(out["XTX"][f], r.XTX) <= 1e-12
        assert np.array_equal(out["X_mean"][f], r.X_mean) and np.array_equal(out["X_std"][f], r.X_std)
    X32, Y32, w32, _ = make_inputs(N, 90, 4, 1, dtype=np.float32, seed=45)
    o32 = OracleCVMatrix(False, False, False, False, dtype=np.float32)
    o32.fit(X32, Y32, w32)
    m32 = CVMatrix(False, False, False, False, dtype=np.float32)
    m32.fit(X32, Y32, w32)
    m32.set_folds(sets)
    out32 = m32.training_batch()
    for f in (1, 20):
        r = o32.fold(sets[f])
        assert out32["XTX"].dtype == np.float32
        assert rel_fro(out32["XTX"][f], r.XTX) <= 1e-5 and rel_fro(out32["XTY"][f], r.XTY) <= 1e-5
