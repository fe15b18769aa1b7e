"""Small workloads for compute-sanitizer (tools/sanitize.sh): the README quick-start shape (cfg 1), the tiny bench
shape, a multi-tile row-split batch, leave-one-out in both forms, leave-few-out, the binade scan, the fused fit + folds
path and the streaming fit.  numpy in / numpy out only (no torch), every result checked against the oracle."""
import os
import sys

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
from cvmatrix_b200 import CVMatrix, Partitioner  # noqa: E402
from cvmatrix_oracle import OracleCVMatrix, make_inputs, rel_fro  # noqa: E402


def check(m, orc, part, folds, tag):
    out = m.training_batch()
    for pos in folds:
        r = orc.fold(part.get_validation_indices(list(part.folds_dict)[pos]))
        assert rel_fro(out["XTX"][pos], r.XTX) <= 1e-12 and rel_fro(out["XTY"][pos], r.XTY) <= 1e-12, tag
        assert np.array_equal(out["X_mean"][pos], r.X_mean) and np.array_equal(out["Y_std"][pos], r.Y_std), tag
    print("ok", tag, flush=True)


def case(N, K, M, P, tag, seed=42, scan=None, loo_mode=None, fused=False):
    X, Y, w, folds = make_inputs(N, K, M, P, seed=seed)
    part = Partitioner(folds)
    orc = OracleCVMatrix()
    orc.fit(X, Y, w)
    m = CVMatrix()
    if scan is not None:
        m.set_scan_mode(scan)
    if loo_mode is not None:
        m.set_loo_mode(loo_mode)
    if fused:
        m.fit(X, Y, w, folds=part)
    else:
        m.fit(X, Y, w)
        m.set_folds(part)
    assert np.array_equal(m.sum_X, orc.sum_X) and m.sum_w == orc.sum_w
    check(m, orc, part, sorted({0, P // 2, P - 1}), tag)
    val = part.get_validation_indices(list(part.folds_dict)[0])
    st = m.training_statistics(val)
    for a, b in zip(st, orc.training_statistics(val)):
        assert np.array_equal(a, b), tag


def case_f32(N, K, M, P, tag):
    """float32 model, un-centred (the regime where numpy-float32 is accurate to 1e-5): the TF32 tensor-core Gram kernels."""
    X, Y, w, folds = make_inputs(N, K, M, P, dtype=np.float32, seed=9)
    orc = OracleCVMatrix(False, False, False, False, dtype=np.float32)
    orc.fit(X, Y, w)
    m = CVMatrix(False, False, False, False, dtype=np.float32)
    m.fit(X, Y, w)
    e = rel_fro(m.XTX, orc.XTX)
    print("f32 totals relFro", e, flush=True)
    assert e <= 1e-5, tag
    part = Partitioner(folds)
    m.set_folds(part)
    out = m.training_batch()
    r = orc.fold(part.get_validation_indices(0))
    assert rel_fro(out["XTX"][0], r.XTX) <= 1e-5 and rel_fro(out["XTY"][0], r.XTY) <= 1e-5, tag
    print("ok", tag, flush=True)


which = sys.argv[1] if len(sys.argv) > 1 else "all"
if which in ("all", "cfg1"):
    case(100, 50, 10, 5, "cfg1 README quick-start")
if which in ("all", "tiny"):
    case(4000, 24, 3, 4, "tiny bench shape")
if which in ("all", "split"):
    case(6100, 130, 5, 2, "two tiles x row-split folds", seed=3)
if which in ("all", "scan"):
    case(5000, 40, 3, 2, "binade scan forced", seed=4, scan=2)
if which in ("all", "loo"):
    case(300, 70, 3, 300, "leave-one-out, streaming form", seed=5, loo_mode=0)
    case(300, 70, 3, 300, "leave-one-out, exact form", seed=5, loo_mode=1)
if which in ("all", "few"):
    case(600, 70, 3, 100, "leave-few-out (6 rows per fold)", seed=6)
if which in ("all", "fused"):
    case(5000, 40, 3, 4, "fused fit + folds", seed=7, fused=True)
if which in ("all", "f32"):
    case_f32(3000, 200, 5, 3, "float32 model (TF32 tensor cores)")
print("SANITIZER_CASES_OK", which)
