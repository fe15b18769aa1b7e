"""
TEST INFRASTRUCTURE.  Pins oracle/cvmatrix_oracle.py to the live reference.

Runs only where /root/reference exists (the build container).  For every combination of
(center_X, center_Y, scale_X, scale_Y) x {weights with zeros, no weights} x ddof {0, 1}
x {Y, no Y} x dtype {f64, f32} x order {"numpy", "explicit"} it demands *bit-identical*
outputs of fit attributes, training_XTX / training_XTY / training_XTX_XTY /
training_statistics, error behaviour, and Partitioner dictionaries.

Usage:  python oracle/check_against_reference.py        (exit code 0 = pinned)
"""

import itertools
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
REF = os.environ.get("CVMATRIX_REFERENCE", "/root/reference")


def _same(a, b):
    if a is None or b is None:
        return a is None and b is None
    a = np.asarray(a)
    b = np.asarray(b)
    return a.shape == b.shape and a.dtype == b.dtype and np.array_equal(a, b, equal_nan=True)


def _same_tree(a, b):
    if isinstance(a, tuple):
        return isinstance(b, tuple) and len(a) == len(b) and all(_same_tree(x, y) for x, y in zip(a, b))
    return _same(a, b)


def _call(f, *args):
    try:
        return ("ok", f(*args))
    except ValueError as e:
        return ("ValueError", str(e))


def main():
    if not os.path.isdir(REF):
        print(f"reference not present at {REF}; nothing to pin against")
        return 2
    sys.path.insert(0, REF)
    from cvmatrix import CVMatrix, Partitioner  # the live reference
    from cvmatrix_oracle import OracleCVMatrix, OraclePartitioner

    rng = np.random.default_rng(7)
    N, K, M = 403, 9, 3
    X0 = rng.normal(size=(N, K)) + 3.0
    X0[:, 4] = 2.5  # constant column -> std replaced by 1
    Y0 = rng.normal(size=(N, M)) * 10
    w0 = rng.random(N)
    w0[rng.random(N) < 0.15] = 0.0
    labels = rng.integers(0, 7, size=N)
    fails = 0
    checks = 0

    # Partitioner: integer labels, mixed hashables, LOO
    for folds in (labels, list(labels), [0, "one", 2, 2, "one", 1.0, True], np.arange(50)):
        a, b = Partitioner(folds).folds_dict, OraclePartitioner(folds).folds_dict
        ok = list(a.keys()) == list(b.keys()) and all(
            _same(a[k], b[k]) for k in a
        )
        checks += 1
        fails += not ok

    part = Partitioner(labels)
    val_sets = [part.get_validation_indices(f) for f in list(part.folds_dict)[:4]]
    val_sets.append(np.array([5]))  # single row
    val_sets.append(np.array([17, 3, 3, 250, -1]))  # unsorted, duplicate, negative
    val_sets.append(np.array([], dtype=int))  # empty

    for dtype in (np.float64, np.float32):
        for use_w, use_Y, ddof in itertools.product((True, False), (True, False), (0, 1)):
            for flags in itertools.product((False, True), repeat=4):
                ref = CVMatrix(*flags, ddof=ddof, dtype=dtype)
                ref.fit(X0, Y0 if use_Y else None, w0 if use_w else None)
                for order in ("numpy", "explicit"):
                    orc = OracleCVMatrix(*flags, ddof=ddof, dtype=dtype, order=order)
                    orc.fit(X0, Y0 if use_Y else None, w0 if use_w else None)
                    pairs = [
                        (ref.XTX, orc.XTX), (ref.XTY, orc.XTY),
                        (ref.sum_X, orc.sum_X), (ref.sum_Y, orc.sum_Y),
                        (ref.sum_sq_X, orc.sum_sq_X), (ref.sum_sq_Y, orc.sum_sq_Y),
                    ]
                    for a, b in pairs:
                        checks += 1
                        if not _same(a, b):
                            fails += 1
                            print("fit attr mismatch", dtype.__name__, use_w, use_Y, ddof, flags, order)
                    if any(flags):
                        checks += 1
                        if use_w:
                            ok = ref.sum_w == orc.sum_w and type(ref.sum_w) is type(orc.sum_w) and ref.num_nonzero_w == orc.nnz_w
                        else:
                            ok = ref.sum_w == orc.sum_w == N
                        fails += not ok
                    for val in val_sets:
                        methods = ["training_XTX", "training_statistics"]
                        methods += ["training_XTY", "training_XTX_XTY"]  # raise without Y: also compared
                        for m in methods:
                            ra = _call(getattr(ref, m), val)
                            rb = _call(getattr(orc, m), val)
                            checks += 1
                            ok = ra[0] == rb[0] and (
                                ra[1] == rb[1] if ra[0] != "ok" else _same_tree(ra[1], rb[1])
                            )
                            if not ok:
                                fails += 1
                                print("fold mismatch", dtype.__name__, use_w, use_Y, ddof, flags, order, m, val[:5])

    # degenerate folds: every training weight zero / nnz <= ddof
    wz = np.zeros(N)
    wz[labels == 0] = 1.0
    val0 = part.get_validation_indices(0)
    for flags in itertools.product((False, True), repeat=4):
        ref = CVMatrix(*flags, ddof=1)
        orc = OracleCVMatrix(*flags, ddof=1, order="explicit")
        ref.fit(X0, Y0, wz)
        orc.fit(X0, Y0, wz)
        for m in ("training_XTX", "training_XTY", "training_XTX_XTY", "training_statistics"):
            ra, rb = _call(getattr(ref, m), val0), _call(getattr(orc, m), val0)
            checks += 1
            ok = ra[0] == rb[0] and (ra[1] == rb[1] if ra[0] != "ok" else _same_tree(ra[1], rb[1]))
            fails += not ok
            if not ok:
                print("degenerate mismatch", flags, m, ra[0], rb[0])
        keep = np.setdiff1d(np.arange(N), val0)[:1]
        wz2 = wz.copy()
        wz2[keep] = 0.5  # exactly one non-zero training weight -> nnz <= ddof
        ref.fit(X0, Y0, wz2)
        orc.fit(X0, Y0, wz2)
        for m in ("training_XTX", "training_XTX_XTY", "training_statistics"):
            ra, rb = _call(getattr(ref, m), val0), _call(getattr(orc, m), val0)
            checks += 1
            ok = ra[0] == rb[0] and (ra[1] == rb[1] if ra[0] != "ok" else _same_tree(ra[1], rb[1]))
            fails += not ok
            if not ok:
                print("ddof mismatch", flags, m, ra, rb)

    # 1-D X / Y (K = M = 1 -> pairwise column sums), integer input, negative weights
    x1 = rng.integers(0, 50, size=300)
    y1 = rng.integers(0, 9, size=300)
    w1 = rng.integers(0, 3, size=300)
    v1 = np.arange(0, 300, 3)
    for dtype in (np.float64, np.float32):
        ref = CVMatrix(dtype=dtype)
        ref.fit(x1, y1, w1)
        for order in ("numpy", "explicit"):
            orc = OracleCVMatrix(dtype=dtype, order=order)
            orc.fit(x1, y1, w1)
            checks += 1
            ok = _same_tree(ref.training_XTX_XTY(v1), orc.training_XTX_XTY(v1))
            fails += not ok
            if not ok:
                print("1-D mismatch", dtype.__name__, order)
    for cls in (CVMatrix, OracleCVMatrix):
        r = _call(cls().fit, X0, Y0, -w0 - 1)
        checks += 1
        fails += r != ("ValueError", "Weights must be non-negative.")

    print(f"{checks} checks, {fails} failures")
    return 0 if fails == 0 else 1


if __name__ == "__main__":
    sys.exit(main())
