"""
Generates tests/golden/*.npz from the LIVE reference (sm00thix/cvmatrix v3.2.1 imported
from /root/reference).  Run in the build container only:

    python tests/golden/make_golden.py

Each fixture stores the exact inputs, the constructor options and every output of the
reference (fit attributes, the four training_* entry points per validation set, raised
ValueError messages) so that tests can check the oracle and the CUDA path anywhere
without the reference being present.  numpy / OpenBLAS versions are recorded in the
manifest because the GEMM bits depend on them (stats do not).
"""

import itertools
import json
import os
import sys

import numpy as np

REF = os.environ.get("CVMATRIX_REFERENCE", "/root/reference")
sys.path.insert(0, REF)
from cvmatrix import CVMatrix, Partitioner  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
STAT_NAMES = ("X_mean", "X_std", "Y_mean", "Y_std")


def _record_call(store, prefix, fn, val):
    try:
        res = fn(val)
    except ValueError as e:
        store[prefix + "/error"] = np.array(str(e))
        return
    if fn.__name__ == "training_statistics":
        mats, stats = (), res
    elif fn.__name__ == "training_XTX_XTY":
        mats, stats = (("XTX", res[0][0]), ("XTY", res[0][1])), res[1]
    elif fn.__name__ == "training_XTX":
        mats, stats = (("XTX", res[0]),), res[1]
    else:
        mats, stats = (("XTY", res[0]),), res[1]
    for name, m in mats:
        store[f"{prefix}/{name}"] = m
    for name, s in zip(STAT_NAMES, stats):
        if s is not None:
            store[f"{prefix}/{name}"] = s


def make_case(store, manifest, name, X, Y, w, val_sets, flags, ddof, dtype, methods, inputs=None):
    model = CVMatrix(*flags, ddof=ddof, dtype=dtype)
    model.fit(X, Y, w)
    if inputs is None:
        inputs = name
    if f"{inputs}/in/X" not in store:
        store[f"{inputs}/in/X"] = np.asarray(X)
        if Y is not None:
            store[f"{inputs}/in/Y"] = np.asarray(Y)
        if w is not None:
            store[f"{inputs}/in/w"] = np.asarray(w)
        for i, v in enumerate(val_sets):
            store[f"{inputs}/in/val{i}"] = np.asarray(v, dtype=np.int64)
    for attr in ("XTX", "XTY", "sum_X", "sum_Y", "sum_sq_X", "sum_sq_Y"):
        a = getattr(model, attr)
        if a is not None:
            store[f"{name}/fit/{attr}"] = a
    if model.sum_w is not None:
        store[f"{name}/fit/sum_w"] = np.asarray(model.sum_w)
        store[f"{name}/fit/num_nonzero_w"] = np.asarray(model.num_nonzero_w)
    for i, v in enumerate(val_sets):
        for m in methods:
            _record_call(store, f"{name}/out/val{i}/{m}", getattr(model, m), np.asarray(v))
    manifest["cases"].append(
        dict(name=name, flags=[bool(f) for f in flags], ddof=int(ddof), dtype=np.dtype(dtype).name,
             has_Y=Y is not None, weighted=w is not None, n_val_sets=len(val_sets), methods=list(methods),
             inputs=inputs)
    )


ALL = ("training_XTX", "training_XTY", "training_XTX_XTY", "training_statistics")


def main():
    manifest = dict(
        reference="sm00thix/cvmatrix v3.2.1 (/root/reference)",
        numpy=np.__version__,
        blas=str(np.show_config(mode="dicts").get("Build Dependencies", {}).get("blas", {}).get("name", "?"))
        + " " + str(np.show_config(mode="dicts").get("Build Dependencies", {}).get("blas", {}).get("version", "?")),
        cases=[],
    )
    store = {}

    # 1. README quick-start (BASELINE.json configs[0]): N=100, K=50, M=10, 5 folds
    rng = np.random.default_rng(42)
    N, K, M, P = 100, 50, 10, 5
    X, Y, w = rng.random((N, K)), rng.random((N, M)), rng.random(N) + 0.1
    part = Partitioner(np.arange(N) % P)
    make_case(store, manifest, "quickstart", X, Y, w, [part.get_validation_indices(f) for f in part.folds_dict],
              (True,) * 4, 1, np.float64, ("training_XTX_XTY", "training_statistics"))

    # 2. full flag sweep x {weights with zeros, none} x ddof x {Y, no Y}, 3 uneven folds + odd index sets
    rng = np.random.default_rng(1)
    N, K, M = 61, 6, 3
    X = rng.normal(size=(N, K)) * np.array([1, 10, 0.1, 5, 1, 2]) + np.arange(K)
    X[:, 3] = -1.25  # constant column
    Y = rng.normal(size=(N, M)) + 100.0
    w = rng.random(N)
    w[rng.random(N) < 0.2] = 0.0
    labels = rng.integers(0, 3, size=N)
    part = Partitioner(labels)
    vals = [part.get_validation_indices(f) for f in part.folds_dict]
    vals += [np.array([7]), np.array([30, 2, 2, -1, 15]), np.array([], dtype=np.int64)]
    for flags in itertools.product((False, True), repeat=4):
        for use_w, use_Y, ddof in itertools.product((True, False), (True, False), (0, 1)):
            tag = "".join("TF"[not f] for f in flags) + f"_w{int(use_w)}_y{int(use_Y)}_d{ddof}"
            if "sweep_inputs/in/X" not in store:
                store["sweep_inputs/in/X"], store["sweep_inputs/in/Y"], store["sweep_inputs/in/w"] = X, Y, w
                for i, v in enumerate(vals):
                    store[f"sweep_inputs/in/val{i}"] = np.asarray(v, dtype=np.int64)
            make_case(store, manifest, "sweep_" + tag, X, Y if use_Y else None, w if use_w else None, vals,
                      flags, ddof, np.float64, ALL, inputs="sweep_inputs")
    # (the sweep shares one input set: X, Y, w are always stored; cases with w0 / y0 ignore them)

    # 3. float32, all flags, weighted and unweighted
    # (well-conditioned uniform data: float32 parity at 1e-5 is only meaningful without
    # catastrophic cancellation - SURVEY.md Appendix B)
    rng32 = np.random.default_rng(2)
    X32, Y32 = rng32.random((N, K), dtype=np.float32), rng32.random((N, M), dtype=np.float32)
    w32 = rng32.random(N, dtype=np.float32)
    w32[::7] = 0
    for use_w in (True, False):
        make_case(store, manifest, f"f32_w{int(use_w)}", X32, Y32, w32 if use_w else None, vals, (True,) * 4, 1,
                  np.float32, ALL)
    make_case(store, manifest, "f32_raw", X32, Y32, w32, vals, (False,) * 4, 1, np.float32, ALL)

    # 4. leave-one-out, weighted (BASELINE.json configs[3] in miniature)
    rng = np.random.default_rng(3)
    N, K, M = 40, 5, 2
    X, Y, w = rng.random((N, K)), rng.random((N, M)), rng.random(N)
    make_case(store, manifest, "loo", X, Y, w, [np.array([i]) for i in range(10)], (True,) * 4, 1, np.float64,
              ("training_XTX_XTY",))

    # 5. K = M = 1 given as 1-D integer arrays (pairwise column sums), integer weights
    rng = np.random.default_rng(4)
    x1, y1, w1 = rng.integers(0, 50, 300), rng.integers(0, 9, 300), rng.integers(0, 3, 300)
    make_case(store, manifest, "one_dim", x1, y1, w1, [np.arange(0, 300, 3), np.arange(150, 300)], (True,) * 4, 1,
              np.float64, ALL)
    make_case(store, manifest, "one_dim_f32", x1, y1, w1, [np.arange(0, 300, 3)], (True,) * 4, 0, np.float32, ALL)

    # 6. degenerate folds: all training weights zero / one non-zero training weight
    rng = np.random.default_rng(5)
    N, K, M = 30, 4, 2
    X, Y = rng.random((N, K)), rng.random((N, M))
    wz = np.zeros(N)
    wz[:10] = rng.random(10) + 0.5
    val = np.arange(10)
    for flags in ((True,) * 4, (False,) * 4, (False, True, False, False), (False, False, True, False)):
        tag = "".join("TF"[not f] for f in flags)
        make_case(store, manifest, "allzero_" + tag, X, Y, wz, [val], flags, 1, np.float64, ALL)
    wz2 = wz.copy()
    wz2[20] = 0.75
    make_case(store, manifest, "ddof_TTTT", X, Y, wz2, [val], (True,) * 4, 1, np.float64, ALL)
    make_case(store, manifest, "ddof_TTFF", X, Y, wz2, [val], (True, True, False, False), 1, np.float64, ALL)
    # unweighted: validation set = everything but one row -> nnz_train (=1) <= ddof
    make_case(store, manifest, "ddof_unweighted", X, Y, None, [np.arange(N - 1)], (True,) * 4, 1, np.float64, ALL)

    # 7. summation-order pin: long columns (sequential order) and > 128 weights (pairwise splits)
    rng = np.random.default_rng(6)
    N, K, M = 3001, 5, 2
    X = rng.random((N, K)) * 1e3
    Y = rng.random((N, M))
    w = rng.random(N)
    part = Partitioner(np.arange(N) % 2)
    make_case(store, manifest, "order_pin", X, Y, w, [part.get_validation_indices(0), np.arange(1000, 1131)],
              (True,) * 4, 1, np.float64, ("training_XTX_XTY",))

    # pack: one flat float64 array per case (f32 / int64 values are exactly representable),
    # layout described in the manifest -> a few hundred zip members instead of ~9000
    packed = {}
    layout = {}
    for key, arr in store.items():
        case, rest = key.split("/", 1)
        arr = np.asarray(arr)
        if arr.dtype.kind in "US":
            layout.setdefault(case, []).append([rest, "str", str(arr)])
            continue
        flat = arr.astype(np.float64).reshape(-1)
        assert np.array_equal(flat.astype(arr.dtype).reshape(arr.shape), arr, equal_nan=True)
        layout.setdefault(case, []).append([rest, arr.dtype.name, list(arr.shape)])
        packed.setdefault(case, []).append(flat)
    np.savez_compressed(os.path.join(OUT, "cvmatrix_golden.npz"),
                        __layout__=np.array(json.dumps(layout, separators=(",", ":"))),
                        **{k: np.concatenate(v) if v else np.zeros(0) for k, v in packed.items()})

    # Partitioner fixtures (JSON: keys must survive mixed types -> stored as repr strings in order)
    pcases = []
    rng = np.random.default_rng(8)
    for nm, folds in (
        ("mod5", (np.arange(23) % 5).tolist()),
        ("random_ints", rng.integers(-3, 4, size=40).tolist()),
        ("mixed_hashables", [0, "one", 2, 2, "one", 1.0, True, (1, 2), None, (1, 2)]),
        ("loo", list(range(12))),
        ("floats", [0.5, 1.5, 0.5, 2.0, 2, 1.5]),
        ("empty", []),
    ):
        d = Partitioner(folds).folds_dict
        pcases.append(dict(name=nm, folds_repr=repr(folds), keys_repr=[repr(k) for k in d],
                           indices=[v.tolist() for v in d.values()], index_dtype=str(next(iter(d.values())).dtype) if d else "int64"))
    manifest["partitioner"] = pcases
    with open(os.path.join(OUT, "manifest.json"), "w") as f:
        json.dump(manifest, f, indent=1)
    size = os.path.getsize(os.path.join(OUT, "cvmatrix_golden.npz"))
    print(f"{len(manifest['cases'])} cases, {len(store)} arrays, {size/1e3:.0f} kB")


if __name__ == "__main__":
    main()
