"""
GPU tests (-m gpu) of the binade scan (cvmatrix_b200/csrc/kernels_scan.cuh): float64 column sums in numpy's
sequential order computed without the dependent-add chain.  The bar is bit-exactness - against the chain kernels
(scan mode 0) and against the numpy oracle - on friendly and on adversarial columns.
"""

import numpy as np
import pytest

from cvmatrix_oracle import OracleCVMatrix, make_inputs

pytestmark = pytest.mark.gpu


def _same(a, b):
    return a.shape == b.shape and np.array_equal(a, b, equal_nan=True) and np.array_equal(np.signbit(a), np.signbit(b))


def _adversarial(N, seed=3):
    """Columns that stress the fast / slow decision: cancelling sums, zeros, ties, huge ranges, specials."""
    rng = np.random.default_rng(seed)
    cols = [
        rng.random(N),                                              # friendly
        rng.standard_normal(N),                                     # sum wanders around zero: never fast
        rng.standard_normal(N) + 3.0,
        np.zeros(N),                                                # identity segments
        np.concatenate([np.zeros(N // 2), np.ones(N - N // 2)]),
        rng.integers(0, 1000, N).astype(np.float64),                # exact sums
        rng.random(N).astype(np.float32).astype(np.float64),        # few mantissa bits: many exact ties
        rng.integers(0, 16, N) * 2.0 ** -53 + (np.arange(N) == 0),  # half-ulp steps on top of 1.0
        np.exp(rng.standard_normal(N) * 8),                         # 30 binades of dynamic range
        -rng.random(N),                                             # negative running sum
        np.where(np.arange(N) == N // 3, 1e300, rng.random(N)),     # one huge element
        np.where(np.arange(N) == N // 2, np.inf, rng.random(N)),
        np.where(np.arange(N) == 2 * N // 3, np.nan, rng.random(N)),
        rng.random(N) * 1e-310,                                     # denormals
        np.tile([1e10, -1e10 + 1.0], N // 2 + 1)[:N] + rng.random(N),
        np.concatenate([[1024.0], -rng.random(N - 1) * 1e-13]),     # creeping along a binade boundary
    ]
    return np.stack(cols, axis=1)


@pytest.mark.parametrize("weighted", [True, False])
def test_scan_equals_chain_and_oracle_friendly(weighted):
    from cvmatrix_b200 import CVMatrix

    X, Y, w, folds = make_inputs(150_001, 70, 3, 2, seed=5)
    X *= 1e3
    w = w if weighted else None
    orc = OracleCVMatrix()
    orc.fit(X, Y, w)
    res = {}
    for mode in (0, 2):
        m = CVMatrix()
        m.set_scan_mode(mode)
        m.fit(X, Y, w)                       # 84 MB: the chunk-pipelined upload, one chunk
        assert (m.scan_launch_count > 0) == (mode == 2)
        vals = [np.flatnonzero(folds == 0), np.flatnonzero(folds == 1), np.array([5, 3, 3, -1, 150000, 77, -150001] * 400)]
        res[mode] = [m.sum_X, m.sum_sq_X, m.sum_Y, m.sum_sq_Y] + [s for v in vals for s in m.training_statistics(v)]
        n0 = m.scan_launch_count
        m.training_statistics(vals[0])
        assert (m.scan_launch_count > n0) == (mode == 2)
    for a, b in zip(res[0], res[2]):
        assert _same(a, b)
    for got, want in zip(res[2][:4], (orc.sum_X, orc.sum_sq_X, orc.sum_Y, orc.sum_sq_Y)):
        assert _same(got, want)
    for v_i, v in enumerate([np.flatnonzero(folds == 0), np.flatnonzero(folds == 1)]):
        for got, want in zip(res[2][4 + 4 * v_i: 8 + 4 * v_i], orc.training_statistics(v)):
            assert _same(got, want)


@pytest.mark.parametrize("spec", ["1", "0"])
def test_scan_adversarial_columns_bit_exact(spec, monkeypatch):
    """spec = "1": fold statistics run passes 1 and 3 in one read of the rows from guessed proxies (k_scan_spec), verified
    by the prefix pass; "0": the four-pass form everywhere."""
    from cvmatrix_b200 import CVMatrix

    monkeypatch.setenv("CVMX_SCAN_SPEC", spec)   # read by cvmx_create

    N = 60_000
    A = _adversarial(N)
    X = np.concatenate([A, A[:, :3] * 7.0, np.random.default_rng(1).random((N, 45))], axis=1)   # 64 columns = 2 groups
    Y = A[:, [1, 5]]
    w = np.random.default_rng(2).random(N)
    w[::17] = 0.0
    folds = np.arange(N) % 2
    with np.errstate(all="ignore"):
        orc = OracleCVMatrix()
        orc.fit(X, Y, w)
        res = {}
        for mode in (0, 2):
            m = CVMatrix()
            m.set_scan_mode(mode)
            m.fit(X, Y, w)
            res[mode] = [m.sum_X, m.sum_sq_X, m.sum_Y, m.sum_sq_Y] + list(m.training_statistics(np.flatnonzero(folds == 1)))
        for a, b in zip(res[0], res[2]):
            assert _same(a, b)
        for got, want in zip(res[2][:4], (orc.sum_X, orc.sum_sq_X, orc.sum_Y, orc.sum_sq_Y)):
            assert _same(got, want)
        for got, want in zip(res[2][4:], orc.training_statistics(np.flatnonzero(folds == 1))):
            assert _same(got, want)


def test_scan_chunked_fit_accumulates_across_chunks():
    """fit of a host matrix larger than one 128 MB upload chunk: the scan continues the running sums chunk by chunk."""
    from cvmatrix_b200 import CVMatrix

    rng = np.random.default_rng(11)
    N, K = 450_000, 48
    X = rng.random((N, K))
    X[:, 5] = rng.standard_normal(N)       # one group has a never-fast column: handed to the chain kernel
    w = rng.random(N)
    res = {}
    for mode in (0, 2):
        m = CVMatrix(center_Y=False, scale_Y=False)
        m.set_scan_mode(mode)
        m.fit(X, None, w)
        res[mode] = (m.sum_X, m.sum_sq_X)
    assert _same(res[0][0], res[2][0]) and _same(res[0][1], res[2][1])
    WX = X * w[:, None]
    assert _same(res[2][0], np.sum(WX, axis=0, keepdims=True)) and _same(res[2][1], np.sum(WX * X, axis=0, keepdims=True))


@pytest.mark.parametrize("n_shards", [1, 3])
def test_scan_column_sharded_statistics(n_shards):
    import ctypes as C

    import torch

    from cvmatrix_b200 import CVMatrix, Partitioner, _lib
    from cvmatrix_b200.distributed import _DevArray

    X, Y, w, folds = make_inputs(40_000, 200, 6, 4, seed=9)
    m = CVMatrix()
    m.fit(X, Y, w)
    m.set_folds(Partitioner(folds))
    m.set_scan_mode(0)
    ref = m.training_batch(out="numpy")
    m.set_scan_mode(2)
    lib, h = m._lib, m._h
    total = None
    for s in range(n_shards):
        sp, sc = C.c_void_p(), C.c_int64()
        _lib.check(lib.cvmx_sharded_stats(h, 0, 4, s, n_shards, C.byref(sp), C.byref(sc)), h)
        _lib.check(lib.cvmx_sharded_stats_wait(h), h)
        m.sync()
        view = torch.as_tensor(_DevArray(sp.value, sc.value, "<f8"), device="cuda")
        total = view.clone() if total is None else total + view
    assert m.scan_launch_count >= n_shards
    ld = lib.cvmx_ld(h)
    st = total.cpu().numpy().reshape(4, 2, ld)
    for f in range(4):
        assert np.array_equal(st[f, 0, :200], ref["X_mean"][f][0]) and np.array_equal(st[f, 1, :200], ref["X_std"][f][0])
        assert np.array_equal(st[f, 0, 200:206], ref["Y_mean"][f][0]) and np.array_equal(st[f, 1, 200:206], ref["Y_std"][f][0])
