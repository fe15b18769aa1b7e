"""
GPU tests (-m gpu) of the row-slab mode (include/cvmx.h "Row-slab mode"; BASELINE config 5: the ROWS of the data set are
sharded across GPUs) that run on ONE GPU, so that the driver's single-GPU box exercises the slab entry points too:

* `RowSlabFolds` with a world of one rank (no process group): chained sums with an empty carry, global weight folds,
  per-slab Gram + owner-side finish;
* two slabs held by two handles on the same device, with the carries handed from one to the other by hand - exactly
  what ranks r and r + 1 do over NCCL (tests/dist_fit_worker.py runs the real thing on 2 GPUs).

Bar: statistics, column sums and weight sums bit-identical to the numpy oracle; matrices to 1e-12 (float64).
"""

import ctypes as C

import numpy as np
import pytest

from cvmatrix_oracle import OracleCVMatrix, make_inputs, rel_fro

pytestmark = pytest.mark.gpu


def _inputs(N=24_000, K=150, M=4, seed=21):
    X, Y, w, _ = make_inputs(N, K, M, 1, seed=seed)
    w[::13] = 0.0
    labels = np.random.default_rng(seed).choice([0, 1, 2, 3], size=N, p=[0.45, 0.3, 0.2, 0.05])
    return X, Y, w, labels


def test_row_slab_world_of_one():
    import torch

    from cvmatrix_b200 import CVMatrix, Partitioner, _lib
    from cvmatrix_b200.distributed import RowSlabFolds

    X, Y, w, labels = _inputs()
    N, K = X.shape
    M = Y.shape[1]
    part = Partitioner(labels)
    orc = OracleCVMatrix()
    orc.fit(X, Y, w)
    m = CVMatrix()
    rs = RowSlabFolds(m, N, K, M, w, block_rows=7000)
    rs.fit((b0, X[b0:b0 + 7000], Y[b0:b0 + 7000]) for b0 in range(0, N, 7000))
    assert rel_fro(m.XTX, orc.XTX) <= 1e-14 and rel_fro(m.XTY, orc.XTY) <= 1e-14
    assert np.array_equal(m.sum_X, orc.sum_X) and np.array_equal(m.sum_sq_X, orc.sum_sq_X)
    assert np.array_equal(m.sum_Y, orc.sum_Y) and np.array_equal(m.sum_sq_Y, orc.sum_sq_Y)
    assert m.sum_w == orc.sum_w and m.num_nonzero_w == orc.nnz_w
    rs.set_folds(part)
    stream = torch.cuda.Stream()
    with torch.cuda.stream(stream):
        _lib.check(m._lib.cvmx_set_stream(m._h, C.c_void_p(stream.cuda_stream)), m._h)
        out = rs.training_batch()
        torch.cuda.synchronize()
        _lib.check(m._lib.cvmx_set_stream(m._h, None), m._h)
    assert (out["fold_begin"], out["fold_end"]) == (0, 4)
    for pos, key in enumerate(part.folds_dict):
        r = orc.fold(part.get_validation_indices(key))
        assert rel_fro(out["XTX"][pos].cpu().numpy(), r.XTX) <= 1e-12 and rel_fro(out["XTY"][pos].cpu().numpy(), r.XTY) <= 1e-12
        st = out["stats"][pos].cpu().numpy()
        assert np.array_equal(st[0, :K], r.X_mean[0]) and np.array_equal(st[1, :K], r.X_std[0])
        assert np.array_equal(st[0, K:], r.Y_mean[0]) and np.array_equal(st[1, K:], r.Y_std[0])


def test_two_slabs_on_one_device_chain_by_hand():
    import torch

    from cvmatrix_b200 import CVMatrix, Partitioner, _lib, sharding

    X, Y, w, labels = _inputs(seed=22)
    N, K = X.shape
    M = Y.shape[1]
    part = Partitioner(labels)
    offsets, indices = part.csr()
    P = offsets.size - 1
    orc = OracleCVMatrix()
    orc.fit(X, Y, w)
    dev = torch.device("cuda", 0)
    wg = torch.from_numpy(w).to(dev)
    vp = lambda t: None if t is None else C.c_void_p(t.data_ptr())  # noqa: E731
    world = 2
    hs, slabs = [], []
    carry = None
    totals = None
    # ---- fit: slab r continues the column chains of slab r - 1; the totals are the sum of the per-slab Grams
    for r in range(world):
        r0, r1 = sharding.slab_rows(r, world, N)
        m = CVMatrix()
        lib, h = m._lib, m._h
        m.fit_begin(r1 - r0, K, M, weighted=True, max_block_rows=8192)
        for b0 in range(r0, r1, 8192):
            b1 = min(r1, b0 + 8192)
            m.fit_rows(b0 - r0, X[b0:b1], Y[b0:b1], w[b0:b1])
        _lib.check(lib.cvmx_fit_end_slab(h, None if carry is None else vp(carry[0]), None if carry is None else vp(carry[1]), vp(wg), N, r0), h)
        ld = int(lib.cvmx_ld(h))
        sp, qp, mc = C.c_void_p(), C.c_void_p(), C.c_int64()
        _lib.check(lib.cvmx_moments_ptr(h, C.byref(sp), C.byref(qp), C.byref(mc)), h)
        from cvmatrix_b200.distributed import _DevArray

        carry = torch.stack([torch.as_tensor(_DevArray(sp.value, ld, "<f8"), device=dev).clone(),
                             torch.as_tensor(_DevArray(qp.value, ld, "<f8"), device=dev).clone()])
        tp, cnt, ldt = C.c_void_p(), C.c_int64(), C.c_int64()
        _lib.check(lib.cvmx_totals_ptr(h, C.byref(tp), C.byref(cnt), C.byref(ldt)), h)
        t = torch.as_tensor(_DevArray(tp.value, cnt.value, "<f8"), device=dev)
        totals = t.clone() if totals is None else totals + t
        hs.append((m, sp, qp, tp, cnt, ld))
        slabs.append((r0, r1))
    for m, sp, qp, tp, cnt, ld in hs:      # what the broadcast / all-reduce leave on every rank
        torch.as_tensor(_DevArray(sp.value, ld, "<f8"), device=dev).copy_(carry[0])
        torch.as_tensor(_DevArray(qp.value, ld, "<f8"), device=dev).copy_(carry[1])
        torch.as_tensor(_DevArray(tp.value, cnt.value, "<f8"), device=dev).copy_(totals)
        torch.cuda.synchronize()
        m._streamed, m.N = True, N
        m._pull_totals()
        assert np.array_equal(m.sum_X, orc.sum_X) and np.array_equal(m.sum_sq_Y, orc.sum_sq_Y)
        assert m.sum_w == orc.sum_w and m.num_nonzero_w == orc.nnz_w
        assert rel_fro(m.XTX, orc.XTX) <= 1e-14 and rel_fro(m.XTY, orc.XTY) <= 1e-14
    # ---- folds: local CSRs + global weight folds, chained fold sums, per-slab Grams summed, finished by slab 0's handle
    raw = torch.zeros((P, 2, hs[0][5]), dtype=torch.float64, device=dev)
    torch.cuda.synchronize()          # the handles work on their own streams
    grams = []
    for (m, *_), (r0, r1) in zip(hs, slabs):
        lib, h = m._lib, m._h
        loc_off, loc_idx = sharding.local_csr(offsets, indices, r0, r1)
        m._upload_csr(loc_off, loc_idx)
        _lib.check(lib.cvmx_set_weight_folds(h, offsets.ctypes.data_as(C.c_void_p), indices.ctypes.data_as(C.c_void_p), P), h)
        _lib.check(lib.cvmx_slab_fold_sums(h, 0, P, vp(raw)), h)      # in place: carry in -> carry out
        m.sync()
    for m, *_ in hs:
        lib, h = m._lib, m._h
        _lib.check(lib.cvmx_slab_finalize_stats(h, 0, P, vp(raw)), h)
        n = lib.cvmx_sharded_gram_count(h, 0, P, 3)
        g = torch.empty((n,), dtype=torch.float64, device=dev)
        _lib.check(lib.cvmx_sharded_gram(h, 0, P, 3, 0, 1, vp(g)), h)
        m.sync()
        grams.append(g)
    gsum = grams[0] + grams[1]
    m0 = hs[0][0]
    oxx = torch.empty((P, K, K), dtype=torch.float64, device=dev)
    oxy = torch.empty((P, K, M), dtype=torch.float64, device=dev)
    ost = torch.empty((P, 2, K + M), dtype=torch.float64, device=dev)
    torch.cuda.synchronize()
    _lib.check(m0._lib.cvmx_sharded_finish(m0._h, 0, 0, P, 3, vp(gsum), vp(oxx), vp(oxy), vp(ost), None, None), m0._h)
    m0.sync()
    for pos, key in enumerate(part.folds_dict):
        r = orc.fold(part.get_validation_indices(key))
        assert rel_fro(oxx[pos].cpu().numpy(), r.XTX) <= 1e-12 and rel_fro(oxy[pos].cpu().numpy(), r.XTY) <= 1e-12
        st = ost[pos].cpu().numpy()
        assert np.array_equal(st[0, :K], r.X_mean[0]) and np.array_equal(st[1, :K], r.X_std[0])
        assert np.array_equal(st[0, K:], r.Y_mean[0]) and np.array_equal(st[1, K:], r.Y_std[0])


def test_three_slabs_decoupled_chain_by_hand():
    """The decoupled slab chain (cvmx_slab_scan_local / _prepare): every slab runs the streaming scan passes from an
    APPROXIMATE start (the sum of the earlier slabs' pass-1 totals) before the exact chains arrive; only the last pass is
    chained.  Sums and statistics must still be bit-identical to numpy's sequential sums over all rows."""
    import torch

    from cvmatrix_b200 import CVMatrix, Partitioner, _lib, sharding
    from cvmatrix_b200.distributed import _DevArray

    X, Y, w, labels = _inputs(N=30_000, seed=23)
    X[:, 3] -= 0.5            # a column whose running sum wanders around zero: segments the scan must hand to the chain kernel
    N, K = X.shape
    M = Y.shape[1]
    part = Partitioner(labels)
    offsets, indices = part.csr()
    P = offsets.size - 1
    orc = OracleCVMatrix()
    orc.fit(X, Y, w)
    dev = torch.device("cuda", 0)
    wg = torch.from_numpy(w).to(dev)
    vp = lambda t: None if t is None else C.c_void_p(t.data_ptr())  # noqa: E731
    world = 3
    hs, slabs, tots = [], [], []
    # ---- fit, phase 1 on every slab: rows in, local scan passes, approximate slab totals
    for r in range(world):
        r0, r1 = sharding.slab_rows(r, world, N)
        m = CVMatrix()
        lib, h = m._lib, m._h
        m.fit_begin(r1 - r0, K, M, weighted=True, max_block_rows=8192)
        for b0 in range(r0, r1, 8192):
            b1 = min(r1, b0 + 8192)
            m.fit_rows(b0 - r0, X[b0:b1], Y[b0:b1], w[b0:b1])
        ld = int(lib.cvmx_ld(h))
        tot = torch.zeros((4 * ld,), dtype=torch.float64, device=dev)
        app = C.c_int32(0)
        _lib.check(lib.cvmx_slab_scan_local(h, -1, 0, vp(wg), N, r0, vp(tot), C.byref(app)), h)
        assert app.value == 1
        m.sync()
        hs.append(m)
        slabs.append((r0, r1))
        tots.append(tot)
    # ---- phase 2 on every slab (what the all-gather provides), then the chain of last passes
    carry, totals, keep = None, None, []
    for r, m in enumerate(hs):
        lib, h = m._lib, m._h
        start = torch.zeros_like(tots[0]) if r == 0 else torch.stack(tots[:r]).sum(0)
        keep.append(start)
        torch.cuda.synchronize()
        _lib.check(lib.cvmx_slab_scan_prepare(h, -1, 0, vp(start)), h)
    for r, m in enumerate(hs):
        lib, h = m._lib, m._h
        r0, r1 = slabs[r]
        _lib.check(lib.cvmx_fit_end_slab(h, None if carry is None else vp(carry[0]), None if carry is None else vp(carry[1]), vp(wg), N, r0), h)
        ld = int(lib.cvmx_ld(h))
        sp, qp, mc = C.c_void_p(), C.c_void_p(), C.c_int64()
        _lib.check(lib.cvmx_moments_ptr(h, C.byref(sp), C.byref(qp), C.byref(mc)), h)
        carry = torch.stack([torch.as_tensor(_DevArray(sp.value, ld, "<f8"), device=dev).clone(),
                             torch.as_tensor(_DevArray(qp.value, ld, "<f8"), device=dev).clone()])
        tp, cnt, ldt = C.c_void_p(), C.c_int64(), C.c_int64()
        _lib.check(lib.cvmx_totals_ptr(h, C.byref(tp), C.byref(cnt), C.byref(ldt)), h)
        t = torch.as_tensor(_DevArray(tp.value, cnt.value, "<f8"), device=dev)
        totals = t.clone() if totals is None else totals + t
    for m in hs:
        lib, h = m._lib, m._h
        ld = int(lib.cvmx_ld(h))
        sp, qp, mc = C.c_void_p(), C.c_void_p(), C.c_int64()
        _lib.check(lib.cvmx_moments_ptr(h, C.byref(sp), C.byref(qp), C.byref(mc)), h)
        tp, cnt, ldt = C.c_void_p(), C.c_int64(), C.c_int64()
        _lib.check(lib.cvmx_totals_ptr(h, C.byref(tp), C.byref(cnt), C.byref(ldt)), h)
        torch.as_tensor(_DevArray(sp.value, ld, "<f8"), device=dev).copy_(carry[0])
        torch.as_tensor(_DevArray(qp.value, ld, "<f8"), device=dev).copy_(carry[1])
        torch.as_tensor(_DevArray(tp.value, cnt.value, "<f8"), device=dev).copy_(totals)
        torch.cuda.synchronize()
        m._streamed, m.N = True, N
        m._pull_totals()
        assert np.array_equal(m.sum_X, orc.sum_X) and np.array_equal(m.sum_sq_X, orc.sum_sq_X)
        assert np.array_equal(m.sum_Y, orc.sum_Y) and np.array_equal(m.sum_sq_Y, orc.sum_sq_Y)
        assert m.sum_w == orc.sum_w and m.num_nonzero_w == orc.nnz_w
    # ---- folds: the same three phases for the fold sums, then statistics bit-exact against the oracle
    ld = int(hs[0]._lib.cvmx_ld(hs[0]._h))
    ftots = []
    for m, (r0, r1) in zip(hs, slabs):
        lib, h = m._lib, m._h
        loc_off, loc_idx = sharding.local_csr(offsets, indices, r0, r1)
        m._upload_csr(loc_off, loc_idx)
        _lib.check(lib.cvmx_set_weight_folds(h, offsets.ctypes.data_as(C.c_void_p), indices.ctypes.data_as(C.c_void_p), P), h)
        tot = torch.zeros((P * 4 * ld,), dtype=torch.float64, device=dev)
        app = C.c_int32(0)
        _lib.check(lib.cvmx_slab_scan_local(h, 0, P, None, 0, 0, vp(tot), C.byref(app)), h)
        assert app.value == 1
        m.sync()
        ftots.append(tot)
    for r, m in enumerate(hs):
        start = torch.zeros_like(ftots[0]) if r == 0 else torch.stack(ftots[:r]).sum(0)
        keep.append(start)
        torch.cuda.synchronize()
        _lib.check(m._lib.cvmx_slab_scan_prepare(m._h, 0, P, vp(start)), m._h)
    raw = torch.zeros((P, 2, ld), dtype=torch.float64, device=dev)
    torch.cuda.synchronize()
    for m in hs:
        _lib.check(m._lib.cvmx_slab_fold_sums(m._h, 0, P, vp(raw)), m._h)
        m.sync()
    m0 = hs[0]
    _lib.check(m0._lib.cvmx_slab_finalize_stats(m0._h, 0, P, vp(raw)), m0._h)
    m0.sync()
    # raw fold sums against numpy's sequential sums, bit for bit
    rawh = raw.cpu().numpy()
    for pos, key in enumerate(part.folds_dict):
        val = part.get_validation_indices(key)
        Z = np.concatenate([X[val], Y[val]], axis=1)
        WZ = w[val][:, None] * Z
        assert np.array_equal(rawh[pos, 0, :K + M], np.sum(WZ, axis=0)), pos
        assert np.array_equal(rawh[pos, 1, :K + M], np.sum(WZ * Z, axis=0)), pos
