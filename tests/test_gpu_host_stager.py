"""
Pageable host inputs (-m gpu): ordinary numpy arrays are uploaded through the library's page-locked ring filled by
worker threads (csrc/host_stager.h), page-locked arrays go straight to the DMA engine, and CVMX_HOST_STAGER=0 leaves
pageable memory to the driver.  All three must leave the same bits on the device - fit totals, moment sums and fold
results identical - for the chunk-pipelined fit (>= 64 MB), the fused fit + folds, and the streaming fit in row blocks.
"""

import os

import numpy as np
import pytest

from cvmatrix_oracle import make_inputs

pytestmark = pytest.mark.gpu


def _pinned_copy(a):
    import torch

    t = torch.empty(a.shape, dtype=torch.from_numpy(a).dtype, pin_memory=True)
    t.numpy()[...] = a
    return t


def _fit(X, Y, w, stager, folds=None, blocks=0):
    from cvmatrix_b200 import CVMatrix, Partitioner

    old = os.environ.get("CVMX_HOST_STAGER")
    os.environ["CVMX_HOST_STAGER"] = "1" if stager else "0"
    try:
        m = CVMatrix(copy=False)          # the switch is read when the handle is created
    finally:
        if old is None:
            del os.environ["CVMX_HOST_STAGER"]
        else:
            os.environ["CVMX_HOST_STAGER"] = old
    if blocks:
        N, K = X.shape
        m.fit_begin(N, K, Y.shape[1], weighted=True, max_block_rows=blocks)
        for b0 in range(0, N, blocks):
            m.fit_rows(b0, X[b0:b0 + blocks], Y[b0:b0 + blocks], w[b0:b0 + blocks])
        m.fit_end()
    elif folds is not None:
        m.fit(X, Y, w, folds=Partitioner(folds))
    else:
        m.fit(X, Y, w)
    return m


def _same(a, b):
    for name in ("XTX", "XTY", "sum_X", "sum_Y", "sum_sq_X", "sum_sq_Y"):
        assert np.array_equal(getattr(a, name), getattr(b, name)), name
    assert a.sum_w == b.sum_w and a.num_nonzero_w == b.num_nonzero_w


def test_pageable_upload_matches_pinned_and_driver_path():
    # 120k x 100 float64 = 96 MB: the chunk-pipelined upload (>= 64 MB), several ring pieces per chunk, a ragged last piece
    X, Y, w, folds = make_inputs(120_001, 100, 9, 4, seed=31)
    Xp, Yp, wp = _pinned_copy(X), _pinned_copy(Y), _pinned_copy(w)
    a = _fit(X, Y, w, stager=True)
    b = _fit(Xp.numpy(), Yp.numpy(), wp.numpy(), stager=True)     # page-locked: no staging
    c = _fit(X, Y, w, stager=False)                                # the driver's bounce buffer
    _same(a, b)
    _same(a, c)
    # fused fit + folds and the batched fold path on top of it
    fa, fb = _fit(X, Y, w, True, folds=folds), _fit(Xp.numpy(), Yp.numpy(), wp.numpy(), True, folds=folds)
    _same(fa, fb)
    oa, ob = fa.training_batch(), fb.training_batch()
    for key in ("XTX", "XTY", "X_mean", "X_std", "Y_mean", "Y_std"):
        assert np.array_equal(oa[key], ob[key]), key


def test_pageable_row_blocks():
    X, Y, w, _ = make_inputs(70_000, 96, 5, 2, seed=32)            # 54 MB in 16k-row blocks of 12 MB: staged per block
    a = _fit(X, Y, w, stager=True, blocks=16_384)
    b = _fit(X, Y, w, stager=False, blocks=16_384)
    _same(a, b)
    c = _fit(X, Y, w, stager=True)
    assert np.array_equal(a.sum_X, c.sum_X) and a.sum_w == c.sum_w
