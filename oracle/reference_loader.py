"""
TEST INFRASTRUCTURE — NOT PRODUCT CODE.

Hands out the CPU implementation that bench.py's `cpu_baseline` / `--impl reference` legs and the parity report
time and compare against:

* kind "reference": the UNMODIFIED reference package staged by `make -C oracle ref` into oracle/_ref/ (a plain copy
  of /root/reference/cvmatrix; oracle/_ref/SHA256SUMS lists the file hashes).  Used whenever it is present.
* kind "port": oracle/cvmatrix_oracle.py with order="numpy" (the same numpy calls; pinned bit-identical to the
  reference by oracle/check_against_reference.py) when oracle/_ref is absent.

Only tests/, __graft_entry__.smoke() and bench.py may import this module.
"""

from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")


def host_threads() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:  # pragma: no cover
        return os.cpu_count() or 1


def use_all_host_threads() -> int:
    """BLAS thread pools at the full core count (torchrun exports OMP_NUM_THREADS=1 to its workers)."""
    n = host_threads()
    try:
        from threadpoolctl import threadpool_limits

        threadpool_limits(limits=n)
    except Exception:  # pragma: no cover
        pass
    return n


def blas_threads() -> int:
    try:
        from threadpoolctl import threadpool_info

        return max([int(p.get("num_threads", 1)) for p in threadpool_info() if p.get("user_api") == "blas"] or [1])
    except Exception:  # pragma: no cover
        return 1


def load():
    """-> (CVMatrix class, Partitioner class, kind)"""
    if os.path.isfile(os.path.join(REF_DIR, "cvmatrix", "cvmatrix.py")):
        if REF_DIR not in sys.path:
            sys.path.insert(0, REF_DIR)
        from cvmatrix import CVMatrix, Partitioner  # the staged, unmodified reference

        return CVMatrix, Partitioner, "reference"
    if HERE not in sys.path:
        sys.path.insert(0, HERE)
    from cvmatrix_oracle import OracleCVMatrix, OraclePartitioner

    class PortCVMatrix(OracleCVMatrix):
        def __init__(self, *a, **k):
            k.setdefault("order", "numpy")
            k.pop("backend", None)
            super().__init__(*a, **k)

    return PortCVMatrix, OraclePartitioner, "port"


def fold_outputs(model, val):
    """(XTX, XTY, (X_mean, X_std, Y_mean, Y_std)) of one validation set, reference call shape."""
    (XTX, XTY), stats = model.training_XTX_XTY(np.asarray(val))
    return XTX, XTY, stats
