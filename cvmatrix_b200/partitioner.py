"""
Partitioner: fold labels -> validation index sets, as one CSR.

Drop-in for ``cvmatrix.partitioner.Partitioner`` (reference cvmatrix/partitioner.py:22-107,
Algorithm 1 of Engstrøm & Jensen 2025).  Same constructor, same ``folds_dict`` (a real
insertion-ordered dict, key order = first appearance, values = ascending int64 index
arrays), same ``get_validation_indices`` and the same ``ValueError(f"Fold {fold} not found.")``.

What is new: all index sets live in ONE contiguous int64 buffer (``indices``) delimited by
``offsets`` — the CSR that ``CVMatrix.set_folds`` uploads to the device once, so a batched
call can run every fold without any per-fold host work.  The dict values are views into
that buffer.  Numeric label arrays are partitioned with a stable sort instead of the
reference's Python loop (0.2-1.5 s at N = 1M); arbitrary hashables use
the same dict-insertion semantics as the reference.
"""

from __future__ import annotations

from collections.abc import Hashable
from typing import Iterable

import numpy as np
import numpy.typing as npt


class Partitioner:
    """
    Parameters
    ----------
    folds : Iterable of Hashable with N elements
        Each unique value defines one fold; the indices of the samples carrying it are the
        validation set of that fold.

    Attributes
    ----------
    folds_dict : dict[Hashable, npt.NDArray[np.int_]]
        key -> ascending int64 indices (views into ``indices``).
    keys : list
        Fold keys in first-appearance order (``list(folds_dict)``).
    offsets : int64 array of P + 1 entries, indices : int64 array of N entries
        CSR of the validation sets, fold ``k`` owns ``indices[offsets[k]:offsets[k+1]]``.
    """

    def __init__(self, folds: Iterable[Hashable]) -> None:
        self.folds_dict: dict[Hashable, npt.NDArray[np.int_]] = {}
        self._init_folds_dict(folds)

    def get_validation_indices(self, fold: Hashable) -> npt.NDArray[np.int_]:
        try:
            return self.folds_dict[fold]
        except KeyError as e:
            raise ValueError(f"Fold {fold} not found.") from e

    # ------------------------------------------------------------------------------------------
    def fold_position(self, fold: Hashable) -> int:
        """Position of a fold key in the CSR (row of ``offsets``)."""
        try:
            return self._pos[fold]
        except KeyError as e:
            raise ValueError(f"Fold {fold} not found.") from e

    @property
    def n_folds(self) -> int:
        return len(self.keys)

    def csr(self):
        """(offsets[P+1], indices[N]) as contiguous int64 arrays."""
        return self.offsets, self.indices

    # ------------------------------------------------------------------------------------------
    def _init_folds_dict(self, folds: Iterable[Hashable]) -> None:
        arr = folds if isinstance(folds, np.ndarray) else None
        if (
            arr is not None
            and arr.ndim == 1
            and arr.dtype.kind in "iufb"
            and not (arr.dtype.kind == "f" and np.isnan(arr).any())
        ):
            keys, offsets, indices = self._partition_numeric(arr)
        else:
            keys, offsets, indices = self._partition_hashable(folds)
        self.keys = keys
        self.offsets = offsets
        self.indices = indices
        self.folds_dict = {k: indices[offsets[p]:offsets[p + 1]] for p, k in enumerate(keys)}
        self._pos = {k: p for p, k in enumerate(keys)}

    @staticmethod
    def _partition_numeric(arr: np.ndarray):
        n = arr.shape[0]
        if n == 0:
            return [], np.zeros(1, dtype=np.int64), np.zeros(0, dtype=np.int64)
        group = first = None
        if arr.dtype.kind in "iub":
            # integer labels in a small range: counting pass instead of a comparison sort
            lo, hi = int(arr.min()), int(arr.max())
            span = hi - lo + 1
            if span <= max(4 * n, 1 << 16) and span <= (1 << 26):
                native = Partitioner._partition_native(arr, lo, span) if (4 * span <= n or span <= (1 << 16)) else None
                if native is not None:
                    return native
                rel = (arr.astype(np.int64, copy=False) - lo) if arr.dtype.kind != "b" else arr.astype(np.int64)
                first_of = np.full(span, n, dtype=np.int64)
                first_of[rel[::-1]] = np.arange(n - 1, -1, -1, dtype=np.int64)   # last write wins = first occurrence
                present = np.flatnonzero(first_of < n)
                order = present[np.argsort(first_of[present], kind="stable")]    # label values by first appearance
                rank = np.empty(span, dtype=np.int64)
                rank[order] = np.arange(order.size)
                group = rank[rel]
                first = first_of[order]
                n_groups = order.size
        if group is None:
            _, first_u, inverse = np.unique(arr, return_index=True, return_inverse=True)
            inverse = inverse.reshape(-1)
            order = np.argsort(first_u, kind="stable")         # unique values by first appearance
            rank = np.empty_like(order)
            rank[order] = np.arange(order.size)
            group = rank[inverse]                              # fold position of every row
            first = first_u[order]
            n_groups = order.size
        counts = np.bincount(group, minlength=n_groups)
        offsets = np.zeros(n_groups + 1, dtype=np.int64)
        np.cumsum(counts, out=offsets[1:])
        if n_groups == n:
            indices = first.astype(np.int64, copy=True)        # every label unique (leave-one-out): no sort needed
        else:
            # numpy's stable sort is a radix sort for 16-bit keys
            small = np.uint16 if n_groups <= (1 << 16) else (np.int32 if n_groups < 2**31 else np.int64)
            indices = np.argsort(group.astype(small, copy=False), kind="stable").astype(np.int64, copy=False)
        keys = [arr[i] for i in first]                         # numpy scalars, like iterating the array
        return keys, offsets, np.ascontiguousarray(indices)

    @staticmethod
    def _partition_native(arr: np.ndarray, lo: int, span: int):
        """Two counting passes in the C library (host code of libcvmx.so); None if the library is not built."""
        try:
            from . import _lib

            lib = _lib.load()
        except (ImportError, OSError):
            return None
        import ctypes as C

        n = arr.shape[0]
        labels = np.ascontiguousarray(arr, dtype=np.int64)
        scratch = np.empty(2 * span, dtype=np.int64)
        first = np.empty(min(span, n), dtype=np.int64)
        offsets = np.empty(min(span, n) + 1, dtype=np.int64)
        indices = np.empty(n, dtype=np.int64)
        p = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
        nf = lib.cvmx_partition_labels(p(labels), n, lo, span, p(scratch), p(first), p(offsets), p(indices))
        if nf < 0:
            return None
        first = first[:nf]
        keys = [arr[i] for i in first]   # numpy scalars, like iterating the array
        return keys, offsets[: nf + 1].copy(), indices

    @staticmethod
    def _partition_hashable(folds: Iterable[Hashable]):
        buckets: dict[Hashable, list[int]] = {}
        for i, label in enumerate(folds):
            rows = buckets.get(label)
            if rows is None:
                buckets[label] = [i]
            else:
                rows.append(i)
        keys = list(buckets)
        offsets = np.zeros(len(keys) + 1, dtype=np.int64)
        if keys:
            np.cumsum([len(buckets[k]) for k in keys], out=offsets[1:])
        indices = np.empty(int(offsets[-1]), dtype=np.int64)
        for p, k in enumerate(keys):
            indices[offsets[p]:offsets[p + 1]] = buckets[k]
        return keys, offsets, indices
