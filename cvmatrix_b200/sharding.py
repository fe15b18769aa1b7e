"""
Pure (device-free) sharding arithmetic of the multi-GPU fold path.  The C library applies the same
rules inside cvmx_sharded_stats / cvmx_sharded_gram (cvmatrix_b200/csrc/cvmx_api.cu); keeping them here
lets the N > 1 logic be tested on CPU with a gloo process group.

The path shards two ways:
  * many folds  -> contiguous fold blocks per rank, no collective;
  * few folds   -> every fold's rows are split across ranks (float64 Gram partials reduced by the fold owner over
                   NVLink peer memory, or one all-reduce),
                   the sequential per-column moment chains are split by column group (one all-reduce of
                   the statistics rows, whose foreign entries are zero).
"""

from __future__ import annotations

GRAM_ROW_ALIGN = 16      # GBK: rows per pipeline stage of k_gram
MOMENT_GROUP_COLS = 32   # MOM_COLS: columns per moment-chain CTA


def fold_block(rank: int, world: int, f0: int, f1: int):
    """Contiguous block of folds [f0, f1) owned by `rank`."""
    n = f1 - f0
    return f0 + rank * n // world, f0 + (rank + 1) * n // world


def use_row_sharding(n_folds: int, world: int) -> bool:
    """Few large folds cannot keep `world` GPUs busy by fold ownership alone."""
    return world > 1 and n_folds < 4 * world


def row_shard(n_rows: int, shard: int, n_shards: int):
    """[begin, end) positions of `shard` inside a fold's index list (stage-aligned, last shard takes the rest)."""
    per = -(-n_rows // n_shards)
    per = -(-per // GRAM_ROW_ALIGN) * GRAM_ROW_ALIGN
    return min(n_rows, shard * per), min(n_rows, (shard + 1) * per)


def column_groups(ld: int, shard: int, n_shards: int, group_cols: int = MOMENT_GROUP_COLS):
    """Column groups (of `group_cols` columns) whose moment chains `shard` computes."""
    n_groups = -(-ld // group_cols)
    return list(range(shard, n_groups, n_shards))


def slab_rows(rank: int, world: int, n_rows: int):
    """Row slab [begin, end) of `rank` when the ROWS of the data set are sharded across ranks (BASELINE config 5)."""
    return rank * n_rows // world, (rank + 1) * n_rows // world


def weighted_slab_bounds(weights, n_rows: int, align: int = GRAM_ROW_ALIGN):
    """Row-slab boundaries [b_0 = 0, ..., b_W = n_rows] with slab r proportional to weights[r] (e.g. the host->device copy
    rate each rank measured: on a box whose GPUs do not reach host memory equally fast, equal slabs make every fit wait
    for the slowest link).  Boundaries are multiples of `align` rows; every rank keeps at least one aligned block when
    the data set allows it."""
    w = [max(float(x), 0.0) for x in weights]
    world = len(w)
    total = sum(w)
    if total <= 0.0:
        w, total = [1.0] * world, float(world)
    bounds, acc = [0], 0.0
    for r in range(world - 1):
        acc += w[r]
        b = int(round(n_rows * acc / total / align)) * align
        lo = bounds[-1] + (align if n_rows >= world * align else 0)
        hi = n_rows - (world - 1 - r) * (align if n_rows >= world * align else 0)
        bounds.append(max(lo, min(b, hi)))
    bounds.append(n_rows)
    return bounds


def local_csr(offsets, indices, row0: int, row1: int, assume_sorted: bool = False):
    """The part of a global CSR of validation sets (ascending row numbers inside every fold) that falls into the row slab
    [row0, row1), renumbered from 0: (local offsets, local indices).  Raises if a fold is not ascending - the chained
    column sums rely on a fold's rows on rank r all preceding those on rank r + 1 (``assume_sorted``: the caller vouches
    for it - a Partitioner's index sets are ascending by construction - and the check, a pass over all N indices, is
    skipped)."""
    import numpy as np

    offsets = np.asarray(offsets, dtype=np.int64)
    indices = np.asarray(indices, dtype=np.int64)
    P = offsets.size - 1
    loc = np.zeros(P + 1, np.int64)
    if not assume_sorted and indices.size > 1:
        step = np.diff(indices) <= 0
        step[offsets[1:-1][(offsets[1:-1] > 0) & (offsets[1:-1] < indices.size)] - 1] = False   # fold boundaries may step down
        if step.any():
            bad = int(np.searchsorted(offsets, np.flatnonzero(step)[0], side="right") - 1)
            raise ValueError(f"fold {bad}: row-slab mode needs strictly ascending validation indices")
    if P <= 4096:
        lo = np.empty(P, np.int64)
        hi = np.empty(P, np.int64)
        for f in range(P):
            idx = indices[offsets[f]:offsets[f + 1]]
            lo[f], hi[f] = np.searchsorted(idx, row0), np.searchsorted(idx, row1)
        np.cumsum(hi - lo, out=loc[1:])
        out = np.empty(int(loc[P]), np.int64)
        for f in range(P):
            np.subtract(indices[offsets[f] + lo[f]:offsets[f] + hi[f]], row0, out=out[loc[f]:loc[f + 1]])
        return loc, out
    inside = (indices >= row0) & (indices < row1)            # many folds: one vectorised pass, order preserved
    csum = np.concatenate([[0], np.cumsum(inside)])
    loc[:] = csum[offsets]
    return loc, indices[inside] - row0
