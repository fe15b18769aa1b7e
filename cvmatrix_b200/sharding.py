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


def local_csr(offsets, indices, row0: int, row1: int):
    """The part of a global CSR of validation sets (ascending row numbers inside every fold) that falls into the row slab
    [row0, row1), renumbered from 0: (local offsets, local indices).  Raises if a fold is not ascending - the chained
    column sums rely on a fold's rows on rank r all preceding those on rank r + 1."""
    import numpy as np

    offsets = np.asarray(offsets, dtype=np.int64)
    indices = np.asarray(indices, dtype=np.int64)
    P = offsets.size - 1
    parts, loc = [], np.zeros(P + 1, np.int64)
    for f in range(P):
        idx = indices[offsets[f]:offsets[f + 1]]
        if idx.size > 1 and np.any(np.diff(idx) <= 0):
            raise ValueError(f"fold {f}: row-slab mode needs strictly ascending validation indices")
        lo, hi = np.searchsorted(idx, row0), np.searchsorted(idx, row1)
        parts.append(idx[lo:hi] - row0)
        loc[f + 1] = loc[f] + (hi - lo)
    return loc, (np.concatenate(parts) if parts else np.zeros(0, np.int64))
