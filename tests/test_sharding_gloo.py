"""
CPU tests of the N > 1 host logic with a world_size-2 gloo process group: fold ownership covers every fold
exactly once; row shards of a fold tile its index list; the all-reduce assembly used by the row-sharded mode
(zero-filled foreign column groups for the statistics, summed float64 Gram partials) reproduces the unsharded
result.  The per-rank compute stand-in is the numpy oracle (test infrastructure), not the CUDA library.
"""

import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from cvmatrix_b200 import sharding

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_partition_rules():
    for world in (1, 2, 3, 8):
        for f0, f1 in ((0, 5), (0, 1000), (3, 4), (10, 10)):
            blocks = [sharding.fold_block(r, world, f0, f1) for r in range(world)]
            assert blocks[0][0] == f0 and blocks[-1][1] == f1
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
        for n in (0, 1, 15, 16, 17, 200_000, 200_003):
            shards = [sharding.row_shard(n, s, world) for s in range(world)]
            assert shards[0][0] == 0 and shards[-1][1] == n
            assert all(shards[i][1] == shards[i + 1][0] for i in range(world - 1))
            assert all(b % sharding.GRAM_ROW_ALIGN == 0 for b, _ in shards if b < n)
        groups = sorted(g for s in range(world) for g in sharding.column_groups(512, s, world))
        assert groups == list(range(16))
    assert sharding.use_row_sharding(5, 8) and not sharding.use_row_sharding(1000, 8) and not sharding.use_row_sharding(5, 1)


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    from cvmatrix_oracle import OracleCVMatrix, make_inputs, colsum

    dist.init_process_group("gloo", rank=rank, world_size=world)
    X, Y, w, folds = make_inputs(4000, 40, 3, 3, seed=5)
    Z = np.hstack([X, Y])
    ld = 64
    orc = OracleCVMatrix(order="numpy")
    orc.fit(X, Y, w)
    ok = True
    for f in range(3):
        val = np.flatnonzero(folds == f)
        # row-sharded raw Gram partial of this rank, summed by all-reduce
        b, e = sharding.row_shard(val.size, rank, world)
        rows = val[b:e]
        part = (X[rows] * w[rows, None]).T @ Z[rows]
        g = torch.from_numpy(part.copy())
        dist.all_reduce(g)
        full = (X[val] * w[val, None]).T @ Z[val]
        ok &= np.allclose(g.numpy(), full, rtol=1e-13, atol=0)
        # column-sharded sequential sums: foreign groups zero, all-reduce(sum) assembles them exactly
        sums = np.zeros(ld)
        for grp in sharding.column_groups(ld, rank, world):
            c0, c1 = grp * 32, min((grp + 1) * 32, Z.shape[1])
            if c0 < c1:
                sums[c0:c1] = colsum((Z[val] * w[val, None])[:, c0:c1], "explicit") if c1 - c0 > 1 else (Z[val, c0] * w[val]).sum()
        s = torch.from_numpy(sums)
        dist.all_reduce(s)
        ref = np.sum(Z[val] * w[val, None], axis=0)
        ok &= np.array_equal(s.numpy()[: Z.shape[1]], ref)
    # fold ownership: every fold finished by exactly one rank
    owned = torch.zeros(3, dtype=torch.int64)
    o0, o1 = sharding.fold_block(rank, world, 0, 3)
    owned[o0:o1] += 1
    dist.all_reduce(owned)
    ok &= bool((owned == 1).all())
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_row_sharded_assembly_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]


def test_row_slab_csr_covers_every_fold_once():
    """Row-slab mode: the local CSRs of all ranks, shifted back by their slab starts, reassemble every fold in order."""
    import numpy as np

    from cvmatrix_b200 import sharding

    rng = np.random.default_rng(0)
    N, P, world = 1003, 7, 3
    labels = rng.integers(0, P, N)
    order = np.argsort(labels, kind="stable")
    offsets = np.concatenate([[0], np.cumsum(np.bincount(labels, minlength=P))]).astype(np.int64)
    got = [[] for _ in range(P)]
    covered = 0
    for r in range(world):
        r0, r1 = sharding.slab_rows(r, world, N)
        covered += r1 - r0
        loc_off, loc_idx = sharding.local_csr(offsets, order, r0, r1)
        assert loc_idx.size == 0 or (loc_idx.min() >= 0 and loc_idx.max() < r1 - r0)
        for f in range(P):
            got[f].append(loc_idx[loc_off[f]:loc_off[f + 1]] + r0)
    assert covered == N
    for f in range(P):
        assert np.array_equal(np.concatenate(got[f]), order[offsets[f]:offsets[f + 1]])
    import pytest

    with pytest.raises(ValueError, match="ascending"):
        sharding.local_csr(np.array([0, 3]), np.array([5, 2, 9]), 0, 10)


def test_weighted_slab_bounds():
    from cvmatrix_b200 import sharding

    b = sharding.weighted_slab_bounds([23, 23, 23, 23, 35, 35, 35, 35], 1_000_000)
    assert b[0] == 0 and b[-1] == 1_000_000 and len(b) == 9 and all(x % 16 == 0 for x in b[:-1])
    sizes = [y - x for x, y in zip(b, b[1:])]
    assert all(s > 0 for s in sizes) and abs(sizes[4] / sizes[0] - 35 / 23) < 0.01
    assert sharding.weighted_slab_bounds([1, 1], 100) == [0, 48, 100]
    assert sharding.weighted_slab_bounds([0, 0, 0], 50) == [0, 16, 32, 50]           # no rates: equal slabs
    assert sharding.weighted_slab_bounds([1, 0, 1, 5], 64) == [0, 16, 32, 48, 64]    # every rank keeps a block
    assert sharding.weighted_slab_bounds([3.0], 10) == [0, 10]


def test_local_csr_paths_agree():
    """Few folds (per-fold binary search) and many folds (one vectorised pass) give the same local CSR; unsorted folds are
    refused unless the caller vouches for them."""
    import numpy as np
    import pytest

    from cvmatrix_b200 import sharding

    rng = np.random.default_rng(5)
    N = 20_000
    for P in (3, 5000):
        labels = rng.integers(0, P, size=N)
        order = np.argsort(labels, kind="stable").astype(np.int64)
        offsets = np.concatenate([[0], np.cumsum(np.bincount(labels, minlength=P))]).astype(np.int64)
        for r0, r1 in ((0, N), (3000, 9000), (N - 1, N), (500, 500)):
            loc, idx = sharding.local_csr(offsets, order, r0, r1)
            loc2, idx2 = sharding.local_csr(offsets, order, r0, r1, assume_sorted=True)
            assert np.array_equal(loc, loc2) and np.array_equal(idx, idx2)
            want = [order[offsets[f]:offsets[f + 1]] for f in range(P)]
            want = [v[(v >= r0) & (v < r1)] - r0 for v in want]
            assert np.array_equal(idx, np.concatenate(want)) and np.array_equal(np.diff(loc), [v.size for v in want])
    with pytest.raises(ValueError, match="fold 1"):
        sharding.local_csr(np.array([0, 2, 5]), np.array([0, 4, 3, 2, 9]), 0, 10)
    # a step down exactly at a fold boundary is fine
    loc, idx = sharding.local_csr(np.array([0, 2, 4]), np.array([5, 9, 0, 3]), 2, 8)
    assert loc.tolist() == [0, 1, 2] and idx.tolist() == [3, 1]
