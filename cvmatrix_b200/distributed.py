"""
Multi-GPU evaluation of fold batches: one process per GPU (torchrun), every rank holds a fitted CVMatrix on
the same data.  The row-sharded mode reduces over NVLink peer memory (symmetric memory: the fold owner sums its
peers' raw Grams inside the epilogue kernel); `torch.distributed` (NCCL) carries the fallback all-reduce and the
slab exchange of the sharded upload.
See cvmatrix_b200/sharding.py for the sharding rules and include/cvmx.h (cvmx_sharded_*) for the device side.
"""

from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np

from . import _lib, sharding
from .cvmatrix import CVMatrix


class _DevArray:
    """Zero-copy view of library-owned device memory for torch (``__cuda_array_interface__``)."""

    def __init__(self, ptr: int, count: int, typestr: str):
        self.__cuda_array_interface__ = {"shape": (count,), "typestr": typestr, "data": (ptr, False), "version": 3}


class ShardedFolds:
    """
    Runs folds ``[f0, f1)`` of ``cvm``'s device CSR across all ranks of ``group``.

    ``training_batch`` returns device tensors for the folds THIS rank owns (a contiguous block, see
    ``sharding.fold_block``): dict(fold_begin, fold_end, XTX, XTY, stats, scal, status).  Many folds: each rank
    simply evaluates its block.  Few folds: every rank computes the raw Gram of its row shard of every fold and the
    moment sums of its column groups; the owner of a fold sums its peers' buffers over NVLink inside the epilogue
    kernel (or one NCCL all-reduce assembles them); each rank then finishes its own folds.
    """

    def __init__(self, cvm: CVMatrix, group=None):
        import torch
        import torch.distributed as dist

        self.cvm, self.group, self.dist, self.torch = cvm, group, dist, torch
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        # profiling aid: run this rank's share of a `shards`-way row-sharded step without any collective
        self.emulate_shards = 0
        self.dev = torch.device("cuda", cvm.device)
        self.tdt = torch.float64 if np.dtype(cvm.dtype) == np.float64 else torch.float32
        self._gram: Optional["torch.Tensor"] = None
        # peer-memory reduction (float64, <= 8 ranks): the raw Grams live in symmetric memory mapped over NVLink and the
        # fold owner sums its peers' fragments inside the epilogue kernel - no all-reduce.  CVMX_PEER_REDUCE=0 disables.
        import os

        self.use_peers = (os.environ.get("CVMX_PEER_REDUCE", "1") != "0" and self.world > 1 and self.world <= 8
                          and self.tdt == torch.float64)
        self._symm = None        # (buffer, handle, elements per half, peer pointer arrays per half)
        self._step = 0

    @property
    def want(self) -> int:
        """XTX always; XTY only when the model was fitted with Y (the reference supports X-only use: training_XTX)."""
        return _lib.WANT_XTX | (_lib.WANT_XTY if (self.cvm.M or 0) else 0)

    def alloc_outputs(self, n_folds: int):
        t, K, M = self.torch, self.cvm.K, self.cvm.M or 0
        return dict(
            XTX=t.empty((n_folds, K, K), dtype=self.tdt, device=self.dev),
            XTY=t.empty((n_folds, K, M), dtype=self.tdt, device=self.dev) if M else None,
            stats=t.empty((n_folds, 2, K + M), dtype=self.tdt, device=self.dev),
            scal=t.empty((n_folds, 2), dtype=self.tdt, device=self.dev),
            status=t.empty((n_folds,), dtype=t.int32, device=self.dev),
        )

    def _setup_symm(self, half_elems: int):
        """Symmetric-memory buffer of 2 x half_elems float64 shared by all ranks of the group; returns None (on every rank)
        unless every rank succeeded, so that all ranks take the same path."""
        t, dist = self.torch, self.dist
        ok, res = 1, None
        try:
            import torch.distributed._symmetric_memory as symm_mem

            grp = self.group if self.group is not None else dist.group.WORLD
            buf = symm_mem.empty(2 * half_elems, dtype=t.float64, device=self.dev)
            hdl = symm_mem.rendezvous(buf, grp)
            ptrs = []
            for half in range(2):
                arr = (C.c_void_p * self.world)(*[int(p) + half * half_elems * 8 for p in hdl.buffer_ptrs])
                ptrs.append(arr)
            res = (buf, hdl, half_elems, ptrs)
        except Exception:   # pragma: no cover - depends on the box (driver support for shareable allocations)
            ok = 0
        flag = t.tensor([ok], dtype=t.int32, device=self.dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.group)
        if int(flag.item()) == 0:
            self.use_peers = False
            return None
        return res

    def training_batch(self, f0: int, f1: int, out: Optional[dict] = None, row_sharded: Optional[bool] = None):
        """Must be called on a non-default torch stream that the handle has been bound to (cvmx_set_stream), so
        that library kernels and NCCL collectives are ordered on one stream."""
        t, cvm = self.torch, self.cvm
        lib, h = cvm._lib, cvm._h
        if row_sharded is None:
            row_sharded = sharding.use_row_sharding(f1 - f0, self.world)
        o0, o1 = sharding.fold_block(self.rank, self.world, f0, f1)
        if out is None:
            out = self.alloc_outputs(max(o1 - o0, 1))
        vp = lambda x: None if x is None else C.c_void_p(x.data_ptr())  # noqa: E731
        want = self.want
        if not row_sharded:
            if o1 > o0:
                _lib.check(lib.cvmx_training_batch(h, o0, o1, want, vp(out["XTX"]), vp(out["XTY"]), vp(out["stats"]), vp(out["scal"]),
                                                   vp(out["status"]), _lib.DEVICE), h)
            return dict(out, fold_begin=o0, fold_end=o1)
        # ---- few large folds: rows of every fold split across ranks --------------------------------------
        sp, sc = C.c_void_p(), C.c_int64()
        shards = self.emulate_shards or self.world
        _lib.check(lib.cvmx_sharded_stats(h, f0, f1, self.rank, shards, C.byref(sp), C.byref(sc)), h)
        n = lib.cvmx_sharded_gram_count(h, f0, f1, want)
        half_elems = n + sc.value
        if self.use_peers and self.world > 1 and (self._symm is None or self._symm[2] < half_elems):
            self._symm = self._setup_symm(half_elems)
        if self._symm is not None and self.world > 1:
            buf, hdl, cap, ptrs = self._symm
            half = self._step & 1            # two halves alternate: a peer may still read step i while step i + 1 is written
            self._step += 1
            gram = buf[half * cap: half * cap + n]
            _lib.check(lib.cvmx_sharded_gram(h, f0, f1, want, self.rank, shards, vp(gram)), h)
            _lib.check(lib.cvmx_sharded_stats_wait(h), h)
            stats = t.as_tensor(_DevArray(sp.value, sc.value, "<f8"), device=self.dev)
            buf[half * cap + n: half * cap + n + sc.value].copy_(stats)
            hdl.barrier(channel=0)           # every rank's Grams and statistics rows are complete and visible
            _lib.check(lib.cvmx_sharded_finish_peers(h, f0, f1, o0, o1, want, ptrs[half], self.world, n, vp(out["XTX"]), vp(out["XTY"]),
                                                     vp(out["stats"]), vp(out["scal"]), vp(out["status"])), h)
            return dict(out, fold_begin=o0, fold_end=o1)
        # one buffer for both reductions: [raw Grams | statistics rows widened to float64] -> ONE all-reduce
        if self._gram is None or self._gram.numel() < n + sc.value:
            self._gram = t.empty((n + sc.value,), dtype=t.float64, device=self.dev)
        gram = self._gram[:n]
        _lib.check(lib.cvmx_sharded_gram(h, f0, f1, want, self.rank, shards, vp(gram)), h)
        _lib.check(lib.cvmx_sharded_stats_wait(h), h)   # the chains ran on a side stream beside the Gram kernel
        if self.world > 1:
            stats = t.as_tensor(_DevArray(sp.value, sc.value, "<f8" if self.tdt == t.float64 else "<f4"), device=self.dev)
            tail = self._gram[n:n + sc.value]
            tail.copy_(stats)
            self.dist.all_reduce(self._gram[:n + sc.value], group=self.group)
            stats.copy_(tail)   # foreign entries were zero: the sum is exact in either dtype
        _lib.check(lib.cvmx_sharded_finish(h, f0, o0, o1, want, vp(gram), vp(out["XTX"]), vp(out["XTY"]), vp(out["stats"]),
                                           vp(out["scal"]), vp(out["status"])), h)
        return dict(out, fold_begin=o0, fold_end=o1)


def fit_row_sharded(cvm: CVMatrix, X, Y=None, weights=None, group=None) -> None:
    """
    ``cvm.fit`` with the Gram pass (X^T W [X|Y], the 2 N K (K+M) flops of fit) split by rows across the ranks of
    ``group``: every rank uploads the data (it needs all rows for its folds anyway), contracts only its row slab and
    the partial totals are combined with ONE all-reduce over NVLink (K x ld elements: 2 MB at K = 500).  The moment
    sums keep numpy's sequential order, so every rank computes them over all rows.  float64 partial sums are added
    by NCCL in ring / tree order, so the totals can differ from the single-GPU ones in the last bits.
    """
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    n = np.asarray(X).shape[0]
    r0, r1 = sharding.fold_block(rank, world, 0, n)
    cvm.fit(X, Y, weights, _gram_rows=(r0, r1))
    if world == 1:
        return
    lib, h = cvm._lib, cvm._h
    ptr, count, ld = C.c_void_p(), C.c_int64(), C.c_int64()
    _lib.check(lib.cvmx_totals_ptr(h, C.byref(ptr), C.byref(count), C.byref(ld)), h)
    f64 = np.dtype(cvm.dtype) == np.float64
    tot = torch.as_tensor(_DevArray(ptr.value, count.value, "<f8" if f64 else "<f4"), device=torch.device("cuda", cvm.device))
    _lib.check(lib.cvmx_sync(h), h)
    dist.all_reduce(tot, group=group)
    torch.cuda.synchronize(cvm.device)
    _lib.check(lib.cvmx_commit_totals(h), h)
    cvm._pull_totals()


def fit_sharded_upload(cvm: CVMatrix, X, Y=None, weights=None, group=None, block_rows: int = 32768) -> None:
    """
    ``cvm.fit`` with the host->device upload AND the Gram pass split by rows across the ranks of ``group``: rank r
    copies only its own row slab over PCIe (``cvmx_fit_rows``; the slab's Gram runs behind the copy), the slabs are
    then exchanged over NVLink (one NCCL broadcast per rank, straight into ``cvmx_data_ptr()``), the numpy-order
    column sums are evaluated per column group (``cvmx_fit_end(rank, world)``: the binade scan) and three
    all-reduces assemble the totals and the two moment rows.  Every rank ends up with the same fitted state as
    ``cvm.fit(X, Y, weights)``, having moved 1 / world of the bytes over its PCIe link.
    """
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    X = cvm._init_mat(X)
    Y = cvm._init_mat(Y) if Y is not None else None
    w = cvm._init_mat(weights).reshape(-1) if weights is not None else None
    N, K = X.shape
    M = Y.shape[1] if Y is not None else 0
    cvm.fit_begin(N, K, M, weighted=w is not None, max_block_rows=block_rows)
    r0, r1 = sharding.fold_block(rank, world, 0, N)
    for b0 in range(r0, r1, block_rows):
        b1 = min(r1, b0 + block_rows)
        cvm.fit_rows(b0, X[b0:b1], None if Y is None else Y[b0:b1], None if w is None else w[b0:b1], gram=True)
    lib, h = cvm._lib, cvm._h
    _lib.check(lib.cvmx_sync(h), h)
    f64 = np.dtype(cvm.dtype) == np.float64
    ts = "<f8" if f64 else "<f4"
    dev = torch.device("cuda", cvm.device)
    if world > 1:
        zp, wp, ld = C.c_void_p(), C.c_void_p(), C.c_int64()
        _lib.check(lib.cvmx_data_ptr(h, C.byref(zp), C.byref(wp), C.byref(ld)), h)
        Z = torch.as_tensor(_DevArray(zp.value, N * ld.value, ts), device=dev)
        wv = torch.as_tensor(_DevArray(wp.value, N, ts), device=dev)
        for r in range(world):
            s0, s1 = sharding.fold_block(r, world, 0, N)
            if s1 > s0:
                src = dist.get_global_rank(group, r) if group is not None else r
                dist.broadcast(Z[s0 * ld.value: s1 * ld.value], src=src, group=group)
                if w is not None:
                    dist.broadcast(wv[s0:s1], src=src, group=group)
        torch.cuda.synchronize(cvm.device)
    cvm.fit_end(rank, world, pull=False)
    if world > 1:
        tp, cnt, ldt = C.c_void_p(), C.c_int64(), C.c_int64()
        _lib.check(lib.cvmx_totals_ptr(h, C.byref(tp), C.byref(cnt), C.byref(ldt)), h)
        sp, qp, mc = C.c_void_p(), C.c_void_p(), C.c_int64()
        _lib.check(lib.cvmx_moments_ptr(h, C.byref(sp), C.byref(qp), C.byref(mc)), h)
        for ptr, count in ((tp, cnt.value), (sp, mc.value), (qp, mc.value)):
            dist.all_reduce(torch.as_tensor(_DevArray(ptr.value, count, ts), device=dev), group=group)
        torch.cuda.synchronize(cvm.device)
        _lib.check(lib.cvmx_commit_totals(h), h)
    cvm._pull_totals()
    cvm.X, cvm.Y = X, Y
    cvm.weights = None if w is None else w.reshape(-1, 1)


def upload_balanced_bounds(N: int, sample, device, group=None, reps: int = 3):
    """Slab boundaries proportional to the host->device copy rate every rank reaches WHILE ALL RANKS COPY AT ONCE.

    On an 8-GPU box the links to host memory are not equal once they are all busy (measured on this pool's B200 boxes,
    0.5 GB per rank from pinned memory: 23 GB/s on four GPUs, 35 GB/s on the other four, against 55 GB/s for one GPU
    alone), and a row-sharded fit is as slow as its slowest upload.  `sample`: a pinned host tensor of this rank (>= 64 MB;
    a slice of the data itself will do).  Collective: every rank of `group` must call it.  Returns a list of W + 1 rows."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return [0, int(N)]
    dev = torch.device("cuda", device) if not isinstance(device, torch.device) else device
    buf = torch.empty(sample.shape, dtype=sample.dtype, device=dev)
    buf.copy_(sample, non_blocking=True)                      # warm-up (page tables, first-touch)
    torch.cuda.synchronize(dev)
    dist.barrier(group=group)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        buf.copy_(sample, non_blocking=True)
    e1.record()
    torch.cuda.synchronize(dev)
    rate = torch.tensor([sample.numel() * sample.element_size() * reps / max(e0.elapsed_time(e1), 1e-3)], dtype=torch.float64, device=dev)
    rates = torch.empty((world,), dtype=torch.float64, device=dev)
    dist.all_gather_into_tensor(rates, rate, group=group)
    return sharding.weighted_slab_bounds(rates.tolist(), int(N))


# ---------------------------------------------------------------------------------------------------------------------
# Row-slab mode (BASELINE config 5): the ROWS of the data set are sharded across the ranks instead of replicated.
# ---------------------------------------------------------------------------------------------------------------------
class RowSlabFolds:
    """
    A data set whose rows are sharded across the ranks of ``group`` (rank r holds rows ``sharding.slab_rows(r, world, N)``):
    fit and fold batches with every reduction over rows cut at the slab boundaries (include/cvmx.h, "Row-slab mode").

        rs = RowSlabFolds(cvm, N, K, M, weights_global)     # cvm: a fresh CVMatrix on this rank's device
        rs.fit(blocks)                                       # blocks: iterable of (global_row0, X_block, Y_block) inside the slab
        rs.set_folds(partitioner)                            # validation sets as GLOBAL row numbers
        out = rs.training_batch()                            # device tensors for the folds this rank owns

    * Gram totals: per-slab partial, one NCCL all-reduce (K x ld elements).
    * Fold Grams: per-slab partials in symmetric memory, summed by the fold owner over NVLink inside the epilogue kernel
      (no collective); NCCL all-reduce fallback when symmetric memory is unavailable or the model is float32.
    * numpy-order column sums (sequential over rows): chained rank to rank - a rank continues the running sums it
      receives and hands them on (2 x ld values per fold per hop), the last rank broadcasts the result.
    * weight sums (pairwise trees over all rows): every rank holds the whole weight vector and evaluates them itself.
    """

    def __init__(self, cvm: CVMatrix, N: int, K: int, M: int, weights=None, group=None, block_rows: int = 65536, bounds=None):
        import torch
        import torch.distributed as dist

        self.cvm, self.group, self.dist, self.torch = cvm, group, dist, torch
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.dev = torch.device("cuda", cvm.device)
        self.f64 = np.dtype(cvm.dtype) == np.float64
        self.tdt = torch.float64 if self.f64 else torch.float32
        self.N, self.K, self.M = int(N), int(K), int(M or 0)
        # bounds: slab boundaries b_0 .. b_W agreed by all ranks (upload_balanced_bounds); default: equal slabs
        if bounds is not None:
            if len(bounds) != self.world + 1 or bounds[0] != 0 or bounds[-1] != self.N or any(a > b for a, b in zip(bounds, bounds[1:])):
                raise ValueError("bounds must be world + 1 ascending row numbers from 0 to N")
            self.row0, self.row1 = int(bounds[self.rank]), int(bounds[self.rank + 1])
        else:
            self.row0, self.row1 = sharding.slab_rows(self.rank, self.world, self.N)
        self.block_rows = int(block_rows)
        self.w_host = None
        if weights is not None and not hasattr(weights, "data_ptr"):
            self.w_host = np.ascontiguousarray(np.asarray(weights, dtype=cvm.dtype).reshape(-1))
            weights = torch.from_numpy(self.w_host).to(self.dev)
        self.w = weights                     # all N weights, on the device (None: unweighted)
        self.use_peers = self.world > 1 and self.world <= 8 and self.f64
        import os

        self.decouple = os.environ.get("CVMX_SLAB_DECOUPLE", "1") != "0"   # scan passes before the carry arrives (float64)
        self._symm = None
        self._step = 0
        self._gram = None
        self.P = 0
        self._warm_chain()

    def _warm_chain(self) -> None:
        """One tiny message down the rank chain: NCCL opens its point-to-point channels lazily (~0.2 s per new pair), which
        would otherwise be paid hop after hop inside the first fit."""
        if self.world == 1:
            return
        t, dist = self.torch, self.dist
        token = t.zeros(1, dtype=t.float32, device=self.dev)
        if self.rank > 0:
            dist.recv(token, src=self._grank(self.rank - 1), group=self.group)
        if self.rank < self.world - 1:
            dist.send(token, dst=self._grank(self.rank + 1), group=self.group)
        dist.broadcast(token, src=self._grank(self.world - 1), group=self.group)
        t.cuda.current_stream(self.dev).synchronize()

    def _grank(self, r: int) -> int:
        return self.dist.get_global_rank(self.group, r) if self.group is not None else r

    def fit(self, blocks) -> None:
        t, dist, cvm = self.torch, self.dist, self.cvm
        lib, h = cvm._lib, cvm._h
        import os
        import time

        timing = os.environ.get("CVMX_SLAB_TIMING", "0") != "0"     # diagnostic: per-phase wall times (adds synchronisations)
        marks = []

        def mark(name):
            if timing:
                t.cuda.synchronize(self.dev)
                marks.append((name, time.perf_counter()))

        mark("begin")
        n_local = self.row1 - self.row0
        cvm.fit_begin(n_local, self.K, self.M, weighted=self.w is not None, max_block_rows=self.block_rows)
        # the weight sums (one CTA over all N weights) run on a side stream while the rows upload
        _lib.check(lib.cvmx_slab_begin(h, None if self.w is None else C.c_void_p(self.w.data_ptr()), self.N, self.row0), h)
        for b0, Xb, Yb in blocks:
            nb = int(Xb.shape[0])
            if b0 < self.row0 or b0 + nb > self.row1:
                raise ValueError("block outside this rank's row slab")
            wb = None
            if self.w is not None:
                if hasattr(Xb, "is_cuda") and Xb.is_cuda:
                    wb = self.w[b0:b0 + nb]
                else:                                    # host blocks take host weights
                    if self.w_host is None:
                        self.w_host = self.w.cpu().numpy()
                    wb = self.w_host[b0:b0 + nb]
            cvm.fit_rows(b0 - self.row0, Xb, Yb if self.M else None, wb, gram=True)
        ld = int(lib.cvmx_ld(h))
        vp = lambda x: None if x is None else C.c_void_p(x.data_ptr())  # noqa: E731
        mark("rows")
        # decoupled chain: every rank runs the streaming scan passes of its slab now; only the last pass waits for the carry
        self._decoupled(-1, 0, 1, ld, w_args=(vp(self.w), self.N, self.row0))
        mark("scan")
        carry = t.zeros((2, ld), dtype=self.tdt, device=self.dev)
        if self.world > 1 and self.rank > 0:
            dist.recv(carry, src=self._grank(self.rank - 1), group=self.group)
            t.cuda.current_stream(self.dev).synchronize()
        mark("recv")
        first = self.rank == 0
        _lib.check(lib.cvmx_fit_end_slab(h, None if first else vp(carry[0]), None if first else vp(carry[1]), vp(self.w), self.N, self.row0), h)
        sp, qp, mc = C.c_void_p(), C.c_void_p(), C.c_int64()
        _lib.check(lib.cvmx_moments_ptr(h, C.byref(sp), C.byref(qp), C.byref(mc)), h)
        ts = "<f8" if self.f64 else "<f4"
        sum_z = t.as_tensor(_DevArray(sp.value, ld, ts), device=self.dev)
        sumsq_z = t.as_tensor(_DevArray(qp.value, ld, ts), device=self.dev)
        mark("fit_end")
        if self.world > 1:
            carry[0].copy_(sum_z)
            carry[1].copy_(sumsq_z)
            if self.rank < self.world - 1:
                dist.send(carry, dst=self._grank(self.rank + 1), group=self.group)
            dist.broadcast(carry, src=self._grank(self.world - 1), group=self.group)   # the last slab's chains are the totals
            sum_z.copy_(carry[0])
            sumsq_z.copy_(carry[1])
            tp, cnt, ldt = C.c_void_p(), C.c_int64(), C.c_int64()
            _lib.check(lib.cvmx_totals_ptr(h, C.byref(tp), C.byref(cnt), C.byref(ldt)), h)
            dist.all_reduce(t.as_tensor(_DevArray(tp.value, cnt.value, ts), device=self.dev), group=self.group)
            t.cuda.current_stream(self.dev).synchronize()
            _lib.check(lib.cvmx_commit_totals(h), h)
        mark("collectives")
        cvm._streamed = True
        cvm.N = self.N
        cvm._pull_totals()
        mark("pull")
        if timing and marks:
            import sys

            print(f"[slab fit rank {self.rank}] " + " ".join(f"{b[0]}={1e3 * (b[1] - a[1]):.2f}ms" for a, b in zip(marks, marks[1:])),
                  file=sys.stderr, flush=True)

    def _decoupled(self, f0: int, f1: int, n_sets: int, ld: int, w_args=(None, 0, 0)) -> bool:
        """Streaming passes of the binade scan for the fit totals (f0 < 0) or the folds [f0, f1) of this slab, before the
        previous slab's chains are known (include/cvmx.h, "Decoupled slab chain"): local totals -> all-gather -> sums of the
        earlier slabs = approximate start -> passes 2 and 3.  Returns False (nothing done, on every rank) when some rank
        cannot take the path (float32, a slab too short); the chained call that follows then does all the work itself."""
        if self.world == 1 or not self.f64 or not self.decouple:
            return False
        t, dist, cvm = self.torch, self.dist, self.cvm
        lib, h = cvm._lib, cvm._h
        n = n_sets * 4 * ld
        mine = t.zeros((n + 1,), dtype=t.float64, device=self.dev)
        app = C.c_int32(0)
        _lib.check(lib.cvmx_slab_scan_local(h, f0, f1, w_args[0], w_args[1], w_args[2], C.c_void_p(mine.data_ptr()), C.byref(app)), h)
        mine[n] = float(app.value)
        allv = t.empty((self.world, n + 1), dtype=t.float64, device=self.dev)
        dist.all_gather_into_tensor(allv, mine, group=self.group)
        if float(allv[:, n].min().item()) < 1.0:          # one host sync per call; every rank sees the same flags
            return False                                   # the planes of a local-only attempt are never used
        start = allv[:self.rank, :n].sum(dim=0) if self.rank > 0 else t.zeros((n,), dtype=t.float64, device=self.dev)
        _lib.check(lib.cvmx_slab_scan_prepare(h, f0, f1, C.c_void_p(start.data_ptr())), h)
        self._start_keepalive = start     # the kernels read it asynchronously
        return True

    def set_folds(self, folds) -> None:
        from .partitioner import Partitioner

        cvm = self.cvm
        is_part = isinstance(folds, Partitioner)
        if is_part:
            offsets, indices = folds.csr()
        else:
            offsets, indices = folds
        offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        indices = np.ascontiguousarray(indices, dtype=np.int64)
        loc_off, loc_idx = sharding.local_csr(offsets, indices, self.row0, self.row1, assume_sorted=is_part)
        cvm._upload_csr(loc_off, loc_idx)
        _lib.check(cvm._lib.cvmx_set_weight_folds(cvm._h, offsets.ctypes.data_as(C.c_void_p), indices.ctypes.data_as(C.c_void_p),
                                                  offsets.size - 1), cvm._h)
        self.P = offsets.size - 1

    def alloc_outputs(self, n_folds: int):
        t, K, M = self.torch, self.K, self.M
        return dict(
            XTX=t.empty((n_folds, K, K), dtype=self.tdt, device=self.dev),
            XTY=t.empty((n_folds, K, M), dtype=self.tdt, device=self.dev) if M else None,
            stats=t.empty((n_folds, 2, K + M), dtype=self.tdt, device=self.dev),
            scal=t.empty((n_folds, 2), dtype=self.tdt, device=self.dev),
            status=t.empty((n_folds,), dtype=t.int32, device=self.dev),
        )

    def _setup_symm(self, half_elems: int):
        return ShardedFolds._setup_symm(self, half_elems)

    def training_batch(self, f0: int = 0, f1: Optional[int] = None, out: Optional[dict] = None):
        """Folds ``[f0, f1)``; returns device tensors for the folds this rank owns (``sharding.fold_block``).  Call on a
        non-default torch stream that the handle has been bound to (cvmx_set_stream)."""
        t, dist, cvm = self.torch, self.dist, self.cvm
        lib, h = cvm._lib, cvm._h
        if f1 is None:
            f1 = self.P
        o0, o1 = sharding.fold_block(self.rank, self.world, f0, f1)
        if out is None:
            out = self.alloc_outputs(max(o1 - o0, 1))
        vp = lambda x: None if x is None else C.c_void_p(x.data_ptr())  # noqa: E731
        want = _lib.WANT_XTX | (_lib.WANT_XTY if self.M else 0)
        ld = int(lib.cvmx_ld(h))
        # 1. the folds' column sums, chained slab to slab in row order (streaming scan passes first, on every rank at once)
        self._decoupled(f0, f1, f1 - f0, ld)
        carry = t.zeros((f1 - f0, 2, ld), dtype=self.tdt, device=self.dev)
        if self.world > 1 and self.rank > 0:
            dist.recv(carry, src=self._grank(self.rank - 1), group=self.group)
        _lib.check(lib.cvmx_slab_fold_sums(h, f0, f1, vp(carry)), h)
        if self.world > 1:
            if self.rank < self.world - 1:
                dist.send(carry, dst=self._grank(self.rank + 1), group=self.group)
            dist.broadcast(carry, src=self._grank(self.world - 1), group=self.group)
        # 2. weight masses (global), means and stds of every fold of the batch, on every rank
        _lib.check(lib.cvmx_slab_finalize_stats(h, f0, f1, vp(carry)), h)
        # 3. this slab's raw Gram of every fold, then the owners sum their peers' over NVLink inside the epilogue kernel
        n = lib.cvmx_sharded_gram_count(h, f0, f1, want)
        if self.use_peers and (self._symm is None or self._symm[2] < n):
            self._symm = self._setup_symm(n)
        if self._symm is not None and self.world > 1:
            buf, hdl, cap, ptrs = self._symm
            half = self._step & 1
            self._step += 1
            gram = buf[half * cap: half * cap + n]
            _lib.check(lib.cvmx_sharded_gram(h, f0, f1, want, 0, 1, vp(gram)), h)
            hdl.barrier(channel=0)
            _lib.check(lib.cvmx_sharded_finish_peers(h, f0, f1, o0, o1, want, ptrs[half], self.world, -1, vp(out["XTX"]), vp(out["XTY"]),
                                                     vp(out["stats"]), vp(out["scal"]), vp(out["status"])), h)
            return dict(out, fold_begin=o0, fold_end=o1)
        if self._gram is None or self._gram.numel() < n:
            self._gram = t.empty((n,), dtype=t.float64, device=self.dev)
        gram = self._gram[:n]
        _lib.check(lib.cvmx_sharded_gram(h, f0, f1, want, 0, 1, vp(gram)), h)
        if self.world > 1:
            dist.all_reduce(gram, group=self.group)
        _lib.check(lib.cvmx_sharded_finish(h, f0, o0, o1, want, vp(gram), vp(out["XTX"]), vp(out["XTY"]), vp(out["stats"]),
                                           vp(out["scal"]), vp(out["status"])), h)
        return dict(out, fold_begin=o0, fold_end=o1)
