"""Worker of tests/test_gpu_multi.py: row-sharded fit + fold-sharded / row-sharded fold batches on WORLD_SIZE GPUs."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import ctypes as C  # noqa: E402

from cvmatrix_b200 import CVMatrix, Partitioner, _lib  # noqa: E402
from cvmatrix_b200 import sharding  # noqa: E402
from cvmatrix_b200.distributed import RowSlabFolds, ShardedFolds, fit_row_sharded, fit_sharded_upload  # noqa: E402
from cvmatrix_oracle import OracleCVMatrix, make_inputs, rel_fro  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
X, Y, w, folds = make_inputs(60_000, 200, 6, 3, seed=13)
orc = OracleCVMatrix()
orc.fit(X, Y, w)
m = CVMatrix(device=local)
fit_row_sharded(m, X, Y, w)
assert rel_fro(m.XTX, orc.XTX) <= 1e-14 and rel_fro(m.XTY, orc.XTY) <= 1e-14
assert np.array_equal(m.sum_X, orc.sum_X) and m.sum_w == orc.sum_w
part = Partitioner(folds)
m.set_folds(part)
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
_lib.check(m._lib.cvmx_set_stream(m._h, C.c_void_p(stream.cuda_stream)), m._h)
modes = []
for use_peers in (True, False):       # peer-memory reduction over NVLink (symmetric memory) and the NCCL all-reduce path
    sf = ShardedFolds(m)
    sf.use_peers = sf.use_peers and use_peers
    for row_sharded in (True, False):
        for rep in range(3):          # several steps: the symmetric buffer alternates between its two halves
            out = sf.training_batch(0, 3, row_sharded=row_sharded)
        torch.cuda.synchronize()
        for pos, f in enumerate(range(out["fold_begin"], out["fold_end"])):
            r = orc.fold(part.get_validation_indices(f))
            K = 200
            assert rel_fro(out["XTX"][pos].cpu().numpy(), r.XTX) <= 1e-12, (use_peers, row_sharded, f)
            assert rel_fro(out["XTY"][pos].cpu().numpy(), r.XTY) <= 1e-12, (use_peers, row_sharded, f)
            assert np.array_equal(out["stats"][pos, 0, :K].cpu().numpy(), r.X_mean[0]), (use_peers, row_sharded, f)
            assert np.array_equal(out["stats"][pos, 1, K:].cpu().numpy(), r.Y_std[0]), (use_peers, row_sharded, f)
    modes.append(sf._symm is not None)
if rank == 0:
    print("PEER_REDUCE_USED", modes)
owned = torch.zeros(3, device="cuda")
owned[out["fold_begin"]:out["fold_end"]] += 1
dist.all_reduce(owned)
assert bool((owned == 1).all())
# sharded upload: every rank copies 1 / world of the rows over PCIe, slabs exchanged over NVLink, column-sharded sums
m2 = CVMatrix(device=local)
fit_sharded_upload(m2, X, Y, w, block_rows=7000)
assert rel_fro(m2.XTX, orc.XTX) <= 1e-13 and rel_fro(m2.XTY, orc.XTY) <= 1e-13
assert np.array_equal(m2.sum_X, orc.sum_X) and np.array_equal(m2.sum_sq_X, orc.sum_sq_X)
assert np.array_equal(m2.sum_Y, orc.sum_Y) and np.array_equal(m2.sum_sq_Y, orc.sum_sq_Y)
assert m2.sum_w == orc.sum_w and m2.num_nonzero_w == orc.nnz_w
val = part.get_validation_indices(1)
(XTX, XTY), stats = m2.training_XTX_XTY(val)
r = orc.fold(val)
assert rel_fro(XTX, r.XTX) <= 1e-12 and rel_fro(XTY, r.XTY) <= 1e-12
for s, g in zip(stats, (r.X_mean, r.X_std, r.Y_mean, r.Y_std)):
    assert np.array_equal(s, g)
# X-only model (no Y): the sharded paths derive `want` from the fitted model (training_XTX is supported without Y)
mx = CVMatrix(device=local)
mx.fit(X, None, w)
ox = OracleCVMatrix()
ox.fit(X, None, w)
mx.set_folds(part)
_lib.check(mx._lib.cvmx_set_stream(mx._h, C.c_void_p(stream.cuda_stream)), mx._h)
for row_sharded in (True, False):
    sfx = ShardedFolds(mx)
    out = sfx.training_batch(0, 3, row_sharded=row_sharded)
    torch.cuda.synchronize()
    assert out["XTY"] is None
    for pos, f in enumerate(range(out["fold_begin"], out["fold_end"])):
        r = ox.fold(part.get_validation_indices(f), want_XTY=False)
        assert rel_fro(out["XTX"][pos].cpu().numpy(), r.XTX) <= 1e-12, ("no Y", row_sharded, f)
        assert np.array_equal(out["stats"][pos, 1, :200].cpu().numpy(), r.X_std[0]), ("no Y", row_sharded, f)
_lib.check(mx._lib.cvmx_set_stream(mx._h, None), mx._h)

# row-slab mode (BASELINE config 5): the ROWS are sharded across ranks; chained column sums, global weight sums,
# per-slab Grams summed by the fold owners over NVLink.  Uneven folds, zero weights, float64 and float32.
N = X.shape[0]
labels = np.random.default_rng(3).choice([0, 1, 2, 3, 4], size=N, p=[0.4, 0.3, 0.15, 0.1, 0.05])
part5 = Partitioner(labels)
w0 = w.copy()
w0[::9] = 0.0
for dt in (np.float64, np.float32):
    Xd, Yd, wd = X.astype(dt), Y.astype(dt), w0.astype(dt)
    od = OracleCVMatrix(dtype=dt)
    od.fit(Xd, Yd, wd)
    ms = CVMatrix(dtype=dt, device=local)
    rs = RowSlabFolds(ms, N, 200, 6, wd, block_rows=9000)
    r0, r1 = sharding.slab_rows(rank, world, N)
    rs.fit((b0, Xd[b0:min(r1, b0 + 9000)], Yd[b0:min(r1, b0 + 9000)]) for b0 in range(r0, r1, 9000))
    tol_t = 1e-13 if dt == np.float64 else 1e-6
    assert rel_fro(ms.XTX, od.XTX) <= tol_t and rel_fro(ms.XTY, od.XTY) <= tol_t, (dt, rel_fro(ms.XTX, od.XTX))
    assert np.array_equal(ms.sum_X, od.sum_X) and np.array_equal(ms.sum_sq_X, od.sum_sq_X), dt
    assert np.array_equal(ms.sum_Y, od.sum_Y) and np.array_equal(ms.sum_sq_Y, od.sum_sq_Y), dt
    assert ms.sum_w == od.sum_w and ms.num_nonzero_w == od.nnz_w, dt
    rs.set_folds(part5)
    _lib.check(ms._lib.cvmx_set_stream(ms._h, C.c_void_p(stream.cuda_stream)), ms._h)
    for rep in range(2):
        out = rs.training_batch()
    torch.cuda.synchronize()
    keys = list(part5.folds_dict)
    for pos, f in enumerate(range(out["fold_begin"], out["fold_end"])):
        r = od.fold(part5.get_validation_indices(keys[f]))
        for name, row, sl, g in (("X_mean", 0, slice(0, 200), r.X_mean), ("X_std", 1, slice(0, 200), r.X_std),
                                 ("Y_mean", 0, slice(200, 206), r.Y_mean), ("Y_std", 1, slice(200, 206), r.Y_std)):
            assert np.array_equal(out["stats"][pos, row, sl].cpu().numpy(), g[0]), ("slab", dt, f, name)
        if dt == np.float64:
            assert rel_fro(out["XTX"][pos].cpu().numpy(), r.XTX) <= 1e-12, ("slab", f, rel_fro(out["XTX"][pos].cpu().numpy(), r.XTX))
            assert rel_fro(out["XTY"][pos].cpu().numpy(), r.XTY) <= 1e-12, ("slab", f, rel_fro(out["XTY"][pos].cpu().numpy(), r.XTY))
    if rank == 0:
        print("SLAB_OK", np.dtype(dt).name, "peers" if rs._symm is not None else "nccl")
    _lib.check(ms._lib.cvmx_set_stream(ms._h, None), ms._h)
dist.barrier()
if rank == 0:
    print("DIST_OK", world)
dist.destroy_process_group()
