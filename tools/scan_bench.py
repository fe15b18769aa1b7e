"""Times the statistics of a fold batch (numpy-order column sums) with the dependent-add chains (scan mode 0) and the
binade scan (mode 2), bit-compares them, and reports the share a rank of an n-way column-sharded step would run.

    python tools/scan_bench.py [N K M P]      (default: cfg 2, N=1M K=500 M=10 P=5)
"""
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from cvmatrix_b200 import CVMatrix, Partitioner, _lib  # noqa: E402

N, K, M, P = (int(a) for a in sys.argv[1:5]) if len(sys.argv) >= 5 else (1_000_000, 500, 10, 5)
rng = np.random.default_rng(42)
X, Y, w = rng.random((N, K)), rng.random((N, M)), rng.random(N)
m = CVMatrix(copy=False)
m.fit(X, Y, w)
m.set_folds(Partitioner(np.arange(N) % P))
lib, h = m._lib, m._h
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
_lib.check(lib.cvmx_set_stream(h, C.c_void_p(stream.cuda_stream)), h)
ld = lib.cvmx_ld(h)
out = {}
res = {}


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(reps):
        fn()
    e1.record(stream)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for mode in (0, 2):
    m.set_scan_mode(mode)
    st = torch.zeros((P, 2, K + M), dtype=torch.float64, device="cuda")

    def full():
        _lib.check(lib.cvmx_training_batch(h, 0, P, 0, None, None, C.c_void_p(st.data_ptr()), None, None, _lib.DEVICE), h)

    out[f"stats_all_columns_ms_mode{mode}"] = timed(full)
    res[mode] = st.cpu().numpy().copy()
    for shards in (2, 4, 8):
        def shard():
            sp, sc = C.c_void_p(), C.c_int64()
            _lib.check(lib.cvmx_sharded_stats(h, 0, P, 0, shards, C.byref(sp), C.byref(sc)), h)
            _lib.check(lib.cvmx_sharded_stats_wait(h), h)

        out[f"stats_rank0_of_{shards}_ms_mode{mode}"] = timed(shard)
out["bit_identical"] = bool(np.array_equal(res[0], res[2]))
out["shape"] = dict(N=N, K=K, M=M, P=P)
print(json.dumps(out))
