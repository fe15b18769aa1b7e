"""One column-sharded (8-way, then 1-way) statistics pass of cfg 2 with the binade scan forced on: the workload of the
ncu captures in tools/final_evidence.sh (profiles/r01_scan_ncu_raw.csv, r01_launches_scan.csv)."""
import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from cvmatrix_b200 import CVMatrix, Partitioner, _lib
N, K, M, P = 1_000_000, 500, 10, 5
rng = np.random.default_rng(42)
X, Y, w = rng.random((N, K)), rng.random((N, M)), rng.random(N)
m = CVMatrix(copy=False); m.fit(X, Y, w); m.set_folds(Partitioner(np.arange(N) % P))
lib, h = m._lib, m._h
m.set_scan_mode(2)
for shards in (8, 1):
    for _ in range(2):
        sp, sc = C.c_void_p(), C.c_int64()
        _lib.check(lib.cvmx_sharded_stats(h, 0, P, 0, shards, C.byref(sp), C.byref(sc)), h)
        _lib.check(lib.cvmx_sharded_stats_wait(h), h)
        m.sync()
