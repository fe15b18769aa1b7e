"""One fold batch between cudaProfilerStart / cudaProfilerStop, for `ncu --profile-from-start off` captures of the fold
path alone (the launches of fit are not profiled):

    ncu --set full --import-source on --clock-control none --profile-from-start off -k regex:k_gram -c 1 \
        -o gpurun_out/x python tools/prof_once.py lmo

cases:  lmo   leave-many-out, cfg 3 at reduced N (300 folds x 1000 rows, K = 500, M = 10, float64): k_gram, fused epilogue
        kfold cfg 2 at reduced N (5 folds x 60k rows): k_gram with row-split partials + k_gram_reduce
        loo   leave-one-out (4000 folds, K = 500, M = 10): k_loo_operands + k_loo_tiles
        f32   wide float32 model (K = 2048, M = 32, 4 folds x 8k rows): k_gram_tc (tcgen05)
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from cvmatrix_b200 import CVMatrix, Partitioner  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "lmo"
rng = np.random.default_rng(42)
dtype = np.float64
if which == "lmo":
    N, K, M, P = 300_000, 500, 10, 300
elif which == "kfold":
    N, K, M, P = 300_000, 500, 10, 5
elif which == "loo":
    N, K, M, P = 4000, 500, 10, 4000
elif which == "f32":
    N, K, M, P, dtype = 32_768, 2048, 32, 4, np.float32
else:
    raise SystemExit("unknown case " + which)
X, Y, w = rng.random((N, K)).astype(dtype), rng.random((N, M)).astype(dtype), (rng.random(N) + 0.1).astype(dtype)
m = CVMatrix(dtype=dtype, copy=False)
m.fit(X, Y, w)
m.set_folds(Partitioner(np.arange(N) % P))
for _ in range(2):
    m.training_batch(out="torch")
torch.cuda.synchronize()
torch.cuda.profiler.start()
m.training_batch(out="torch")
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("PROF_ONCE_OK", which, m.launch_count)
