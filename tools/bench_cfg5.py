"""
BASELINE.json config 5 at FULL size on one B200: N = 2,000,000, K = 5,000, M = 100, 10 folds, weighted,
center + scale, float64 (82 GB resident) or float32 (41 GB).  The matrix cannot exist on the host, so it is
produced block by block ON THE DEVICE (torch.rand, seeded per block) and fed through the streaming fit
(cvmx_fit_begin / cvmx_fit_rows / cvmx_fit_end); the fold path is the ordinary batched call.

    python tools/bench_cfg5.py [--dtype f64|f32] [--rows 2000000] [--steps 2]

Prints one JSON line: fit seconds, fold-matrices/s of the fold path, DMMA-pipe roofline fraction, and a parity
check of fold 0's statistics / matrices against a float64 torch recomputation on a 3-column-block sample.
"""
import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
from cvmatrix_b200 import CVMatrix, Partitioner, _lib  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--dtype", default="f64", choices=["f64", "f32"])
ap.add_argument("--rows", type=int, default=2_000_000)
ap.add_argument("--cols", type=int, default=5000)
ap.add_argument("--resp", type=int, default=100)
ap.add_argument("--folds", type=int, default=10)
ap.add_argument("--steps", type=int, default=2)
ap.add_argument("--block", type=int, default=65536)
args = ap.parse_args()
N, K, M, P, B = args.rows, args.cols, args.resp, args.folds, args.block
npdt, tdt = (np.float64, torch.float64) if args.dtype == "f64" else (np.float32, torch.float32)
dev = torch.device("cuda", 0)


def block(b0, b1):
    g = torch.Generator(device=dev)
    g.manual_seed(42_000 + b0 // B)
    blk = torch.rand((b1 - b0, K + M + 1), dtype=torch.float64, generator=g, device=dev).to(tdt)
    return blk[:, :K], blk[:, K:K + M], blk[:, K + M].contiguous()


m = CVMatrix(dtype=npdt, device=0)
lib, h = m._lib, m._h
torch.cuda.synchronize()
t0 = time.perf_counter()
m.fit_begin(N, K, M, weighted=True, max_block_rows=B)
t_gen = 0.0
for b0 in range(0, N, B):
    b1 = min(N, b0 + B)
    tg = time.perf_counter()
    Xb, Yb, wb = block(b0, b1)
    torch.cuda.synchronize()
    t_gen += time.perf_counter() - tg
    m.fit_rows(b0, Xb, Yb, wb)
del Xb, Yb, wb
m.fit_end()
torch.cuda.synchronize()
fit_s = time.perf_counter() - t0 - t_gen
torch.cuda.empty_cache()

folds = np.arange(N) % P
part = Partitioner(folds)
m.set_folds(part)
stream = torch.cuda.Stream(dev)
torch.cuda.set_stream(stream)
_lib.check(lib.cvmx_set_stream(h, C.c_void_p(stream.cuda_stream)), h)
oxx = torch.empty((P, K, K), dtype=tdt, device=dev)
oxy = torch.empty((P, K, M), dtype=tdt, device=dev)
ost = torch.empty((P, 2, K + M), dtype=tdt, device=dev)
vp = lambda t: C.c_void_p(t.data_ptr())  # noqa: E731


def step():
    _lib.check(lib.cvmx_training_batch(h, 0, P, 3, vp(oxx), vp(oxy), vp(ost), None, None, _lib.DEVICE), h)


step()
torch.cuda.synchronize()
_lib.check(lib.cvmx_profile_enable(h, 1), h)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(stream)
for _ in range(args.steps):
    step()
e1.record(stream)
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / args.steps
pm, pn = (C.c_double * 3)(), (C.c_int64 * 3)()
_lib.check(lib.cvmx_profile_read(h, pm, pn), h)
_lib.check(lib.cvmx_profile_enable(h, 0), h)
gram_ms = pm[1] / args.steps
flops = 2.0 * N * K * (K + M)
with open(os.path.join(ROOT, "profiles", "r01_fp64_calibration.json")) as f:
    peak = float(json.load(f)["dmma_tflops_bps2_w16_acc16"])

# ---- parity on a sample: fold 0, statistics of the first 64 columns and the leading 64 x 64 block of XTX -------
S = 64
idx = torch.from_numpy(part.get_validation_indices(0)).to(dev)
zp, wp, ld = C.c_void_p(), C.c_void_p(), C.c_int64()
_lib.check(lib.cvmx_data_ptr(h, C.byref(zp), C.byref(wp), C.byref(ld)), h)


class _Dev:
    def __init__(self, ptr, count, ts):
        self.__cuda_array_interface__ = {"shape": (count,), "typestr": ts, "data": (ptr, False), "version": 3}


ts = "<f8" if args.dtype == "f64" else "<f4"
Z = torch.as_tensor(_Dev(zp.value, N * ld.value, ts), device=dev).view(N, ld.value)
wv = torch.as_tensor(_Dev(wp.value, N, ts), device=dev)
Xs = Z[:, :S].double()
wd = wv.double()
tr = torch.ones(N, dtype=torch.bool, device=dev)
tr[idx] = False
Xt, wt = Xs[tr], wd[tr]
sw = wt.sum()
mean = (Xt * wt[:, None]).sum(0) / sw
nnz = (wt != 0).sum()
var = ((Xt - mean) ** 2 * wt[:, None]).sum(0) / ((nnz - 1) * sw / nnz)
std = var.sqrt()
Xc = (Xt - mean) / std
ref = (Xc * wt[:, None]).T @ Xc
got = oxx[0, :S, :S].double()
err_xx = float((got - ref).norm() / ref.norm())
err_mean = float((ost[0, 0, :S].double() - mean).abs().max() / mean.abs().max())
err_std = float((ost[0, 1, :S].double() - std).abs().max() / std.abs().max())

line = {
    "workload": f"cfg5 full size: N={N} K={K} M={M} {args.dtype} weighted center+scale {P}-fold, data generated on the device",
    "resident_gb": N * ld.value * (8 if args.dtype == "f64" else 4) / 1e9,
    "fit_s_excluding_generation": fit_s, "generation_s": t_gen,
    "fit_tflops_full_count": flops / fit_s / 1e12,
    "fold_path_ms_per_step": ms, "value_fold_matrices_per_s": P / (ms * 1e-3),
    "k_gram_ms_per_step": gram_ms, "stats_ms_per_step": pm[0] / args.steps,
    "roofline": {"bound": "tensor", "achieved": flops / (gram_ms * 1e-3) / 1e12, "peak": peak, "unit": "TFLOP/s",
                 "frac": flops / (gram_ms * 1e-3) / 1e12 / peak, "note": "full flop count 2 N K (K+M), no symmetry credit"},
    "parity_sample": {"relfro_XTX_64x64_vs_torch_f64_recomputation": err_xx, "rel_max_mean": err_mean, "rel_max_std": err_std},
    "gpu_launches": m.launch_count, "scan_launches": m.scan_launch_count,
}
print(json.dumps(line))
