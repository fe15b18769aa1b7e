#!/usr/bin/env python
"""
bench.py - fold matrices / second of the cvmatrix hot path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config cfg2|cfg3|cfg4] [--impl native|reference]

One "step" = one pass of the batched fold path (cvmx_training_batch: weight masses, numpy-order
moments, DMMA Gram downdate, fused centering/scaling epilogue) over ALL folds of the workload, inputs
resident in HBM, outputs written to HBM.  `value` = folds processed by all ranks / max-over-ranks device
time.  `e2e` = the reference benchmark's own definition (benchmarks/benchmark.py:52-158): Partitioner +
fit (host->device copy of X, Y, w from pinned memory) + all folds + device->host copy of every result,
through the public CVMatrix API.  `cpu_baseline` / `--impl reference` time the numpy restatement of the
reference (oracle/cvmatrix_oracle.py, order="numpy": the same numpy calls the reference makes) on the
box's host cores.

Workloads (BASELINE.json configs; inputs per benchmarks/benchmark.py:223-232, seed 42, uniform [0,1)):
  cfg2  N=1,000,000 K=500 M=10 float64 weighted center+scale, 5 folds      (default; metric config)
  cfg3  same data, 1,000 folds
  cfg4  leave-one-out N=20,000 K=500 M=10 (20,000 folds)
Inputs (4.09 GB) are far larger than the 126 MB L2, so consecutive steps cannot reuse cached inputs.
"""

from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CONFIGS = {
    "cfg2": dict(N=1_000_000, K=500, M=10, P=5, name="N=1M K=500 M=10 f64 weighted center+scale 5-fold"),
    "cfg3": dict(N=1_000_000, K=500, M=10, P=1000, name="N=1M K=500 M=10 f64 weighted center+scale 1000-fold"),
    "cfg4": dict(N=20_000, K=500, M=10, P=20_000, name="LOO N=20k K=500 M=10 f64 weighted center+scale"),
    # contract self-test shape (tests/test_bench_contract.py)
    "tiny": dict(N=4_000, K=24, M=3, P=4, name="self-test: N=4k K=24 M=3 f64 4-fold"),
    # cfg 5 (wide, K=5000 M=100) at reduced N: the full N=2M matrix is 80 GB and cannot be generated on the host
    "cfg5s": dict(N=100_000, K=5000, M=100, P=10, name="wide K=5000 M=100 f64 weighted center+scale 10-fold at N=100k (cfg 5 scaled to 1/20 of its rows)"),
    # reduced shapes for ncu captures only (same per-CTA work as cfg2 / cfg3 / cfg4, fewer CTAs)
    "prof2": dict(N=200_000, K=500, M=10, P=5, name="profiling: N=200k K=500 M=10 f64 5-fold"),
    "prof3": dict(N=200_000, K=500, M=10, P=200, name="profiling: N=200k K=500 M=10 f64 200-fold"),
    "prof4": dict(N=2_000, K=500, M=10, P=2_000, name="profiling: LOO N=2k K=500 M=10 f64"),
}
METRIC = "fold matrices/sec"
UNIT = "fold-matrices/s"


def fp64_peak_tflops():
    """Measured DMMA.8x8x4 issue-rate peak on this pool's B200 (tools/ubench_fp64.cu ->
    profiles/r01_fp64_calibration.json); MEASURED_PEAKS.json carries no FP64 figure."""
    try:
        with open(os.path.join(ROOT, "profiles", "r01_fp64_calibration.json")) as f:
            d = json.load(f)
        return float(d["dmma_tflops_bps2_w16_acc16"]), "measured DMMA.8x8x4 peak (profiles/r01_fp64_calibration.json)"
    except Exception:
        return 37.0, "fallback: nominal B200 FP64 37 TFLOP/s"


def hbm_peak_gbs():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def make_host_inputs(cfg, pinned):
    """Seeded synthetic inputs written straight into (optionally pinned) host buffers."""
    N, K, M, P = cfg["N"], cfg["K"], cfg["M"], cfg["P"]
    rng = np.random.default_rng(42)
    if pinned:
        import torch

        Xt = torch.empty((N, K), dtype=torch.float64, pin_memory=True)
        Yt = torch.empty((N, M), dtype=torch.float64, pin_memory=True)
        wt = torch.empty((N,), dtype=torch.float64, pin_memory=True)
        X, Y, w = Xt.numpy(), Yt.numpy(), wt.numpy()
        keep = (Xt, Yt, wt)
    else:
        X, Y, w = np.empty((N, K)), np.empty((N, M)), np.empty(N)
        keep = None
    rng.random(out=X)
    rng.random(out=Y)
    rng.random(out=w)
    folds = np.arange(N) % P
    return X, Y, w, folds, keep


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 7 for n, v in zip(names, r[3:7]) if v.lower().startswith("active")})
        busy = [s for s in sm if s > 0.5 * (max(mx) if mx else 1)] or sm
        return {"sm_mhz": float(np.median(busy)) if busy else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def run_reference(args, cfg, rank, world):
    """Reference arm: the reference's CPU path (numpy restatement, all host threads) on a bounded row sample of
    the same workload; one step = Partitioner + fit + every fold, as benchmarks/benchmark.py times it."""
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    from cvmatrix_oracle import OracleCVMatrix, OraclePartitioner

    N, K, M, P = cfg["N"], cfg["K"], cfg["M"], cfg["P"]
    X, Y, w, folds, _ = make_host_inputs(cfg, pinned=False)

    def one_step(n_rows, n_folds):
        t0 = time.perf_counter()
        part = OraclePartitioner(folds[:n_rows])
        m = OracleCVMatrix(dtype=np.float64, copy=False, order="numpy")
        m.fit(X[:n_rows], Y[:n_rows], w[:n_rows])
        t1 = time.perf_counter()
        for f in list(part.folds_dict)[:n_folds]:
            m.training_XTX_XTY(part.get_validation_indices(f))
        t2 = time.perf_counter()
        return t1 - t0, t2 - t1

    # calibrate on a small slice, then size the sample so the whole run stays within ~3 minutes
    n_cal = min(N, 50_000 if P <= 1000 else N)
    f_cal = min(P, 50)
    fit_s, fold_s = one_step(n_cal, f_cal)
    est_full = fit_s * N / n_cal + (fold_s / f_cal) * P * (N / n_cal if P <= 1000 else 1.0)
    budget = 170.0 / max(1, args.steps + args.warmup)
    if P <= 1000:  # large folds: sample rows (cost is linear in N), keep every fold
        frac = min(1.0, budget / est_full)
        n_rows = max(P * 20, int(N * frac) // P * P)
        n_folds, scale = P, N / n_rows
    else:          # leave-one-out: full fit, sample folds (cost is linear in the number of folds)
        n_rows = N
        n_folds = int(max(50, min(P, (budget - fit_s) / max(fold_s / f_cal, 1e-9))))
        scale = None
    for _ in range(args.warmup):
        one_step(n_rows, n_folds)
    t_fit = t_fold = 0.0
    for _ in range(args.steps):
        a, b = one_step(n_rows, n_folds)
        t_fit += a
        t_fold += b
    t_fit /= args.steps
    t_fold /= args.steps
    if scale is not None:
        step_s = (t_fit + t_fold) * scale
        sample = f"first {n_rows} of {N} rows, all {P} folds, time scaled linearly by {scale:.2f}"
    else:
        step_s = t_fit + t_fold * (P / n_folds)
        sample = f"full fit, first {n_folds} of {P} folds, fold time scaled by {P / n_folds:.1f}"
    value = P / step_s
    cores = os.cpu_count()
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": step_s * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": cfg["name"], "step": "Partitioner + fit + all folds (training_XTX_XTY), host arrays, copy=False"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                         "fit_s_sample": t_fit, "folds_s_sample": t_fold},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--config", default="cfg2", choices=sorted(CONFIGS))
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--e2e-out", default="numpy", choices=["numpy", "pinned"],
                    help="host destination of the e2e results: fresh numpy arrays (the reference's convention) or the reused page-locked pool")
    args = ap.parse_args()
    cfg = CONFIGS[args.config]
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    args.warmup = max(args.warmup, 3) if args.impl == "native" else args.warmup

    if args.impl == "reference":
        run_reference(args, cfg, rank, world)
        return

    import torch
    import torch.distributed as dist

    from cvmatrix_b200 import CVMatrix, Partitioner, _lib

    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        # NCCL_DEBUG=VERSION / INFO print to stdout; the contract is ONE JSON line there
        if os.environ.get("NCCL_DEBUG", "").upper() in ("VERSION", "INFO"):
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()

    N, K, M, P = cfg["N"], cfg["K"], cfg["M"], cfg["P"]
    X, Y, w, folds, keep = make_host_inputs(cfg, pinned=True)
    part = Partitioner(folds)
    m = CVMatrix(dtype=np.float64, copy=False, device=local_rank)
    lib, h = m._lib, m._h
    t0 = time.perf_counter()
    m.fit(X, Y, w)
    fit_upload_s = time.perf_counter() - t0
    m.set_folds(part)
    # folds are sharded across ranks in contiguous blocks; no data-path collective
    f0, f1 = rank * P // world, (rank + 1) * P // world
    Pl = f1 - f0
    # LOO writes 2.04 MB per fold: keep the resident output window bounded (it is rewritten every chunk)
    chunk = min(max(Pl, 1), 4096)
    oxx = torch.empty((chunk, K, K), dtype=torch.float64, device=dev)
    oxy = torch.empty((chunk, K, M), dtype=torch.float64, device=dev)
    ost = torch.empty((chunk, 2, K + M), dtype=torch.float64, device=dev)
    osc = torch.empty((chunk, 2), dtype=torch.float64, device=dev)
    oss = torch.empty((chunk,), dtype=torch.int32, device=dev)
    stream = torch.cuda.Stream(dev)   # a real (non-default) stream shared by the library and the timing events
    torch.cuda.set_stream(stream)
    _lib.check(lib.cvmx_set_stream(h, C.c_void_p(stream.cuda_stream)), h)
    vp = lambda t: C.c_void_p(t.data_ptr())  # noqa: E731

    from cvmatrix_b200 import sharding
    from cvmatrix_b200.distributed import ShardedFolds

    row_sharded = sharding.use_row_sharding(P, world)
    sf = ShardedFolds(m) if world > 1 else None
    emulate = int(os.environ.get("BENCH_EMULATE_SHARDS", "0"))   # profiling aid: rank 0's share of an N-way sharded step
    if emulate and world == 1:
        sf = ShardedFolds(m)
        sf.emulate_shards = emulate
        row_sharded = True
    outs = dict(XTX=oxx, XTY=oxy, stats=ost, scal=osc, status=oss)

    def step():
        if row_sharded:
            # few large folds: rows of every fold split across ranks, 1 NCCL all-reduce, owners finish their folds
            sf.training_batch(0, P, out=outs, row_sharded=True)
            return
        for c0 in range(f0, f1, chunk):
            c1 = min(f1, c0 + chunk)
            _lib.check(lib.cvmx_training_batch(h, c0, c1, 3, vp(oxx), vp(oxy), vp(ost), vp(osc), vp(oss), _lib.DEVICE), h)

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    _lib.check(lib.cvmx_profile_enable(h, 1), h)
    launches0 = m.launch_count
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        step()
    e1.record(stream)
    torch.cuda.synchronize()
    barrier()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    launches = m.launch_count - launches0
    prof_ms = (C.c_double * 3)()
    prof_n = (C.c_int64 * 3)()
    _lib.check(lib.cvmx_profile_read(h, prof_ms, prof_n), h)
    _lib.check(lib.cvmx_profile_enable(h, 0), h)
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    ms_per_step = ms / args.steps
    value = P / (ms_per_step * 1e-3)

    # ---- roofline of the dominant kernel (k_gram: DMMA Gram + fused epilogue) on this rank ---------------
    n_val_total = int(part.offsets[f1] - part.offsets[f0]) if not row_sharded else int(part.offsets[P]) // world
    flops_per_step = 2.0 * n_val_total * K * (K + M)             # full (no symmetry credit), SURVEY.md 8(d)
    bytes_per_step = 2.0 * 8 * K * (K + M) * Pl                  # read total + write result per fold
    gram_ms = prof_ms[1] / max(1, args.steps)                    # all k_gram launches of one step
    gram_launches = prof_n[1] / max(1, args.steps)
    peak_tf, peak_src = fp64_peak_tflops()
    peak_bw, bw_src = hbm_peak_gbs()
    t_flop, t_byte = flops_per_step / (peak_tf * 1e12), bytes_per_step / (peak_bw * 1e9)
    if t_flop >= t_byte:
        ach = flops_per_step / (gram_ms * 1e-3) / 1e12 if gram_ms > 0 else None
        roof = {"bound": "tensor", "achieved": ach, "peak": peak_tf, "unit": "TFLOP/s", "frac": ach / peak_tf if ach else None,
                "peak_source": peak_src + "; FP64 tensor pipe"}
    else:
        ach = bytes_per_step / (gram_ms * 1e-3) / 1e9 if gram_ms > 0 else None
        roof = {"bound": "hbm", "achieved": ach, "peak": peak_bw, "unit": "GB/s", "frac": ach / peak_bw if ach else None,
                "peak_source": bw_src}
    if roof["bound"] == "tensor" and ach:
        # flops the kernel actually issues: tiles on / above the diagonal only, diagonal tiles at 3/4
        TI, TJ = -(-K // 128), -(-(K + M) // 128)
        issued_tiles = sum((0.75 if bj == bi else 1.0) for bi in range(TI) for bj in range(bi, TJ))
        issued = 2.0 * n_val_total * 128 * 128 * issued_tiles
        roof["issued_flops_per_step"] = issued
        roof["issued_frac_of_peak"] = issued / (gram_ms * 1e-3) / 1e12 / peak_tf
        roof["note"] = ("achieved / frac use the FULL flop count 2 N_val K (K+M) of SURVEY.md 8(d); XTX is symmetric, so only "
                        "upper-triangular tiles are computed and frac can exceed 1 - issued_frac_of_peak is the DMMA pipe's own load")
    roof.update({"kernel": "k_gram<double>", "traffic": None, "kernel_ms_per_step": gram_ms, "kernel_launches_per_step": gram_launches,
                 "stats_ms_per_step": prof_ms[0] / max(1, args.steps), "reduce_ms_per_step": prof_ms[2] / max(1, args.steps),
                 "algorithmic_flops_per_step": flops_per_step, "algorithmic_bytes_per_step": bytes_per_step,
                 "step_roofline_frac": max(t_flop, t_byte) / (ms_per_step * 1e-3)})
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            roof["traffic"] = json.load(f).get(args.config)
    except Exception:
        pass

    # ---- end to end through the public API, host buffers in, host arrays out ------------------------------
    e2e = None
    if not args.no_e2e:
        from cvmatrix_b200.distributed import fit_sharded_upload

        if world == 1:
            _lib.check(lib.cvmx_set_stream(h, None), h)
        e2e_steps = args.steps if cfg["P"] <= 1000 else max(1, min(args.steps, 2))
        out_bytes = 0
        host_out = None
        if row_sharded and world > 1:   # pinned host buffers for the folds this rank owns
            o0, o1 = sharding.fold_block(rank, world, 0, P)
            n_own = max(o1 - o0, 1)
            host_out = {k: torch.empty((n_own,) + tuple(outs[k].shape[1:]), dtype=outs[k].dtype, pin_memory=True)
                        for k in ("XTX", "XTY", "stats", "scal", "status")}

        parts = {"partitioner_ms": 0.0, "fit_ms": 0.0, "set_folds_ms": 0.0, "folds_ms": 0.0}

        def e2e_step():
            nonlocal out_bytes
            t0 = time.perf_counter()
            p2 = Partitioner(folds)
            t1 = time.perf_counter()
            if world > 1:
                # every rank uploads 1 / world of the rows over its own PCIe link; slabs are exchanged over NVLink
                fit_sharded_upload(m, X, Y, w)
                t2 = time.perf_counter()
                m.set_folds(p2)
            else:
                # fit + set_folds in one call: the folds partition the rows, so every row is contracted once, per fold,
                # behind the upload (XtWX = sum of the fold Grams) and training_batch only finishes the folds
                m.fit(X, Y, w, folds=p2)
                t2 = time.perf_counter()
            t3 = time.perf_counter()
            out_bytes = 0
            if row_sharded and world > 1:
                res = sf.training_batch(0, P, out=outs, row_sharded=True)
                n = res["fold_end"] - res["fold_begin"]
                for k, hbuf in host_out.items():
                    if n > 0:
                        hbuf[:n].copy_(outs[k][:n], non_blocking=True)
                        out_bytes += hbuf[:n].numel() * hbuf.element_size()
                torch.cuda.synchronize()
            else:
                for c0 in range(f0, f1, chunk):
                    r = m.training_batch(c0, min(f1, c0 + chunk), out=args.e2e_out)
                    out_bytes += r["XTX"].nbytes + r["XTY"].nbytes + 2 * r["X_mean"].nbytes + 2 * r["Y_mean"].nbytes
            t4 = time.perf_counter()
            for k, v in zip(parts, (t1 - t0, t2 - t1, t3 - t2, t4 - t3)):
                parts[k] += v * 1e3

        e2e_step()
        for k in parts:
            parts[k] = 0.0
        barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            e2e_step()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        h2d = int(X.nbytes + Y.nbytes + w.nbytes) // world + int(part.indices.nbytes + part.offsets.nbytes)
        if world > 1:
            t = torch.tensor([dt, float(out_bytes)], dtype=torch.float64, device=dev)
            dist.all_reduce(t[:1], op=dist.ReduceOp.MAX)
            dist.all_reduce(t[1:], op=dist.ReduceOp.SUM)
            dt, out_bytes = float(t[0].item()), int(t[1].item())
            h2d *= world
        e2e = {"value": P / (dt / e2e_steps), "unit": UNIT, "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": int(out_bytes), "ms_per_step": dt / e2e_steps * 1e3, "steps": e2e_steps,
               "includes": ("Partitioner + fit (H2D from pinned host memory" + (f": 1/{world} of the rows per rank, slabs exchanged over NVLink" if world > 1 else "; fused with the fold Grams when the folds partition the rows")
                            + ") + set_folds + all folds + D2H of every output"),
               "host_outputs": args.e2e_out, "breakdown_ms": {k: v / e2e_steps for k, v in parts.items()}}

    # ---- CPU baseline: numpy restatement of the reference on this box's host cores (rank 0, N = 1 only) ----
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        from cvmatrix_oracle import OracleCVMatrix, OraclePartitioner

        t0 = time.perf_counter()
        op = OraclePartitioner(folds)
        orc = OracleCVMatrix(dtype=np.float64, copy=False, order="numpy")
        orc.fit(X, Y, w)
        t_fit = time.perf_counter() - t0
        keys = list(op.folds_dict)
        n_f = len(keys) if P <= 5 else max(5, min(len(keys), int(12.0 / (9.0 / P if P <= 1000 else 0.005))))
        n_f = min(n_f, 2000)
        t0 = time.perf_counter()
        for k in keys[:n_f]:
            orc.training_XTX_XTY(op.get_validation_indices(k))
        t_folds = (time.perf_counter() - t0) * (len(keys) / n_f)
        cpu = {"value": P / t_folds, "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
               "sample": f"full-size fit ({t_fit:.1f} s, not in value) then {n_f} of {P} folds, fold-path time scaled to {P} folds",
               "fit_s": t_fit, "folds_s": t_folds, "e2e_value": P / (t_fit + t_folds)}
        del orc

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": cfg["name"], "N": N, "K": K, "M": M, "folds": P, "parallelism": ((f"rows of each fold sharded x{world} + " + ("peer-memory reduction over NVLink (symmetric memory, no all-reduce)" if (sf is not None and sf._symm is not None) else "1 NCCL all-reduce")) if row_sharded else f"fold-sharded x{world}"),
                       "l2_policy": "inputs (4.09 GB) larger than L2; no flush needed" if N * K * 8 > 2e8 else "inputs smaller than L2 (LOO): outputs (>=8 GB per step) stream through L2",
                       "step": "batched fold path over all folds, inputs resident in HBM, outputs to HBM"},
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roof, "cpu_baseline": cpu,
            "fit_with_h2d_s": fit_upload_s,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
