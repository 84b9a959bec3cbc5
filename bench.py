#!/usr/bin/env python
"""bench.py -- headline benchmark of the sparse x dense hot path on B200.

Metric (BASELINE.json): SpMM nnz*RHS/s (+ HBM roofline fraction) for the binary-CSR
A_mul_Bn product with 32 right-hand sides on a synthetic 10M x 1M matrix with 200M
nonzeros ("C2"), on 1/2/4/8 GPUs, next to the reference's OpenMP CPU path.

  python bench.py --gpus N --steps K --warmup W            our arm (one rank per GPU under torchrun)
  python bench.py --impl reference --gpus N --steps K ...  the reference's own CPU code (rank 0 only)

A "step" is one full Y = A X pass over one rank's matrix.  N > 1: every rank holds its
own C2-sized row shard (rows are independent, X is replicated, Y stays row-sharded: no
data-path collective) => weak scaling; value = total nnz*R of all ranks / max-over-ranks
time.  Inputs are larger than L2 (cols 0.8 GB, X 256 MB, Y 2.56 GB), so no L2 flush is
needed between iterations.
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

WORKLOADS = {
    # name: (nrow, ncol, nnz, R, dist, seed)
    "c2": (10_000_000, 1_000_000, 200_000_000, 32, 0, 0x5EED0002),
    "c2_small": (1_000_000, 100_000, 20_000_000, 32, 0, 0x5EED0002),
}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def alg_bytes(nrow, nnz, R):
    """B_alg = nnz*(4 + 8R) + 4(N+1) + 8NR  (dense operand counted per gather: X > L2)."""
    return nnz * (4 + 8 * R) + 4 * (nrow + 1) + 8 * nrow * R


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except (ValueError, IndexError):
                continue
            names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
            for name, val in zip(names, r[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------- CPU arm
def cpu_sample_matrix(ncol, R, seed, sample_rows, per_row):
    """A bounded sample of the workload: `sample_rows` rows with the workload's column count, row
    degree distribution (uniform COO => Poisson(per_row)) and dense operand (full X)."""
    import libfastsparse_b200 as fs
    import oracle
    nnz = int(sample_rows * per_row)
    rows, cols, _ = fs.synth_coo_host(seed, 0, nnz, sample_rows, ncol)       # input generation only
    row_ptr, ccols, _ = oracle.csr_from_coo(sample_rows, rows, cols)
    X = np.ascontiguousarray(np.sin(7.0 * np.arange(ncol)[:, None] + 17.0 * np.arange(R)[None, :] + 0.3))
    return nnz, row_ptr, np.ascontiguousarray(ccols), X


def cpu_time_steps(ncol, R, seed, sample_rows, per_row, steps, warmup):
    """Times the reference's own bcsr_A_mul_B32n (csr.h:283-302, unmodified, OpenMP, all host threads)
    when oracle/_ref is present, else the oracle port.  Returns (seconds per step, nnz, kind, cores)."""
    import oracle
    from oracle import dp, ip
    nnz, row_ptr, cols, X = cpu_sample_matrix(ncol, R, seed, sample_rows, per_row)
    Y = np.zeros((sample_rows, R))
    if oracle.REF is not None:
        kind, cores = "reference", oracle.REF.ref_num_threads()
        run = lambda: oracle.REF.ref_bcsr_mul(32, dp(Y), sample_rows, ncol, nnz, ip(row_ptr), ip(cols), dp(X), R)
    else:
        kind, cores = "port", oracle.O.fso_num_threads()
        run = lambda: oracle.O.fso_csr_A_mul_Bn(dp(Y), sample_rows, ip(row_ptr), ip(cols), None, dp(X), R)
    for _ in range(max(warmup, 1)):
        run()
    t0 = time.perf_counter()
    for _ in range(steps):
        run()
    dt = (time.perf_counter() - t0) / steps
    return dt, nnz, kind, cores


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    # The reference arm uses every host thread it can: torchrun exports OMP_NUM_THREADS=1 to its workers, which
    # would silently time the reference's OpenMP loops on one core.  Set before the OpenMP runtime is loaded.
    usable = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    os.environ["OMP_NUM_THREADS"] = os.environ.get("FSB_REF_THREADS", str(usable))
    nrow, ncol, nnz, R, dist, seed = WORKLOADS[args.workload]
    sample_rows = min(nrow, args.sample_rows)
    dt, snnz, kind, cores = cpu_time_steps(ncol, R, seed, sample_rows, nnz / nrow, args.steps, args.warmup)
    value = snnz * R / dt
    sample = f"{sample_rows} rows x {ncol} cols, {snnz} nnz of the {args.workload} matrix (same row degree, full X), per step"
    line = {
        "impl": "reference", "metric": "spmm_nnz_rhs_per_s", "value": value, "unit": "nnz*RHS/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"{args.workload}: binary CSR {nrow}x{ncol}, {nnz} nnz, A_mul_Bn R={R}",
                   "function": "bcsr_A_mul_B32n (csr.h:283-302)" if kind == "reference" else "oracle port of bcsr_A_mul_B32n"},
        "cpu_baseline": {"value": value, "unit": "nnz*RHS/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "nnz*RHS/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# ------------------------------------------------------------------------------------------- GPU arm
def run_ours(args):
    import torch
    import torch.distributed as dist

    import libfastsparse_b200 as fs

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (libfastsparse_b200 has no CPU fallback)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    nrow, ncol, nnz, R, dkind, seed = WORKLOADS[args.workload]

    # every rank: its own C2-sized row shard (seed offset per rank), X replicated
    A = fs.DeviceMatrix.synth(seed + 1000003 * rank, dkind, nnz, nrow, ncol)
    c = torch.arange(ncol, device="cuda", dtype=torch.float64)[:, None]
    k = torch.arange(R, device="cuda", dtype=torch.float64)[None, :]
    X = torch.sin(7.0 * c + 17.0 * k + 0.3).reshape(-1).contiguous()
    Y = torch.empty(nrow * R, dtype=torch.float64, device="cuda")
    pinned = None
    if args.tune:      # pin the launch configuration instead of the per-handle autotune (profiling runs)
        vals = [int(v) for v in args.tune.split(",")]
        algo, tw, g, vec, slabs, rb = vals[:6]
        deep = vals[6] if len(vals) > 6 else 0
        fs.check(fs.lib().fsb_tune_csr_algo(algo, rb, 0))
        fs.check(fs.lib().fsb_tune_csr_spmm(tw, g, vec, slabs))
        fs.check(fs.lib().fsb_tune_csr_staged(deep))
        pinned = (R, max(slabs, 1), bool(deep))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        A.spmm(X, R, out=Y)
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    l0 = fs.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(args.steps):
        A.spmm(X, R, out=Y)
    ev1.record()
    barrier()
    launches = fs.launch_count() - l0
    tuned = pinned or A.tuning()
    clocks = sampler.stop() if sampler else None
    ms = ev0.elapsed_time(ev1) / args.steps
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    value = world * nnz * R / (ms * 1e-3)

    # end-to-end through the host-pointer C-ABI call (what the drop-in headers invoke):
    # pinned host X in, pinned host Y out, both copies inside the timed region
    e2e_steps = max(2, min(args.steps, 5))
    Xh = torch.empty(ncol * R, dtype=torch.float64).pin_memory()
    Xh.copy_(X.cpu())
    Yh = torch.empty(nrow * R, dtype=torch.float64).pin_memory()
    xp = C.cast(Xh.data_ptr(), C.POINTER(C.c_double)); yp = C.cast(Yh.data_ptr(), C.POINTER(C.c_double))
    fs.check(fs.lib().fsb_spmm_host(A.h, yp, xp, R))
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        fs.check(fs.lib().fsb_spmm_host(A.h, yp, xp, R))
    barrier()
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    if world > 1:
        t = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_value = world * nnz * R / e2e_s
    parity = float((Yh[: 64 * R].cuda() - Y[: 64 * R]).abs().max())

    if rank == 0:
        peak, peak_src = measured_peaks()
        ab = alg_bytes(nrow, nnz, R)
        achieved = ab / (ms * 1e-3) / 1e9
        traffic = None      # ncu DRAM bytes of one product in the configuration this run used (column passes)
        tp = os.path.join(ROOT, "profiles", "c2_spmm_traffic.json")
        if args.workload == "c2" and os.path.exists(tp):
            try:
                traffic = json.load(open(tp))["by_column_passes"][str(tuned[1])]["dram_bytes_per_product"]
            except Exception:
                traffic = None
        line = {
            "metric": "spmm_nnz_rhs_per_s", "value": value, "unit": "nnz*RHS/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"{args.workload}: binary CSR {nrow}x{ncol}, {nnz} nnz per GPU, A_mul_Bn R={R} (bcsr_A_mul_Bn / _B32n)",
                       "parallelism": f"row-sharded x{world}, X replicated, no collective", "l2": "inputs larger than L2 (no flush)",
                       "x_pattern": "sin(7c+17k+0.3)", "tune": args.tune or "auto",
                       "kernel": "csr_spmm_staged%s_kernel, %d column pass(es) per step" % ("_deep" if tuned[2] else "", tuned[1])},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "dram_gbs": (traffic / (ms * 1e-3) / 1e9) if traffic else None,
                         "dram_frac": (traffic / (ms * 1e-3) / 1e9 / peak) if traffic else None,
                         "alg_bytes_per_launch": ab, "launches_per_step": tuned[1],
                         "note": "achieved = algorithmic bytes of one product / time of one product (all its column-pass launches); "
                                 "traffic = ncu dram bytes of one product (profiles/c2_spmm_traffic.json), dram_gbs / dram_frac = that traffic / this run's time (frac > 1 on the "
                                 "algorithmic count means L2 served part of the X gathers, not that work was skipped)", "peak_source": peak_src,
                         "frac_of_nominal_8TBs": achieved / 8000.0},
            "e2e": {"value": e2e_value, "unit": "nnz*RHS/s", "h2d_bytes_per_step": ncol * R * 8, "d2h_bytes_per_step": nrow * R * 8,
                    "ms_per_step": e2e_s * 1e3, "api": "fsb_spmm_host (bcsr_A_mul_Bn drop-in path), pinned host buffers",
                    "check_vs_device_path_max_abs": parity},
            "gpu_launches": int(launches), "clocks": clocks,
        }
        if world == 1 and not args.no_cpu:
            sample_rows = 500_000
            dt, snnz, kind, cores = cpu_time_steps(ncol, R, seed, sample_rows, nnz / nrow, 5, 1)
            line["cpu_baseline"] = {"value": snnz * R / dt, "unit": "nnz*RHS/s", "cores": cores, "kind": kind,
                                    "sample": f"{sample_rows} rows x {ncol} cols, {snnz} nnz (same row degree, full X), "
                                              f"bcsr_A_mul_B32n, 5 passes after 1 warm-up, {dt * 1e3:.1f} ms/pass"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--tune", default="", help="algo,tw,g,vec,slabs,rb[,deep] override of the SpMM launch heuristic (see tools/sweep.py)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--sample-rows", type=int, default=1_000_000, help="--impl reference: rows of the workload each CPU step processes")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
