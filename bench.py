#!/usr/bin/env python
"""bench.py -- headline benchmark of the sparse x dense hot path on B200.

Metric (BASELINE.json): SpMM nnz*RHS/s (+ HBM roofline fraction) for the binary-CSR A_mul_Bn product with 32
right-hand sides on ONE synthetic 10M x 1M matrix with 200M nonzeros ("C2"), on 1/2/4/8 GPUs, next to the
reference's OpenMP CPU path.

  python bench.py --gpus N --steps K --warmup W            our arm (one rank per GPU under torchrun)
  python bench.py --impl reference --gpus N --steps K ...  the reference's own CPU code on the SAME matrix (rank 0 only)

A "step" is one full Y = A X pass over the whole matrix.
  N = 1: the matrix lives on one GPU.
  N > 1: STRONG scaling, as north_star states it -- the one matrix is cut by fsb_partition_rows into N nnz-balanced
         contiguous row shards (fsb_csr_row_slice), X is replicated, Y stays row-sharded: no data-path collective.
         value = nnz*R of the whole matrix / max-over-ranks time.  The same line carries a `collectives` block for the
         paths that DO communicate: C3 At_mul_B + allreduce, the C5 operator A'(A X) + lambda X and the
         C5 block-CG iteration, each with its single-GPU time measured in the same process, the speed-up, and parity
         (sharded vs single-GPU result, all ranks bit-identical) -- so the multi-GPU correctness check runs wherever
         the bench runs.  `replicas` keeps last round's weak-scaling figure (every rank its own full matrix).
Inputs are larger than L2 (cols 0.8 GB, X 256 MB, Y 2.56 GB at N = 1), so no L2 flush is needed between iterations;
for N > 1 the shards' inputs are still > L2 (X alone is 256 MB).

The reference arm never imports libfastsparse_b200: its inputs come from oracle/'s own restatement of the generator
(oracle/fsoracle.c fso_synth_coo; tests/test_oracle_golden.py checks both generators bit for bit).
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
L2_BYTES = 126e6


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


def compulsory_bytes(nrow, ncol, nnz, R, passes):
    """Every array once per column pass it is touched in: cols and row_ptr per pass, X and Y once."""
    return passes * (4 * nnz + 4 * (nrow + 1)) + 8 * ncol * R + 8 * nrow * R


def config_for(workload, world):
    """The `config` object -- identical in both arms so the driver can match them."""
    nrow, ncol, nnz, R, _, seed = WORKLOADS[workload]
    par = "one GPU" if world == 1 else (f"one matrix row-partitioned over {world} GPUs (nnz-balanced contiguous row shards), "
                                        "X replicated, Y row-sharded, no data-path collective")
    return {"workload": f"{workload}: binary CSR {nrow}x{ncol}, {nnz} nnz, A_mul_Bn R={R} (bcsr_A_mul_Bn / bcsr_A_mul_B32n, csr.h:257-302)",
            "parallelism": par, "x_pattern": "X[c][k] = sin(7c+17k+0.3) (bench_a_mul_b.c:149-152)", "seed": hex(seed),
            "l2": "inputs larger than L2 (no flush)"}


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
def cpu_build(workload, sample_rows):
    """The workload's matrix for the CPU legs, made by oracle/ alone (no product code): the synthetic COO stream and
    the reference's own new_bcsr (csr.h:30-67) when oracle/_ref is present, else the oracle port.
    sample_rows = 0: the full matrix.  sample_rows > 0 (bounded cpu_baseline leg): `sample_rows` rows with the
    workload's column count, row-degree distribution and dense operand."""
    import oracle
    from oracle import ip
    nrow, ncol, nnz, R, dist, seed = WORKLOADS[workload]
    if sample_rows and sample_rows < nrow:
        nnz = int(sample_rows * (nnz / nrow))
        nrow = sample_rows
    rows, cols, _ = oracle.synth_coo(seed, dist, nnz, nrow, ncol)
    if oracle.REF is not None:
        row_ptr = np.zeros(nrow + 1, np.int32)
        ccols = np.zeros(max(nnz, 1), np.int32)
        oracle.REF.ref_new_bcsr(nnz, nrow, ncol, ip(rows), ip(cols), ip(row_ptr), ip(ccols))
        ccols = ccols[:nnz]
    else:
        row_ptr, ccols, _ = oracle.csr_from_coo(nrow, rows, cols)
    del rows, cols
    X = np.ascontiguousarray(np.sin(7.0 * np.arange(ncol)[:, None] + 17.0 * np.arange(R)[None, :] + 0.3))
    return nrow, ncol, nnz, R, row_ptr, np.ascontiguousarray(ccols), X


def cpu_time_steps(workload, sample_rows, steps, warmup):
    """Times the reference's own bcsr_A_mul_B32n (csr.h:283-302, unmodified, OpenMP, all host threads) when
    oracle/_ref is present, else the oracle port.  Returns (seconds per step, nrow, nnz, kind, cores)."""
    import oracle
    from oracle import dp, ip
    nrow, ncol, nnz, R, row_ptr, cols, X = cpu_build(workload, sample_rows)
    Y = np.zeros((nrow, R))
    if oracle.REF is not None:
        kind, cores = "reference", oracle.REF.ref_num_threads()
        run = lambda: oracle.REF.ref_bcsr_mul(32, dp(Y), nrow, ncol, nnz, ip(row_ptr), ip(cols), dp(X), R)
    else:
        kind, cores = "port", oracle.O.fso_num_threads()
        run = lambda: oracle.O.fso_csr_A_mul_Bn(dp(Y), nrow, ip(row_ptr), ip(cols), None, dp(X), R)
    for _ in range(max(warmup, 1)):
        run()
    t0 = time.perf_counter()
    for _ in range(steps):
        run()
    dt = (time.perf_counter() - t0) / steps
    return dt, nrow, nnz, kind, cores


def set_host_threads():
    # torchrun exports OMP_NUM_THREADS=1 to its workers, which would silently time the reference's OpenMP loops on
    # one core.  Set before the OpenMP runtime is loaded.
    usable = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    os.environ["OMP_NUM_THREADS"] = os.environ.get("FSB_REF_THREADS", str(usable))


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    set_host_threads()
    wnrow, ncol, wnnz, R, _, _ = WORKLOADS[args.workload]
    dt, nrow, nnz, kind, cores = cpu_time_steps(args.workload, args.sample_rows, args.steps, args.warmup)
    value = nnz * R / dt
    full = nrow == wnrow
    sample = (f"the full {args.workload} matrix ({nrow} rows, {nnz} nnz) per step" if full else
              f"{nrow} rows x {ncol} cols, {nnz} nnz of the {args.workload} matrix (same row degree, full X), per step")
    line = {
        "impl": "reference", "metric": "spmm_nnz_rhs_per_s", "value": value, "unit": "nnz*RHS/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_for(args.workload, args.gpus),
        "detail": {"function": "bcsr_A_mul_B32n (csr.h:283-302), unmodified reference, OpenMP" if kind == "reference"
                   else "oracle port of bcsr_A_mul_B32n", "structure": "new_bcsr (csr.h:30-67) on the synthetic COO",
                   "whole_workload": full},
        "cpu_baseline": {"value": value, "unit": "nnz*RHS/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "nnz*RHS/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# ------------------------------------------------------------------------------------------- GPU arm
def traffic_for(workload, kernel, passes):
    """ncu DRAM bytes of one product for exactly this kernel build and pass count (profiles/c2_spmm_traffic.json),
    or (None, reason) -- never a number measured on a different kernel."""
    tp = os.path.join(ROOT, "profiles", "c2_spmm_traffic.json")
    if workload != "c2":
        return None, "no ncu capture for this workload"
    try:
        d = json.load(open(tp))
        e = d["by_kernel"][f"{kernel}/{passes}"]
        return float(e["dram_bytes_per_product"]), f"{e['source']} ({kernel}, {passes} column pass(es))"
    except (OSError, KeyError, ValueError) as ex:
        return None, f"no ncu capture of {kernel}/{passes} in profiles/c2_spmm_traffic.json ({type(ex).__name__})"


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
        fs.comm_init_from_torch()
    nrow, ncol, nnz, R, dkind, seed = WORKLOADS[args.workload]
    L = fs.lib()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v):
        if world == 1:
            return float(v)
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- the ONE matrix; for N > 1 every rank builds it on its device and keeps its nnz-balanced row shard
    full = fs.DeviceMatrix.synth(seed, dkind, nnz, nrow, ncol)
    r0, r1 = 0, nrow
    A = full
    if world > 1:
        rp, _, _ = full.download_csr()
        b = fs.partition_rows(rp, world)
        r0, r1 = int(b[rank]), int(b[rank + 1])
        A = full.row_slice(r0, r1)
        del rp
    nloc = r1 - r0
    c = torch.arange(ncol, device="cuda", dtype=torch.float64)[:, None]
    k = torch.arange(R, device="cuda", dtype=torch.float64)[None, :]
    X = torch.sin(7.0 * c + 17.0 * k + 0.3).reshape(-1).contiguous()
    del c, k
    Y = torch.empty(nloc * R, dtype=torch.float64, device="cuda")
    pinned = None
    if args.tune:      # pin the launch configuration instead of the per-handle autotune (profiling runs)
        vals = [int(v) for v in args.tune.split(",")]
        algo, tw, g, vec, slabs, rb = vals[:6]
        deep = vals[6] if len(vals) > 6 else 0
        fs.check(L.fsb_tune_csr_algo(algo, rb, 0))
        fs.check(L.fsb_tune_csr_spmm(tw, g, vec, slabs))
        fs.check(L.fsb_tune_csr_staged(deep))
        pinned = (R, max(slabs, 1), bool(deep))

    def timed_steps(M, Xd, Yd, steps, warmup):
        for _ in range(warmup):
            M.spmm(Xd, R, out=Yd)
        barrier()
        l0 = fs.launch_count()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        ev0.record()
        for _ in range(steps):
            M.spmm(Xd, R, out=Yd)
        ev1.record()
        barrier()
        return max_over_ranks(ev0.elapsed_time(ev1) / steps), fs.launch_count() - l0

    warm = max(args.warmup, 3)
    for _ in range(2):
        A.spmm(X, R, out=Y)      # first product: per-handle autotune (outside every timed region)
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    ms, launches = timed_steps(A, X, Y, args.steps, warm)
    clocks = sampler.stop() if sampler else None
    tuned = pinned or A.tuning()
    value = nnz * R / (ms * 1e-3)          # the whole matrix, whatever N

    # ---- end to end through the host-pointer C-ABI call (what the drop-in headers invoke): host X in, host Y (this
    # rank's row slab) out, both copies inside the timed region.  Pinned buffers first, then plain malloc'd ones.
    e2e_steps = max(2, min(args.steps, 5))
    Xh = torch.empty(ncol * R, dtype=torch.float64).pin_memory()
    Xh.copy_(X.cpu())
    Yh = torch.empty(nloc * R, dtype=torch.float64).pin_memory()
    xp = C.cast(Xh.data_ptr(), C.POINTER(C.c_double)); yp = C.cast(Yh.data_ptr(), C.POINTER(C.c_double))
    if world > 1:
        A.set_row_sharded(True)      # sharded handle: every rank uploads 1/N of X, the rest arrives over NVLink (all-gather)
        fs.check(L.fsb_tune(b"host_x_allgather", 1))
    fs.check(L.fsb_spmm_host(A.h, yp, xp, R))
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        fs.check(L.fsb_spmm_host(A.h, yp, xp, R))
    barrier()
    e2e_s = max_over_ranks((time.perf_counter() - t0) / e2e_steps)
    # parity of the host path against the device path over a strided sample of ALL of this rank's rows
    stride = max(1, nloc // 65536)
    idx = torch.arange(0, nloc, stride)
    got = Yh.view(nloc, R)[idx].cuda()
    want = Y.view(nloc, R)[idx.cuda()]
    parity = max_over_ranks(float((got - want).abs().max()) if nloc else 0.0)
    # pageable (malloc'd) operands: what a C caller of the reference passes (bench_a_mul_b.c:125-139)
    Xm = np.empty(ncol * R, dtype=np.float64); Xm[:] = Xh.numpy()
    Ym = np.empty(max(nloc * R, 1), dtype=np.float64); Ym[:] = 0.0      # touched once: page faults are not the product's
    xmp = Xm.ctypes.data_as(C.POINTER(C.c_double)); ymp = Ym.ctypes.data_as(C.POINTER(C.c_double))
    fs.check(L.fsb_spmm_host(A.h, ymp, xmp, R))
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        fs.check(L.fsb_spmm_host(A.h, ymp, xmp, R))
    barrier()
    pageable_s = max_over_ranks((time.perf_counter() - t0) / e2e_steps)
    pageable_parity = max_over_ranks(float(np.abs(Ym[: nloc * R].reshape(nloc, R)[::stride] - Yh.view(nloc, R)[idx].numpy()).max()) if nloc else 0.0)
    if world > 1:
        A.set_row_sharded(False)
    h2d_step = (ncol * R * 8) // world if world > 1 else ncol * R * 8
    del Xm, Ym, Yh

    # ---- N > 1 extras: last round's replica (weak-scaling) figure, and the paths that communicate
    extras = {}
    if world > 1:
        Yfull = torch.empty(nrow * R, dtype=torch.float64, device="cuda")
        full.spmm(X, R, out=Yfull)
        ms_rep, _ = timed_steps(full, X, Yfull, max(3, args.steps // 4), 2)
        slab_equal = bool(torch.equal(Yfull[r0 * R: r1 * R], Y)) if nloc else True
        flag = torch.tensor([1 if slab_equal else 0], device="cuda"); dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        extras["replicas"] = {"what": "every rank multiplies its own copy of the full matrix (weak scaling, last round's headline)",
                              "value": world * nnz * R / (ms_rep * 1e-3), "unit": "nnz*RHS/s", "ms_per_step": ms_rep}
        extras["shard_parity"] = {"row_slab_bit_identical_to_single_gpu_product": bool(flag.item())}
        del Yfull
    full_handle = full if world > 1 else None
    del Y
    if world > 1 and not args.no_collectives:
        try:
            extras["collectives"] = collectives_block(args, fs, torch, dist, world, rank, full_handle, A, r0, r1, max_over_ranks, barrier)
        except Exception as ex:      # deterministic failures hit every rank alike; the headline line must survive them
            extras["collectives"] = {"error": f"{type(ex).__name__}: {ex}"[:500]}
    if full_handle is not None:
        full_handle.free()

    if rank == 0:
        peak, peak_src = measured_peaks()
        kernel = "csr_spmm_staged%s_kernel" % ("_deep" if tuned[2] else "")
        passes = int(tuned[1])
        t_s = ms * 1e-3
        ab = alg_bytes(nrow, nnz, R)
        cb = compulsory_bytes(nrow, ncol, nnz, R, passes) + (world - 1) * 8 * ncol * R      # X is read once per GPU
        traffic, traffic_src = traffic_for(args.workload, kernel, passes) if world == 1 else (None, "ncu capture is single-GPU")
        dram_gbs = traffic / t_s / 1e9 if traffic else None
        cfg = config_for(args.workload, world)
        line = {
            "metric": "spmm_nnz_rhs_per_s", "value": value, "unit": "nnz*RHS/s", "n_gpus": world, "steps": args.steps,
            "warmup": warm, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": cfg,
            "detail": {"kernel": f"{kernel}, {passes} column pass(es) per step", "tune": args.tune or "auto (timed once per handle, outside the timed region)",
                       "rows_per_rank": nloc if world > 1 else nrow},
            "roofline": {
                "bound": "hbm", "unit": "GB/s", "peak": peak * world, "peak_source": peak_src + (f" x {world} GPUs" if world > 1 else ""),
                # frac = REAL DRAM traffic of one product (ncu) / this run's time / peak: a physical fraction, <= 1.
                # When this exact kernel has no ncu capture, the compulsory bytes stand in (a lower bound on traffic).
                "achieved": dram_gbs if traffic else cb / t_s / 1e9,
                "frac": (dram_gbs if traffic else cb / t_s / 1e9) / (peak * world),
                "basis": "ncu dram bytes of one product" if traffic else "compulsory bytes (no ncu capture of this kernel: lower bound)",
                "traffic": traffic, "traffic_source": traffic_src,
                "alg_bytes_per_step": ab, "alg_achieved": ab / t_s / 1e9, "alg_frac": ab / t_s / 1e9 / (peak * world),
                "compulsory_bytes_per_step": cb, "compulsory_frac": cb / t_s / 1e9 / (peak * world),
                "launches_per_step": passes,
                "note": "alg_* = SURVEY 8(d) algorithmic bytes (every X gather counted as HBM traffic, so alg_frac can exceed 1 when L2 "
                        "serves gathers); compulsory_* = every array once per pass it is touched in; the product is bound by the L2 "
                        "gather throughput (profiles/r2_gather_ceiling.md), not by DRAM",
            },
            "e2e": {"value": nnz * R / e2e_s, "unit": "nnz*RHS/s", "h2d_bytes_per_step": h2d_step, "d2h_bytes_per_step": nloc * R * 8,
                    "ms_per_step": e2e_s * 1e3, "api": "fsb_spmm_host (what bcsr_A_mul_Bn in include/fastsparse/csr.h calls), pinned host buffers"
                    + ("; per rank: 1/N of X up (rest by NVLink all-gather), its row slab of Y down" if world > 1 else ""),
                    "check_vs_device_path_max_abs": parity, "check_rows": f"every {stride}th row of all rows",
                    "pageable_ms": pageable_s * 1e3, "pageable_value": nnz * R / pageable_s,
                    "pageable_note": "same call with malloc'd (pageable) X and Y: pinned bounce ring + host copy threads (fsb_hostcopy.cu)",
                    "pageable_check_max_abs": pageable_parity},
            "gpu_launches": int(launches), "clocks": clocks,
        }
        if world > 1:
            # context only: the single-GPU ncu traffic of the same kernel build, assumed to split evenly over the row shards
            t1, src1 = traffic_for(args.workload, kernel, passes)
            if t1:
                line["roofline"]["estimate_from_1gpu_capture"] = {
                    "traffic": t1, "achieved": t1 / t_s / 1e9, "frac": t1 / t_s / 1e9 / (peak * world),
                    "note": "NOT measured at this N: " + src1 + "; every shard runs the same kernel on 1/N of the rows against the full X"}
        line.update(extras)
        if world == 1 and not args.no_cpu:
            # the C caller: tests/_build/time_dropin (struct BinaryCSR + malloc'd operands + bcsr_A_mul_Bn through the header)
            exe = os.path.join(ROOT, "tests", "_build", "time_dropin")
            if os.path.exists(exe):
                del A, X
                full.free()
                torch.cuda.empty_cache()
                try:
                    r = subprocess.run([exe, str(nrow), str(ncol), str(nnz), str(R), "5", hex(seed)], capture_output=True, text=True, timeout=600)
                    line["e2e"]["dropin_c"] = json.loads(r.stdout.strip().splitlines()[-1]) if r.returncode == 0 else {"error": (r.stderr or r.stdout)[-300:]}
                except Exception as ex:      # the C leg is informative; never lose the line over it
                    line["e2e"]["dropin_c"] = {"error": repr(ex)[:300]}
            set_host_threads()
            sample_rows = min(nrow, 500_000)
            dt, srows, snnz, kind, cores = cpu_time_steps(args.workload, sample_rows, 5, 1)
            line["cpu_baseline"] = {"value": snnz * R / dt, "unit": "nnz*RHS/s", "cores": cores, "kind": kind,
                                    "sample": f"{srows} rows x {ncol} cols, {snnz} nnz (same row degree, full X), "
                                              f"bcsr_A_mul_B32n, 5 passes after 1 warm-up, {dt * 1e3:.1f} ms/pass"}
        print(json.dumps(line), flush=True)
    if world > 1:
        fs.comm_finalize()
        dist.destroy_process_group()
    return 0


def collectives_block(args, fs, torch, dist, world, rank, full_c2, shard_c2, r0, r1, max_over_ranks, barrier):
    """The row-sharded paths with a real exchange step, at this N, next to their single-GPU time measured in
    the same process on the full matrix, with parity.  Times: CUDA events after a barrier, max over ranks."""
    nrow, ncol, nnz, R, _, _ = WORKLOADS[args.workload]
    L = fs.lib()
    reps = 5

    def timed(fn, reps, sync_ranks):
        for _ in range(2):
            fn()
        if sync_ranks:
            barrier()
        else:
            torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        if sync_ranks:
            barrier()
        else:
            torch.cuda.synchronize()
        return max_over_ranks(e0.elapsed_time(e1) / reps)

    def ranks_agree(t):
        s = t.double().sum().reshape(1).clone()
        lo, hi = s.clone(), s.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        return bool(lo.item() == hi.item())

    def rel_err(a, b):
        return max_over_ranks(float((a - b).abs().max() / b.abs().max()))

    out = {"n_gpus": world, "timing": "CUDA events, max over ranks; single-GPU times measured in the same process on the full matrix"}
    # ---- C3: double CSR, y = A'x with the [F] partial sum-allreduced (sdm_At_mul_B, dsparse.h:54-62)
    fullv = fs.DeviceMatrix.synth(0x5EED0003, 0, nnz, nrow, ncol, with_vals=True)
    rp, _, _ = fullv.download_csr()
    b = fs.partition_rows(rp, world)
    q0, q1 = int(b[rank]), int(b[rank + 1])
    del rp
    sh = fullv.row_slice(q0, q1)
    sh.set_row_sharded(True)
    x = (torch.sin(7.0 * torch.arange(ncol, device="cuda", dtype=torch.float64) + 0.3) / 10).contiguous()
    yfull = fullv.spmm(x, 1)
    z1 = torch.empty(ncol, dtype=torch.float64, device="cuda"); zs = torch.empty_like(z1)
    ms1 = timed(lambda: fullv.spmm_t(yfull, 1, out=z1), reps, False)
    ysh = yfull[q0:q1].contiguous()
    msn = timed(lambda: sh.spmm_t(ysh, 1, out=zs), reps, True)
    out["c3_At_mul_B_allreduce"] = {"what": "double CSR 10Mx1M 200M nnz, y = A'x: per-shard partial + NCCL sum-allreduce of [F] fp64",
                                    "ms_1gpu": ms1, "ms": msn, "speedup": ms1 / msn, "nnz_per_s": nnz / msn * 1e3,
                                    "allreduce_bytes": ncol * 8,
                                    "parity": {"max_rel_err_vs_1gpu": rel_err(zs, z1), "ranks_bit_identical": ranks_agree(zs)}}
    sh.free(); fullv.free()
    del fullv, sh, x, yfull, ysh, z1, zs
    torch.cuda.empty_cache()
    # ---- C5: operator A'(A X) + lambda X (bsbm_AtA, cg.h:9-22) and the block-CG iteration, R = 32, on the C2 matrix
    shard_c2.set_row_sharded(True)
    g = torch.Generator(device="cuda"); g.manual_seed(5)
    Bm = torch.randn(ncol * R, dtype=torch.float64, device="cuda", generator=g)      # same on every rank (same seed)
    K1 = torch.empty(ncol * R, dtype=torch.float64, device="cuda"); Ks = torch.empty_like(K1)
    full_c2.ata(Bm, R, lam=15.0, out=K1)
    ms1 = timed(lambda: full_c2.ata(Bm, R, lam=15.0, out=K1), reps, False)
    shard_c2.ata(Bm, R, lam=15.0, out=Ks)
    msn = timed(lambda: shard_c2.ata(Bm, R, lam=15.0, out=Ks), reps, True)
    out["c5_operator"] = {"what": "A'(A X) + lambda X, R=32, two gather passes per shard + NCCL sum-allreduce of the [F][32] partial",
                          "ms_1gpu": ms1, "ms": msn, "speedup": ms1 / msn, "nnz_rhs_per_s": 2 * nnz * R / msn * 1e3,
                          "allreduce_bytes": ncol * R * 8,
                          "parity": {"max_rel_err_vs_1gpu": rel_err(Ks, K1), "ranks_bit_identical": ranks_agree(Ks)}}
    del K1, Ks

    def solve(M, sync_ranks):
        M.cg(Bm, R, lam=15.0, tol=1e-30, max_iter=2)        # warm-up: transposes, autotune, workspace
        if sync_ranks:
            barrier()
        else:
            torch.cuda.synchronize()
        t0 = time.perf_counter()
        Xs, it = M.cg(Bm, R, lam=15.0, tol=1e-6)
        if sync_ranks:
            barrier()
        else:
            torch.cuda.synchronize()
        return Xs, it, max_over_ranks(time.perf_counter() - t0)

    X1, it1, s1 = solve(full_c2, False)
    fs.check(L.fsb_tune_cg_dist(0))
    Xn, itn, sn = solve(shard_c2, True)
    res = (shard_c2.ata(Xn, R, lam=15.0) - Bm).reshape(ncol, R).norm(dim=0) / Bm.reshape(ncol, R).norm(dim=0)
    per1, pern = s1 / (it1 + 1) * 1e3, sn / (itn + 1) * 1e3
    G = world
    out["c5_block_cg"] = {"what": "block CG (lambda I + A'A) X = B, R=32, lambda=15, tol=1e-6; A row-sharded, CG vectors sharded over F: "
                                  "reduce-scatter of the A'(A P) partial overlapped with its computation, all-gather of P, allreduce of the R x R Grams",
                          "iterations_1gpu": it1, "iterations": itn, "ms_per_iteration_1gpu": per1, "ms_per_iteration": pern,
                          "speedup": per1 / pern, "nnz_rhs_per_s": 2 * nnz * R / pern * 1e3,
                          "nvlink_bytes_per_iteration_per_rank": {"reduce_scatter_in": (G - 1) * ncol * R * 8 // G,
                                                                  "all_gather_in": (G - 1) * ncol * R * 8 // G, "gram_allreduce": 2 * R * R * 8},
                          "parity": {"max_rel_err_vs_1gpu": rel_err(Xn, X1), "max_rel_residual": float(res.max()),
                                     "ranks_bit_identical": ranks_agree(Xn)}}
    shard_c2.set_row_sharded(False)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--tune", default="", help="algo,tw,g,vec,slabs,rb[,deep] override of the SpMM launch heuristic (see tools/sweep.py)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline and C-caller legs")
    ap.add_argument("--no-collectives", action="store_true", help="N > 1: skip the collectives block")
    ap.add_argument("--sample-rows", type=int, default=0,
                    help="--impl reference: rows of the workload each CPU step processes (0 = the whole matrix, the default)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
