"""Measure every BASELINE.json config on one GPU (C2-C5) with the kernel's roofline and,
where cheap, the reference CPU function on a bounded sample.  One JSON line per measurement.

    python tools/bench_all.py [--small] [--out gpurun_out/bench_all.jsonl]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import libfastsparse_b200 as fs  # noqa: E402
from bench import measured_peaks  # noqa: E402

L2 = 126e6


def timed(fn, reps, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def spmm_bytes(nnz, nout, nin, R, idx_bytes, extra=0):
    """matrix arrays, output and row_ptr once; dense input per gather when it exceeds L2, else once."""
    dense = nnz * 8 * R if 8 * nin * R > L2 else 8 * nin * R
    return nnz * idx_bytes + dense + 4 * (nout + 1) + 8 * nout * R + extra


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--small", action="store_true")
    ap.add_argument("--out", default="")
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--only", default="")
    args = ap.parse_args()
    N, F, NNZ = (1_000_000, 100_000, 20_000_000) if args.small else (10_000_000, 1_000_000, 200_000_000)
    peak, _ = measured_peaks()
    out = open(args.out, "w") if args.out else None

    def emit(**kw):
        kw["frac_of_measured_peak"] = kw["alg_gbs"] / peak if "alg_gbs" in kw else None
        line = json.dumps(kw)
        print(line, flush=True)
        if out:
            out.write(line + "\n"); out.flush()

    want = set(args.only.split(",")) if args.only else None
    on = lambda k: want is None or k in want

    # ---------------- C3: double CSR SpMV + At_mul_B
    if on("c3"):
        A = fs.DeviceMatrix.synth(0x5EED0003, 0, NNZ, N, F, with_vals=True)
        x = (torch.sin(7.0 * torch.arange(F, device="cuda", dtype=torch.float64) + 0.3) / 10).contiguous()
        y = torch.empty(N, dtype=torch.float64, device="cuda")
        ms = timed(lambda: A.spmm(x, 1, out=y), args.reps)
        ab = spmm_bytes(NNZ, N, F, 1, 12)
        emit(config="C3 double CSR SpMV (csr_A_mul_B)", ms=ms, nnz_per_s=NNZ / ms * 1e3, alg_bytes=ab, alg_gbs=ab / ms / 1e6)
        z = torch.empty(F, dtype=torch.float64, device="cuda")
        A.spmm_t(y, 1, out=z)
        ms = timed(lambda: A.spmm_t(y, 1, out=z), args.reps)
        ab = spmm_bytes(NNZ, F, N, 1, 12)
        emit(config="C3 double CSR At_mul_B (stored transpose, sdm_At_mul_B)", ms=ms, nnz_per_s=NNZ / ms * 1e3, alg_bytes=ab, alg_gbs=ab / ms / 1e6)
        for mode in (0, 1):
            ms = timed(lambda: A.ata(x, 1, mode=mode, out=z), args.reps)
            ab = (2 * 12 * NNZ + 4 * (N + F + 2) + 16 * N + 16 * F) if mode == 0 else (12 * NNZ + 4 * (N + 1) + 16 * F)
            emit(config=f"C3 double CSR A'(A x) mode {mode} ({'two gather passes' if mode == 0 else 'fused red.add scatter'})", ms=ms,
                 nnz_visits_per_s=2 * NNZ / ms * 1e3, alg_bytes=ab, alg_gbs=ab / ms / 1e6)
        # binary SpMV on the same structure (bcsr_A_mul_B) and fused A'A (parallel_bcsr_AA_mul_B)
        del A
        B = fs.DeviceMatrix.synth(0x5EED0003, 0, NNZ, N, F)
        ms = timed(lambda: B.spmm(x, 1, out=y), args.reps)
        ab = spmm_bytes(NNZ, N, F, 1, 4)
        emit(config="binary CSR SpMV (bcsr_A_mul_B)", ms=ms, nnz_per_s=NNZ / ms * 1e3, alg_bytes=ab, alg_gbs=ab / ms / 1e6)
        for mode in (0, 1):
            ms = timed(lambda: B.ata(x, 1, mode=mode, out=z), args.reps)
            ab = (2 * 4 * NNZ + 4 * (N + F + 2) + 16 * N + 16 * F) if mode == 0 else (8 * NNZ + 4 * (N + 1) + 16 * F)
            emit(config=f"binary CSR A'(A x) mode {mode} (bcsr_AA_mul_B / parallel_bcsr_AA_mul_B)", ms=ms,
                 nnz_visits_per_s=2 * NNZ / ms * 1e3, alg_bytes=ab, alg_gbs=ab / ms / 1e6)
        for R in (2, 4, 8, 16):
            X = torch.randn(F * R, dtype=torch.float64, device="cuda"); Y = torch.empty(N * R, dtype=torch.float64, device="cuda")
            ms = timed(lambda: B.spmm(X, R, out=Y), args.reps)
            ab = spmm_bytes(NNZ, N, F, R, 4)
            emit(config=f"binary CSR SpMM R={R} (bcsr_A_mul_B{R if R <= 8 else 'n'})", ms=ms, nnz_rhs_per_s=NNZ * R / ms * 1e3, alg_bytes=ab, alg_gbs=ab / ms / 1e6)
            del X, Y
        del B

    # ---------------- C4: power-law columns, R = 32, blocked (Hilbert) and column-blocked formats
    if on("c4"):
        R = 32
        M = fs.DeviceMatrix.synth(0x5EED0004, 1, NNZ, N, F, keep_coo=True)
        rows, cols, _ = M.coo
        X = torch.randn(F * R, dtype=torch.float64, device="cuda"); Y = torch.empty(N * R, dtype=torch.float64, device="cuda")
        ms = timed(lambda: M.spmm(X, R, out=Y), args.reps)
        ab = spmm_bytes(NNZ, N, F, R, 4)
        emit(config="C4 power-law cols, binary CSR SpMM R=32", ms=ms, nnz_rhs_per_s=NNZ * R / ms * 1e3, alg_bytes=ab, alg_gbs=ab / ms / 1e6)
        Yref = Y.clone()
        for bs in (512, 256):
            t0 = time.perf_counter()
            Bk = fs.DeviceMatrix.blocked_from_coo_tensors(N, F, rows, cols, None, bs, order=1)
            torch.cuda.synchronize(); build_s = time.perf_counter() - t0
            # the default product path runs the CSR kernels on a row-stable CSR view built on first use: its one-off
            # cost (time of the first product minus a steady-state one) and the HBM it adds are part of the picture
            b0 = Bk.bytes(); t0 = time.perf_counter()
            Bk.spmm(X, R, out=Y); torch.cuda.synchronize()
            first_s = time.perf_counter() - t0
            ms = timed(lambda: Bk.spmm(X, R, out=Y), args.reps)
            ab = NNZ * (8 + 8 * R) + 8 * N * R
            emit(config=f"C4 BlockedSBM bs={bs} Hilbert-sorted (bsbm_A_mul_Bn R=32), CSR view (default)", ms=ms, nnz_rhs_per_s=NNZ * R / ms * 1e3, alg_bytes=ab,
                 alg_gbs=ab / ms / 1e6, maxdiff_vs_csr=float((Y - Yref).abs().max()), device_build_s=build_s,
                 view_build_and_autotune_s=first_s - ms * 1e-3, view_extra_hbm_bytes=Bk.bytes() - b0, handle_hbm_bytes=Bk.bytes())
            if bs == 512:   # the format's own kernel (opt-in, fsb_tune_formats(1)): kept for comparison
                fs.check(fs.lib().fsb_tune_formats(1))
                ms_n = timed(lambda: Bk.spmm(X, R, out=Y), max(3, args.reps // 2))
                fs.check(fs.lib().fsb_tune_formats(0))
                emit(config=f"C4 BlockedSBM bs={bs} Hilbert-sorted, native blocked_spmm_kernel (opt-in)", ms=ms_n, nnz_rhs_per_s=NNZ * R / ms_n * 1e3,
                     alg_bytes=ab, alg_gbs=ab / ms_n / 1e6, maxdiff_vs_csr=float((Y - Yref).abs().max()))
            del Bk
        t0 = time.perf_counter()
        Cb = fs.DeviceMatrix.cbcsr_from_coo_tensors(N, F, rows, cols, 65536)
        torch.cuda.synchronize(); build_s = time.perf_counter() - t0
        b0 = Cb.bytes(); t0 = time.perf_counter()
        Cb.spmm(X, R, out=Y); torch.cuda.synchronize()
        first_s = time.perf_counter() - t0
        ms = timed(lambda: Cb.spmm(X, R, out=Y), args.reps)
        ab = NNZ * (4 + 8 * R) + 4 * (Cb.nblocks * N + 1) + 8 * N * R
        emit(config="C4 ColBinaryCSR colblock=65536 (cbcsr_A_mul_Bn R=32), CSR view (default)", ms=ms, nnz_rhs_per_s=NNZ * R / ms * 1e3, alg_bytes=ab, alg_gbs=ab / ms / 1e6,
             maxdiff_vs_csr=float((Y - Yref).abs().max()), device_build_s=build_s, nblocks=Cb.nblocks,
             view_build_and_autotune_s=first_s - ms * 1e-3, view_extra_hbm_bytes=Cb.bytes() - b0, handle_hbm_bytes=Cb.bytes())
        fs.check(fs.lib().fsb_tune_formats(1))
        ms_n = timed(lambda: Cb.spmm(X, R, out=Y), max(3, args.reps // 2))
        fs.check(fs.lib().fsb_tune_formats(0))
        emit(config="C4 ColBinaryCSR colblock=65536, native cbcsr_spmm_kernel (opt-in)", ms=ms_n, nnz_rhs_per_s=NNZ * R / ms_n * 1e3, alg_bytes=ab,
             alg_gbs=ab / ms_n / 1e6, maxdiff_vs_csr=float((Y - Yref).abs().max()))
        # the C4 preprocessing pipeline of the reference (bench_a_mul_b.c:173-205): sort_sbm -> new_bsbm(512) -> sort_bsbm.
        # Device: global Hilbert sort of the COO in HBM, then the blocked builder with per-block Hilbert order.
        # Host (the reference's serial algorithm, bit-exact twin in fsb_host.cpp) timed on a 1/20 sample, scaled linearly
        # (the sorts are O(n log n): the scaled figure is a lower bound).
        rs, cs = rows.clone(), cols.clone()
        torch.cuda.synchronize(); t0 = time.perf_counter()
        fs.check(fs.lib().fsb_sort_coo_hilbert_dev(N, F, NNZ, rs.data_ptr(), cs.data_ptr(), None))
        torch.cuda.synchronize(); sort_s = time.perf_counter() - t0
        t0 = time.perf_counter()
        Bk2 = fs.DeviceMatrix.blocked_from_coo_tensors(N, F, rs, cs, None, 512, order=1)
        torch.cuda.synchronize(); blk_s = time.perf_counter() - t0
        del Bk2
        ns = NNZ // 20
        hr = rows[:ns].cpu().numpy().copy(); hc = cols[:ns].cpu().numpy().copy()
        t0 = time.perf_counter()
        fs.check(fs.lib().fsb_host_sort_coo_hilbert(N, F, ns, fs.api._ip(hr), fs.api._ip(hc), None))
        host_sort_s = time.perf_counter() - t0
        A_h = fs.new_sbm(N, F, ns, hr, hc)
        t0 = time.perf_counter(); B_h = fs.new_bsbm(A_h, 512); host_blk_s = time.perf_counter() - t0
        B_d = fs.new_bsbm(A_h, 512)                       # the drop-in call on a HOST structure, through the device
        t0 = time.perf_counter(); fs.sort_bsbm(B_h, device=False); host_bsort_s = time.perf_counter() - t0
        t0 = time.perf_counter(); fs.sort_bsbm(B_d, device=True); dropin_dev_bsort_s = time.perf_counter() - t0
        same = all(np.array_equal(B_h.rows[b], B_d.rows[b]) and np.array_equal(B_h.cols[b], B_d.cols[b]) for b in range(B_h.nblocks))
        emit(config="C4 preprocessing: sort_sbm -> new_bsbm(512) -> sort_bsbm", device_sort_sbm_s=sort_s, device_new_bsbm_plus_sort_bsbm_s=blk_s,
             device_total_s=sort_s + blk_s, host_sample_nnz=ns, host_sort_sbm_s_sample=host_sort_s, host_new_bsbm_s_sample=host_blk_s,
             host_sort_bsbm_s_sample=host_bsort_s, host_total_s_scaled_to_full=(host_sort_s + host_blk_s + host_bsort_s) * (NNZ / ns),
             dropin_sort_bsbm_on_host_struct_via_device_s_sample=dropin_dev_bsort_s, dropin_device_sort_matches_host=bool(same))
        del rs, cs, A_h, B_h, B_d, hr, hc
        x = torch.randn(F, dtype=torch.float64, device="cuda"); y = torch.empty(N, dtype=torch.float64, device="cuda")
        ms = timed(lambda: Cb.spmm(x, 1, out=y), args.reps)
        ab = 4 * NNZ + 4 * (Cb.nblocks * N + 1) + 8 * N + 8 * F
        emit(config="C4 ColBinaryCSR colblock=65536 (cbcsr_A_mul_B R=1)", ms=ms, nnz_per_s=NNZ / ms * 1e3, alg_bytes=ab, alg_gbs=ab / ms / 1e6)
        del Cb, M, X, Y, Yref, rows, cols

    # ---------------- C5: block CG (lambda I + A'A) X = B, R = 32, on the C2 matrix
    if on("c5"):
        R = 32
        M = fs.DeviceMatrix.synth(0x5EED0002, 0, NNZ, N, F)
        g = torch.Generator(device="cuda"); g.manual_seed(5)
        Nn = torch.randn(N * R, dtype=torch.float64, device="cuda", generator=g)
        E = torch.randn(F * R, dtype=torch.float64, device="cuda", generator=g)
        Bm = M.spmm_t(Nn, R) + (15.0 ** 0.5) * E          # B = A'N + sqrt(lambda) E   (bench_a_mul_b.c:334-347)
        del Nn, E
        M.cg(Bm, R, lam=15.0, tol=1e-30, max_iter=2)     # warm-up: per-handle autotune of A (A' was tuned by spmm_t above)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        Xs, it = M.cg(Bm, R, lam=15.0, tol=1e-6)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
        res = (M.ata(Xs, R, lam=15.0) - Bm).reshape(F, R).norm(dim=0) / Bm.reshape(F, R).norm(dim=0)
        per_it = dt / max(it + 1, 1)
        ab = (NNZ * (4 + 8 * R) + 4 * (N + 1) + 8 * N * R) + (NNZ * (4 + 8 * R) + 4 * (F + 1) + 8 * F * R) + 11 * 8 * F * R
        emit(config="C5 block CG R=32 lambda=15 tol=1e-6 (bsbm_cgn on CSR + cached transpose)", iterations=it, seconds=dt, ms_per_iteration=per_it * 1e3,
             nnz_rhs_per_s=2 * NNZ * R / per_it, alg_bytes=ab, alg_gbs=ab / per_it / 1e9, max_rel_residual=float(res.max()))
        ms = timed(lambda: M.ata(Xs, R, lam=15.0), 5)
        emit(config="C5 operator (lambda I + A'A) X, R=32, two gather passes", ms=ms, nnz_rhs_per_s=2 * NNZ * R / ms * 1e3)
    if out:
        out.close()


if __name__ == "__main__":
    main()
