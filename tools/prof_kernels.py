"""Launch every kernel family of the library a fixed, small number of times at the BASELINE.json
sizes, for ncu (one plain run first, then the same command under ncu).

    python tools/prof_kernels.py [--only spmv,formats,ata,cg,small_r] [--small]

Each family launches its kernels twice (one warm, one to read); nothing here times anything --
timing lives in bench.py / tools/bench_all.py.  Kernel names to filter on:
    csr_spmm_staged_kernel csr_spmm_kernel csr_stream_kernel stream_fixup_kernel csr_ata_fused_kernel
    blocked_spmm_kernel cbcsr_spmm_kernel gram_* cg_* small_solve_kernel
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import libfastsparse_b200 as fs  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="")
    ap.add_argument("--small", action="store_true")
    ap.add_argument("--cg-iters", type=int, default=3)
    args = ap.parse_args()
    N, F, NNZ = (1_000_000, 100_000, 20_000_000) if args.small else (10_000_000, 1_000_000, 200_000_000)
    want = set(args.only.split(",")) if args.only else None
    on = lambda k: want is None or k in want
    L = fs.lib()
    twice = range(2)

    if on("spmv"):
        # C3: double / binary SpMV with both R = 1 kernels (team-per-row default, merge-path stream)
        A = fs.DeviceMatrix.synth(0x5EED0003, 0, NNZ, N, F, with_vals=True)
        x = (torch.sin(7.0 * torch.arange(F, device="cuda", dtype=torch.float64) + 0.3) / 10).contiguous()
        y = torch.empty(N, dtype=torch.float64, device="cuda")
        z = torch.empty(F, dtype=torch.float64, device="cuda")
        for _ in twice:
            A.spmm(x, 1, out=y)
        fs.check(L.fsb_tune_csr_algo(3, 0, 0))
        for _ in twice:
            A.spmm(x, 1, out=y)
        fs.check(L.fsb_tune_csr_algo(0, 0, 0))
        for _ in twice:
            A.spmm_t(y, 1, out=z)
        if on("ata"):
            for _ in twice:
                A.ata(x, 1, mode=1, out=z)
        del A
        torch.cuda.synchronize()

    if on("small_r"):
        B = fs.DeviceMatrix.synth(0x5EED0003, 0, NNZ, N, F)
        for R in (2, 4, 8):
            X = torch.randn(F * R, dtype=torch.float64, device="cuda")
            Y = torch.empty(N * R, dtype=torch.float64, device="cuda")
            for _ in twice:
                B.spmm(X, R, out=Y)
            del X, Y
        del B
        torch.cuda.synchronize()

    if on("formats"):
        # C4: power-law columns, R = 32, native blocked / column-blocked kernels
        R = 32
        M = fs.DeviceMatrix.synth(0x5EED0004, 1, NNZ, N, F, keep_coo=True)
        rows, cols, _ = M.coo
        X = torch.randn(F * R, dtype=torch.float64, device="cuda")
        Y = torch.empty(N * R, dtype=torch.float64, device="cuda")
        for _ in twice:
            M.spmm(X, R, out=Y)
        fs.check(L.fsb_tune_formats(1))
        Bk = fs.DeviceMatrix.blocked_from_coo_tensors(N, F, rows, cols, None, 512, order=1)
        for _ in twice:
            Bk.spmm(X, R, out=Y)
        del Bk
        Cb = fs.DeviceMatrix.cbcsr_from_coo_tensors(N, F, rows, cols, 65536)
        for _ in twice:
            Cb.spmm(X, R, out=Y)
        del Cb, M, X, Y, rows, cols
        fs.check(L.fsb_tune_formats(0))
        torch.cuda.synchronize()

    if on("cg"):
        # C5: a few block-CG iterations on the C2 matrix (every dense kernel + both products)
        R = 32
        M = fs.DeviceMatrix.synth(0x5EED0002, 0, NNZ, N, F)
        g = torch.Generator(device="cuda"); g.manual_seed(5)
        Bm = torch.randn(F * R, dtype=torch.float64, device="cuda", generator=g)
        Xs, it = M.cg(Bm, R, lam=15.0, tol=1e-30, max_iter=args.cg_iters)
        torch.cuda.synchronize()
        print("cg iterations run:", it)
    print("prof_kernels done; launches:", fs.launch_count())


if __name__ == "__main__":
    main()
