"""The Macau-style caller loop with the matrix resident in HBM (bench_a_mul_b.c:331-360):
per sample, draw noise, form B = A'N + sqrt(lambda) E and solve (A'A + lambda I) X = B by block CG.
Prints one JSON line: solves per second, mean iterations, per-phase milliseconds.

    python tools/macau_loop.py [--small] [--samples 5] [--R 32]
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import libfastsparse_b200 as fs  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--small", action="store_true")
    ap.add_argument("--samples", type=int, default=5)
    ap.add_argument("--R", type=int, default=32)
    ap.add_argument("--lam", type=float, default=15.0)
    ap.add_argument("--tol", type=float, default=1e-6)
    args = ap.parse_args()
    N, F, NNZ = (1_000_000, 100_000, 20_000_000) if args.small else (10_000_000, 1_000_000, 200_000_000)
    R = args.R
    t0 = time.perf_counter()
    A = fs.DeviceMatrix.synth(0x5EED0002, 0, NNZ, N, F)
    B = torch.empty(F * R, dtype=torch.float64, device="cuda")
    X = torch.empty(F * R, dtype=torch.float64, device="cuda")
    A.noise_rhs(R, args.lam, 0, out=B); A.cg(B, R, lam=args.lam, tol=args.tol, out=X)     # warm-up: builds the cached transpose, autotunes
    torch.cuda.synchronize()
    setup = time.perf_counter() - t0
    rhs_ms, cg_ms, its = [], [], []
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    for s in range(1, args.samples + 1):
        e[0].record()
        A.noise_rhs(R, args.lam, s, out=B)
        e[1].record()
        _, it = A.cg(B, R, lam=args.lam, tol=args.tol, out=X)
        e[2].record()
        torch.cuda.synchronize()
        rhs_ms.append(e[0].elapsed_time(e[1])); cg_ms.append(e[1].elapsed_time(e[2])); its.append(it)
    res = (A.ata(X, R, lam=args.lam) - B).reshape(F, R).norm(dim=0) / B.reshape(F, R).norm(dim=0)
    per = (sum(rhs_ms) + sum(cg_ms)) / args.samples
    print(json.dumps(dict(config=f"Macau-style loop: binary CSR {N}x{F}, {NNZ} nnz resident, R={R}, lambda={args.lam}, tol={args.tol}",
                          samples=args.samples, solves_per_s=1e3 / per, ms_per_sample=per, rhs_ms=sum(rhs_ms) / args.samples,
                          cg_ms=sum(cg_ms) / args.samples, iterations=its, nnz_rhs_per_s=(2 * NNZ * R * (sum(its) + len(its))) / (sum(cg_ms) * 1e-3),
                          max_rel_residual_last=float(res.max()), setup_s=setup)), flush=True)


if __name__ == "__main__":
    main()
