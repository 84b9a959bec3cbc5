"""Sweep the CSR SpMM launch parameters on a bench workload (one GPU).

    python tools/sweep.py [--workload c2] [--reps 5] [--combos "algo,tw,g,vec,slabs,rb;..."]

algo 1 = team-per-row kernel (tw, g, vec, slabs), algo 2 = staged row-block kernel (g, vec,
slabs, rb = rows per CTA, 0 = automatic).  Prints one JSON line per setting: ms per product,
nnz*RHS/s, algorithmic GB/s, max |diff| against the first setting."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import libfastsparse_b200 as fs  # noqa: E402
from bench import WORKLOADS, alg_bytes  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="c2")
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--R", type=int, default=0)
    ap.add_argument("--vals", action="store_true")
    ap.add_argument("--dist", type=int, default=-1)
    ap.add_argument("--transpose", action="store_true", help="time the product with the cached transpose (A' X)")
    ap.add_argument("--combos", default="")
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    nrow, ncol, nnz, R, dkind, seed = WORKLOADS[args.workload]
    if args.R:
        R = args.R
    if args.dist >= 0:
        dkind = args.dist
    # box calibration: plain device copy bandwidth (read + write bytes), to compare runs on different boxes
    a = torch.empty(1 << 28, dtype=torch.float32, device="cuda"); b = torch.empty_like(a)
    for _ in range(3):
        b.copy_(a)
    torch.cuda.synchronize()
    c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    c0.record()
    for _ in range(10):
        b.copy_(a)
    c1.record(); torch.cuda.synchronize()
    print(json.dumps({"calibration_copy_gbs": 2 * a.numel() * 4 * 10 / (c0.elapsed_time(c1) * 1e-3) / 1e9}), flush=True)
    del a, b
    A = fs.DeviceMatrix.synth(seed, dkind, nnz, nrow, ncol, with_vals=args.vals)
    nin, nout = (nrow, ncol) if args.transpose else (ncol, nrow)
    X = torch.randn(nin * R, dtype=torch.float64, device="cuda")
    Y = torch.empty(nout * R, dtype=torch.float64, device="cuda")
    run = (lambda: A.spmm_t(X, R, out=Y)) if args.transpose else (lambda: A.spmm(X, R, out=Y))
    if args.combos:
        combos = [tuple(int(v) for v in c.split(",")) for c in args.combos.split(";")]   # optional 7th value: staged build (0 lean, 1 deep)
    else:
        combos = [(0, 0, 0, 0, 0, 0)]
        for vec in (4, 2, 1):
            if R % vec:
                continue
            g = 1
            while g * vec < R and g < 32:
                g *= 2
            for rb in (0, 64, 256):
                combos.append((2, 0, g, vec, 1, rb))
            for tw in (32, 16, 8):
                if tw >= g:
                    combos.append((1, tw, g, vec, 1, 0))
    ref = None
    rows = []
    for combo in combos:
        algo, tw, g, vec, slabs, rb = combo[:6]
        deep = combo[6] if len(combo) > 6 else -1
        xslabs = combo[7] if len(combo) > 7 else 0      # optional 8th value: contiguous (repacked) column slabs of the dense operand
        fs.check(fs.lib().fsb_tune(b"x_slabs", xslabs))
        fs.check(fs.lib().fsb_tune_csr_staged(deep))
        fs.check(fs.lib().fsb_tune_csr_algo(algo, rb, 0))
        fs.check(fs.lib().fsb_tune_csr_spmm(tw, g, vec, slabs))
        for _ in range(2):
            run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.reps):
            run()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.reps
        if ref is None:
            ref = Y.clone()
            err = 0.0
        else:
            err = float((Y - ref).abs().max())
        idx_bytes = 12 if args.vals else 4
        ab = nnz * (idx_bytes + 8 * R) + 4 * (nout + 1) + 8 * nout * R if 8 * nin * R > 126e6 else \
            nnz * idx_bytes + 4 * (nout + 1) + 8 * nout * R + 8 * nin * R
        row = dict(algo=algo, tw=tw, g=g, vec=vec, slabs=slabs, rb=rb, deep=deep, x_slabs=xslabs, ms=ms, nnz_rhs_per_s=nnz * R / ms * 1e3, alg_gbs=ab / ms / 1e6, maxdiff=err)
        rows.append(row)
        print(json.dumps(row), flush=True)
    if args.out:
        json.dump(rows, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
