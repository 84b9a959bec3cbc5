"""Sweep the CSR SpMM launch parameters on the C2 workload (one GPU).

    python tools/sweep.py [--workload c2] [--reps 5]

Prints one line per (TW, G, VEC, slabs): ms per product, nnz*RHS/s, algorithmic GB/s.
Used to pick the heuristic in kernels_csr.cu; results are copied into profiles/."""
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
    ap.add_argument("--combos", default="")
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    nrow, ncol, nnz, R, dkind, seed = WORKLOADS[args.workload]
    if args.R:
        R = args.R
    A = fs.DeviceMatrix.synth(seed, dkind, nnz, nrow, ncol)
    X = torch.randn(ncol * R, dtype=torch.float64, device="cuda")
    Y = torch.empty(nrow * R, dtype=torch.float64, device="cuda")
    ref = None
    if args.combos:
        combos = [tuple(int(v) for v in c.split(",")) for c in args.combos.split(";")]
    else:
        combos = [(0, 0, 0, 0)]
        for vec in (4, 2, 1):
            g = max(1, R // vec)
            if g > 32:
                continue
            for tw in (32, 16, 8):
                if tw >= g:
                    combos.append((tw, g, vec, 1))
        for slabs in (2, 4):
            for vec in (4, 2):
                g = max(1, R // slabs // vec)
                for tw in (32, 16):
                    if tw >= g:
                        combos.append((tw, g, vec, slabs))
    rows = []
    for tw, g, vec, slabs in combos:
        fs.check(fs.lib().fsb_tune_csr_spmm(tw, g, vec, slabs))
        for _ in range(2):
            A.spmm(X, R, out=Y)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.reps):
            A.spmm(X, R, out=Y)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.reps
        if ref is None:
            ref = Y.clone()
            err = 0.0
        else:
            err = float((Y - ref).abs().max())
        row = dict(tw=tw, g=g, vec=vec, slabs=slabs, ms=ms, nnz_rhs_per_s=nnz * R / ms * 1e3, alg_gbs=alg_bytes(nrow, nnz, R) / ms / 1e6, maxdiff=err)
        rows.append(row)
        print(json.dumps(row), flush=True)
    if args.out:
        json.dump(rows, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
