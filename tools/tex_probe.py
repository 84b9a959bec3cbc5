"""Gathers through the texture pipe (knob staged_tex) against LDG on every product that runs the staged kernel:
binary SpMV, narrow SpMM (R = 2, 4, 8, 16), C4 and C2 at R = 32 (one and two column passes, lean and deep builds).

    python tools/tex_probe.py [--small] [--out gpurun_out/tex_probe.jsonl]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import libfastsparse_b200 as fs  # noqa: E402
from tools.bench_all import timed  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--small", action="store_true")
    ap.add_argument("--out", default="")
    ap.add_argument("--reps", type=int, default=10)
    args = ap.parse_args()
    N, F, NNZ = (1_000_000, 100_000, 20_000_000) if args.small else (10_000_000, 1_000_000, 200_000_000)
    L = fs.lib()
    out = open(args.out, "w") if args.out else None

    def emit(**kw):
        print(json.dumps(kw), flush=True)
        if out:
            out.write(json.dumps(kw) + "\n"); out.flush()

    def ab(name, fn, y):
        ref = None
        for rep in range(2):
            for tex in (0, 1):
                fs.check(L.fsb_tune(b"staged_tex", tex))
                ms = timed(fn, args.reps)
                if ref is None:
                    ref = y.clone()
                emit(product=name, gathers="texture" if tex else "LDG", rep=rep, ms=ms, maxdiff=float((y - ref).abs().max()))
        fs.check(L.fsb_tune(b"staged_tex", -1))

    x = (torch.sin(7.0 * torch.arange(F, device="cuda", dtype=torch.float64) + 0.3) / 10).contiguous()
    for with_vals in (False, True):
        B = fs.DeviceMatrix.synth(0x5EED0003, 0, NNZ, N, F, with_vals=with_vals)
        kind = "double" if with_vals else "binary"
        if not with_vals:
            y = torch.empty(N, dtype=torch.float64, device="cuda")
            ab("C3 binary SpMV (staged, 2 lanes per row)", lambda: B.spmm(x, 1, out=y), y)
        for R in (2, 4, 8, 16):
            X = torch.randn(F * R, dtype=torch.float64, device="cuda"); Y = torch.empty(N * R, dtype=torch.float64, device="cuda")
            ab(f"C3 {kind} SpMM R={R} (staged)", lambda: B.spmm(X, R, out=Y), Y)
            del X, Y
        del B
    R = 32
    for key, seed, dist in (("C4 power-law columns", 0x5EED0004, 1), ("C2 uniform columns", 0x5EED0002, 0)):
        M = fs.DeviceMatrix.synth(seed, dist, NNZ, N, F)
        X = torch.randn(F * R, dtype=torch.float64, device="cuda"); Y = torch.empty(N * R, dtype=torch.float64, device="cuda")
        for slabs in (1, 2):
            for deep in (0, 1):
                fs.check(L.fsb_tune_csr_algo(2, 0, 0)); fs.check(L.fsb_tune_csr_spmm(0, 16 // slabs, 2, slabs)); fs.check(L.fsb_tune_csr_staged(deep))
                ab(f"{key}, binary SpMM R=32, {slabs} pass(es), {'deep' if deep else 'lean'} build", lambda: M.spmm(X, R, out=Y), Y)
        fs.check(L.fsb_tune_csr_algo(0, 0, 0)); fs.check(L.fsb_tune_csr_spmm(0, 0, 0, 0)); fs.check(L.fsb_tune_csr_staged(-1))
        del M, X, Y


if __name__ == "__main__":
    main()
