"""C2 / C4 binary SpMM R = 32 on the staged kernel: rows per CTA (= shared-memory footprint, hence the carve-out the
driver picks and the L1 left for gathers in flight) x explicit carve-out.

    python tools/staged_rb_probe.py [--small] [--only c2,c4] [--out gpurun_out/staged_rb_probe.jsonl]
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
    ap.add_argument("--only", default="")
    ap.add_argument("--out", default="")
    ap.add_argument("--reps", type=int, default=8)
    args = ap.parse_args()
    N, F, NNZ = (1_000_000, 100_000, 20_000_000) if args.small else (10_000_000, 1_000_000, 200_000_000)
    want = set(args.only.split(",")) if args.only else None
    L = fs.lib()
    out = open(args.out, "w") if args.out else None
    R = 32
    for key, seed, dist, slabs, deep in (("c2", 0x5EED0002, 0, 2, 1), ("c4", 0x5EED0004, 1, 1, 1), ("c4", 0x5EED0004, 1, 1, 0)):
        if want and key not in want:
            continue
        M = fs.DeviceMatrix.synth(seed, dist, NNZ, N, F)
        X = torch.randn(F * R, dtype=torch.float64, device="cuda"); Y = torch.empty(N * R, dtype=torch.float64, device="cuda")
        fs.check(L.fsb_tune_csr_spmm(0, 16 // slabs, 2, slabs)); fs.check(L.fsb_tune_csr_staged(deep))
        ref = None
        for rb in (0, 32, 64, 128, 256):
            for co in (28, 44, 58):
                fs.check(L.fsb_tune_csr_algo(2, rb, 0)); fs.check(L.fsb_tune(b"staged_carveout", co))
                ms = timed(lambda: M.spmm(X, R, out=Y), args.reps)
                if ref is None:
                    ref = Y.clone()
                line = dict(workload=key, deep=deep, passes=slabs, rows_per_cta=rb, carveout_pct=co, ms=ms, maxdiff=float((Y - ref).abs().max()))
                print(json.dumps(line), flush=True)
                if out:
                    out.write(json.dumps(line) + "\n"); out.flush()
        del M, X, Y
    fs.check(L.fsb_tune_csr_algo(0, 0, 0)); fs.check(L.fsb_tune_csr_staged(-1)); fs.check(L.fsb_tune_csr_spmm(0, 0, 0, 0))


if __name__ == "__main__":
    main()
