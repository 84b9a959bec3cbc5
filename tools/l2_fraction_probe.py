"""C2 / C5-operator products with a FRACTIONAL L2 evict_last policy on the dense operand (keep a hashed subset of the
128 MB slab resident, stream the rest) against the shipped evict_last 1.0 and no hints.

    python tools/l2_fraction_probe.py [--small] [--out gpurun_out/l2_fraction_probe.jsonl]
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

KINDS = {0: "evict_last 1.0 (shipped)", 1: "no hints", 3: "evict_last 0.5 / evict_first", 4: "evict_last 0.75 / evict_first",
         5: "evict_last 0.375 / evict_first", 6: "evict_last 0.5 / normal"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--small", action="store_true")
    ap.add_argument("--out", default="")
    ap.add_argument("--reps", type=int, default=8)
    args = ap.parse_args()
    N, F, NNZ = (1_000_000, 100_000, 20_000_000) if args.small else (10_000_000, 1_000_000, 200_000_000)
    L = fs.lib()
    out = open(args.out, "w") if args.out else None
    R = 32
    M = fs.DeviceMatrix.synth(0x5EED0002, 0, NNZ, N, F)
    X = torch.randn(F * R, dtype=torch.float64, device="cuda"); Y = torch.empty(N * R, dtype=torch.float64, device="cuda")
    M.spmm(X, R, out=Y)
    ref = Y.clone()
    for slabs in (2, 1):
        fs.check(L.fsb_tune_csr_spmm(0, 16 // slabs, 2, slabs)); fs.check(L.fsb_tune_csr_staged(1))
        for rep in range(2):
            for kind, label in KINDS.items():
                fs.check(L.fsb_tune_csr_algo(2, 0, kind * 100))
                ms = timed(lambda: M.spmm(X, R, out=Y), args.reps)
                line = dict(workload="C2 binary SpMM R=32", passes=slabs, x_policy=label, rep=rep, ms=ms, maxdiff=float((Y - ref).abs().max()))
                print(json.dumps(line), flush=True)
                if out:
                    out.write(json.dumps(line) + "\n"); out.flush()
    fs.check(L.fsb_tune_csr_algo(0, 0, 0)); fs.check(L.fsb_tune_csr_staged(-1)); fs.check(L.fsb_tune_csr_spmm(0, 0, 0, 0))


if __name__ == "__main__":
    main()
