"""How much of an SM's 256 KB should stay L1?  A gather that misses L1 holds a line there until its data returns, so the
number of gathers in flight -- and with it the gather rate -- depends on the shared-memory carve-out the driver picks
from the kernel's footprint x resident CTAs.  Times the gather-bound products under explicit carve-outs
(knobs staged_carveout / stream_carveout, percent of the 228 KB maximum; -1 = the driver's choice).

    python tools/carveout_probe.py [--small] [--only c3d,c3b,r4,r8,c4,c2] [--out gpurun_out/carveout_probe.jsonl]
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

CARVE = (-1, 14, 28, 44, 58, 72, 86, 100)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--small", action="store_true")
    ap.add_argument("--only", default="")
    ap.add_argument("--out", default="")
    ap.add_argument("--reps", type=int, default=10)
    args = ap.parse_args()
    N, F, NNZ = (1_000_000, 100_000, 20_000_000) if args.small else (10_000_000, 1_000_000, 200_000_000)
    want = set(args.only.split(",")) if args.only else None
    on = lambda k: want is None or k in want
    L = fs.lib()
    out = open(args.out, "w") if args.out else None

    def sweep(name, knob, fn, check):
        ref = None
        for co in CARVE:
            fs.check(L.fsb_tune(knob.encode(), co))
            ms = timed(fn, args.reps)
            got = check()
            if ref is None:
                ref = got.clone()
            line = dict(product=name, knob=knob, carveout_pct=co, ms=ms, maxdiff=float((got - ref).abs().max()))
            print(json.dumps(line), flush=True)
            if out:
                out.write(json.dumps(line) + "\n"); out.flush()
        fs.check(L.fsb_tune(knob.encode(), -1))

    x = (torch.sin(7.0 * torch.arange(F, device="cuda", dtype=torch.float64) + 0.3) / 10).contiguous()
    if on("c3d"):
        A = fs.DeviceMatrix.synth(0x5EED0003, 0, NNZ, N, F, with_vals=True)
        y = torch.empty(N, dtype=torch.float64, device="cuda"); z = torch.empty(F, dtype=torch.float64, device="cuda")
        for minb in (6, 4):
            fs.check(L.fsb_tune(b"stream_tma_minb", minb))
            sweep(f"C3 double SpMV, TMA-fed stream kernel built for {minb} CTAs/SM", "stream_carveout", lambda: A.spmm(x, 1, out=y), lambda: y)
            sweep(f"C3 double At_mul_B (x-blocked), TMA-fed stream kernel built for {minb} CTAs/SM", "stream_carveout", lambda: A.spmm_t(y, 1, out=z), lambda: z)
        del A
    B = None
    if on("c3b") or on("r4") or on("r8"):
        B = fs.DeviceMatrix.synth(0x5EED0003, 0, NNZ, N, F)
    if on("c3b"):
        y = torch.empty(N, dtype=torch.float64, device="cuda")
        sweep("C3 binary SpMV, staged kernel (2 lanes per row)", "staged_carveout", lambda: B.spmm(x, 1, out=y), lambda: y)
    for R, key in ((4, "r4"), (8, "r8")):
        if on(key):
            X = torch.randn(F * R, dtype=torch.float64, device="cuda"); Y = torch.empty(N * R, dtype=torch.float64, device="cuda")
            sweep(f"binary SpMM R={R}, staged kernel", "staged_carveout", lambda: B.spmm(X, R, out=Y), lambda: Y)
            del X, Y
    del B
    for key, seed, dist, label in (("c4", 0x5EED0004, 1, "C4 power-law columns"), ("c2", 0x5EED0002, 0, "C2 uniform columns")):
        if on(key):
            R = 32
            M = fs.DeviceMatrix.synth(seed, dist, NNZ, N, F)
            X = torch.randn(F * R, dtype=torch.float64, device="cuda"); Y = torch.empty(N * R, dtype=torch.float64, device="cuda")
            M.spmm(X, R, out=Y)   # autotune (pass count, build) under the driver's carve-out
            sweep(f"{label}, binary SpMM R=32, staged kernel ({json.dumps(M.tuning()) if hasattr(M, 'tuning') else ''})", "staged_carveout",
                  lambda: M.spmm(X, R, out=Y), lambda: Y)
            del M, X, Y


if __name__ == "__main__":
    main()
