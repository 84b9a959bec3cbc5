"""C3 double SpMV on the staged row-block kernel (values staged in shared memory, two lanes per row) under explicit
rows-per-CTA / staging capacity / carve-out / build settings, against the merge-path stream kernel.

    python tools/staged_r1_probe.py [--small] [--out gpurun_out/staged_r1_probe.jsonl]
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
    A = fs.DeviceMatrix.synth(0x5EED0003, 0, NNZ, N, F, with_vals=True)
    x = (torch.sin(7.0 * torch.arange(F, device="cuda", dtype=torch.float64) + 0.3) / 10).contiguous()
    y = torch.empty(N, dtype=torch.float64, device="cuda")

    def emit(**kw):
        print(json.dumps(kw), flush=True)
        if out:
            out.write(json.dumps(kw) + "\n"); out.flush()

    fs.check(L.fsb_tune(b"stream_tma_minb", 6))
    ms = timed(lambda: A.spmm(x, 1, out=y), args.reps)
    ref = y.clone()
    emit(kernel="stream (TMA-fed, 6 CTAs/SM)", ms=ms)
    for g in (2, 1, 4):
        fs.check(L.fsb_tune_csr_spmm(0, g, 1, 0))
        for rb in (32, 64, 128):
            for cap in (1, 0):
                for deep in (0, 1):
                    for co in (44, 58, 72):
                        fs.check(L.fsb_tune_csr_algo(2, rb, cap))
                        fs.check(L.fsb_tune_csr_staged(deep))
                        fs.check(L.fsb_tune(b"staged_carveout", co))
                        ms = timed(lambda: A.spmm(x, 1, out=y), args.reps)
                        emit(kernel="staged", lanes_per_row=g, rows_per_cta=rb, cap_mult=cap or 1.5, deep=deep, carveout_pct=co, ms=ms,
                             maxrel=float(((y - ref).abs() / (ref.abs() + 1)).max()))
    fs.check(L.fsb_tune_csr_algo(0, 0, 0)); fs.check(L.fsb_tune_csr_staged(-1)); fs.check(L.fsb_tune_csr_spmm(0, 0, 0, 0))


if __name__ == "__main__":
    main()
