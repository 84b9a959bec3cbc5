"""Binary A'x (cached transpose: 200 entries per row at C3; x-blocked cells: ~67) -- warp-per-row kernel (the automatic
choice for binary long regular rows) against the TMA-fed merge-path stream kernel, with and without x-blocking.

    python tools/binary_t_probe.py [--small] [--out gpurun_out/binary_t_probe.jsonl]
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
    for dist, label in ((0, "C3 structure (uniform columns)"), (1, "C4 structure (power-law columns)")):
        B = fs.DeviceMatrix.synth(0x5EED0003 + dist, dist, NNZ, N, F)
        y = (torch.sin(3.0 * torch.arange(N, device="cuda", dtype=torch.float64) + 0.1)).contiguous()
        z = torch.empty(F, dtype=torch.float64, device="cuda")
        ref = None
        for xb in (1, 0):
            for algo in (0, 3):
                fs.check(L.fsb_tune(b"t_xblock", xb)); fs.check(L.fsb_tune_csr_algo(algo, 0, 0))
                ms = timed(lambda: B.spmm_t(y, 1, out=z), args.reps)
                if ref is None:
                    ref = z.clone()
                line = dict(matrix=label, product="binary A'x", x_blocked=bool(xb), kernel="stream (forced)" if algo == 3 else "automatic", ms=ms,
                            maxrel=float(((z - ref).abs() / (ref.abs() + 1)).max()))
                print(json.dumps(line), flush=True)
                if out:
                    out.write(json.dumps(line) + "\n"); out.flush()
        fs.check(L.fsb_tune(b"t_xblock", 1)); fs.check(L.fsb_tune_csr_algo(0, 0, 0))
        del B


if __name__ == "__main__":
    main()
