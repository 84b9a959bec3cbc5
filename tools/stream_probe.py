"""C3 double SpMV / transposed SpMV with the merge-path stream kernel under a list of knob settings (one process, one
matrix): which form of the kernel (per-thread loads or TMA-fed), how many resident CTAs per SM (= how much of the
256 KB is left to L1 for outstanding gather misses), gathers by LDG or through the texture pipe.  (The run recorded in
profiles/r2w_stream_probe.jsonl also had a gathers-with-L1::no_allocate setting, removed from the kernel since.)

    python tools/stream_probe.py [--small] [--out gpurun_out/stream_probe.jsonl]
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

SETTINGS = [
    {"stream_tma": 0},
    {"stream_tma": 1, "stream_tma_minb": 8, "stream_tex": 0},
    {"stream_tma": 1, "stream_tma_minb": 6, "stream_tex": 0},
    {"stream_tma": 1, "stream_tma_minb": 4, "stream_tex": 0},
    {"stream_tma": 1, "stream_tma_minb": 8, "stream_tex": 1},
    {"stream_tma": 1, "stream_tma_minb": 6, "stream_tex": 1},
    {"stream_tma": 1, "stream_tma_minb": 4, "stream_tex": 1},
]
ALL = ("stream_tma", "stream_tma_minb", "stream_tex")
DEFAULTS = {"stream_tma": 1, "stream_tma_minb": 6, "stream_tex": 1}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--small", action="store_true")
    ap.add_argument("--out", default="")
    ap.add_argument("--reps", type=int, default=10)
    args = ap.parse_args()
    N, F, NNZ = (1_000_000, 100_000, 20_000_000) if args.small else (10_000_000, 1_000_000, 200_000_000)
    L = fs.lib()
    out = open(args.out, "w") if args.out else None
    for with_vals in (True, False):
        A = fs.DeviceMatrix.synth(0x5EED0003, 0, NNZ, N, F, with_vals=with_vals)
        x = (torch.sin(7.0 * torch.arange(F, device="cuda", dtype=torch.float64) + 0.3) / 10).contiguous()
        y = torch.empty(N, dtype=torch.float64, device="cuda")
        z = torch.empty(F, dtype=torch.float64, device="cuda")
        fs.check(L.fsb_tune_csr_algo(3, 0, 0))      # the stream kernel for binary matrices too
        yref = zref = None
        for s in SETTINGS:
            for k in ALL:
                fs.check(L.fsb_tune(k.encode(), s.get(k, DEFAULTS[k])))
            ms = timed(lambda: A.spmm(x, 1, out=y), args.reps)
            mt = timed(lambda: A.spmm_t(y, 1, out=z), args.reps) if with_vals else None
            if yref is None:
                yref, zref = y.clone(), z.clone()
            line = dict(matrix="C3 double" if with_vals else "C3 binary", knobs=s, spmv_ms=ms, spmv_t_ms=mt,
                        maxdiff=float((y - yref).abs().max()), maxdiff_t=float((z - zref).abs().max()) if with_vals else None)
            print(json.dumps(line), flush=True)
            if out:
                out.write(json.dumps(line) + "\n"); out.flush()
        fs.check(L.fsb_tune_csr_algo(0, 0, 0))
        for k in ALL:
            fs.check(L.fsb_tune(k.encode(), DEFAULTS[k]))
        del A


if __name__ == "__main__":
    main()
