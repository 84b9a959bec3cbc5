"""Where does the x-blocked transpose start to pay?  Transposed SpMV (double CSR, 20 entries per row, 1 M columns) for
operands of 24 .. 80 MB, plain cached transpose against the x-blocked one with 16 / 24 / 32 MB blocks.
    python tools/xblock_threshold.py > gpurun_out/xblock_threshold.jsonl"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import libfastsparse_b200 as fs  # noqa: E402


def timed(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


L = fs.lib()
F = 1_000_000
for nrow in (3_000_000, 4_000_000, 5_000_000, 6_000_000, 8_000_000, 10_000_000):
    nnz = nrow * 20
    x = torch.randn(nrow, dtype=torch.float64, device="cuda")
    z = torch.empty(F, dtype=torch.float64, device="cuda")
    out = {"nrow": nrow, "x_MB": nrow * 8 / 1e6}
    fs.check(L.fsb_tune(b"t_xblock", 0))
    A = fs.DeviceMatrix.synth(0x5EED0003, 0, nnz, nrow, F, with_vals=True)
    out["plain_ms"] = timed(lambda: A.spmm_t(x, 1, out=z))
    ref = z.clone()
    A.free()
    fs.check(L.fsb_tune(b"t_xblock", 1)); fs.check(L.fsb_tune(b"t_xblock_min_kb", 1))
    for kb in (16 << 10, 24 << 10, 32 << 10, 48 << 10):
        if kb << 10 >= nrow * 8:
            continue
        fs.check(L.fsb_tune(b"t_xblock_kb", kb))
        A = fs.DeviceMatrix.synth(0x5EED0003, 0, nnz, nrow, F, with_vals=True)
        out[f"xblock_{kb >> 10}MB_ms"] = timed(lambda: A.spmm_t(x, 1, out=z))
        out[f"xblock_{kb >> 10}MB_relerr"] = float((z - ref).abs().max() / ref.abs().max())
        A.free()
    print(json.dumps(out), flush=True)
