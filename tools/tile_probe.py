"""C3 double SpMV / transposed SpMV on the TMA-fed stream kernel: gathers through LDG against gathers through a linear
texture (knob stream_tex), per build (stream_tma_minb); FSB_LIB selects a tile-size variant build.

    python tools/tile_probe.py
"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import libfastsparse_b200 as fs
from tools.bench_all import timed
N, F, NNZ = 10_000_000, 1_000_000, 200_000_000
L = fs.lib()
for with_vals in (True, False):
    A = fs.DeviceMatrix.synth(0x5EED0003, 0, NNZ, N, F, with_vals=with_vals)
    x = (torch.sin(7.0 * torch.arange(F, device="cuda", dtype=torch.float64) + 0.3) / 10).contiguous()
    y = torch.empty(N, dtype=torch.float64, device="cuda"); z = torch.empty(F, dtype=torch.float64, device="cuda")
    fs.check(L.fsb_tune_csr_algo(3, 0, 0))
    ref = None
    for tex in (0, 1):
        for minb in (4, 6, 8):
            fs.check(L.fsb_tune(b"stream_tma_minb", minb)); fs.check(L.fsb_tune(b"stream_tex", tex))
            ms = timed(lambda: A.spmm(x, 1, out=y), 10); mt = timed(lambda: A.spmm_t(y, 1, out=z), 10)
            if ref is None:
                ref = (y.clone(), z.clone())
            print(json.dumps(dict(lib=os.path.basename(os.environ.get("FSB_LIB", "default")), matrix="double" if with_vals else "binary", gathers="texture" if tex else "LDG",
                                  minb=minb, spmv_ms=ms, spmv_t_ms=mt, maxdiff=float((y - ref[0]).abs().max()), maxdiff_t=float((z - ref[1]).abs().max()))), flush=True)
    fs.check(L.fsb_tune_csr_algo(0, 0, 0))
    del A
