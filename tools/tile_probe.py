import json, os, sys
sys.path.insert(0, "/root/repo")
import torch
import libfastsparse_b200 as fs
from tools.bench_all import timed
N, F, NNZ = 10_000_000, 1_000_000, 200_000_000
L = fs.lib()
A = fs.DeviceMatrix.synth(0x5EED0003, 0, NNZ, N, F, with_vals=True)
x = (torch.sin(7.0 * torch.arange(F, device="cuda", dtype=torch.float64) + 0.3) / 10).contiguous()
y = torch.empty(N, dtype=torch.float64, device="cuda"); z = torch.empty(F, dtype=torch.float64, device="cuda")
for minb in (4, 6, 8):
    fs.check(L.fsb_tune(b"stream_tma_minb", minb))
    ms = timed(lambda: A.spmm(x, 1, out=y), 10); mt = timed(lambda: A.spmm_t(y, 1, out=z), 10)
    print(json.dumps(dict(lib=os.path.basename(os.environ.get("FSB_LIB", "default")), minb=minb, spmv_ms=ms, spmv_t_ms=mt, chk=float(y.sum()))), flush=True)
