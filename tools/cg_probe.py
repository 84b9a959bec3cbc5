"""Where does a block-CG solve spend its wall time?  Runs the C5 solve with max_iter = 1, 2, 4, 8
(slope = per-iteration cost, intercept = per-solve overhead) and one full solve.  With FSB_CG_TRACE=1
the library also prints its own per-phase wall times on stderr."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import libfastsparse_b200 as fs  # noqa: E402

small = "--small" in sys.argv
N, F, NNZ = (1_000_000, 100_000, 20_000_000) if small else (10_000_000, 1_000_000, 200_000_000)
R = 32
A = fs.DeviceMatrix.synth(0x5EED0002, 0, NNZ, N, F)
B = A.noise_rhs(R, 15.0, 1)
X = torch.empty(F * R, dtype=torch.float64, device="cuda")
A.cg(B, R, lam=15.0, tol=1e-30, max_iter=2, out=X)   # warm-up: transpose, autotune
torch.cuda.synchronize()
for mi in (1, 2, 4, 8):
    t0 = time.perf_counter()
    _, it = A.cg(B, R, lam=15.0, tol=1e-30, max_iter=mi, out=X)
    torch.cuda.synchronize()
    print(json.dumps({"max_iter": mi, "iterations": it, "ms": (time.perf_counter() - t0) * 1e3}), flush=True)
t0 = time.perf_counter()
_, it = A.cg(B, R, lam=15.0, tol=1e-6, out=X)
torch.cuda.synchronize()
print(json.dumps({"full_solve_iterations": it, "ms": (time.perf_counter() - t0) * 1e3}), flush=True)
