#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest6.log 2>&1; echo "rc=$?" >> gpurun_out/pytest6.log
# C2: L2 policy on/off (cap_mult hundreds digit 1 = no hints), column slabs
timeout 600 python tools/sweep.py --workload c2 --reps 8 --combos "2,0,16,2,1,0;2,0,16,2,1,0;2,0,8,2,2,0;2,0,16,1,2,0;2,0,4,2,4,0;2,0,8,1,4,0;2,0,32,1,1,0;2,0,8,4,1,0" > gpurun_out/sweep6_c2_l2on.log 2>&1
python - <<'PY' > gpurun_out/sweep6_c2_l2off.log 2>&1
import sys, os, json
sys.path.insert(0, os.getcwd())
import torch, libfastsparse_b200 as fs
from bench import WORKLOADS, alg_bytes
nrow, ncol, nnz, R, dk, seed = WORKLOADS["c2"]
A = fs.DeviceMatrix.synth(seed, dk, nnz, nrow, ncol)
X = torch.randn(ncol*R, dtype=torch.float64, device="cuda"); Y = torch.empty(nrow*R, dtype=torch.float64, device="cuda")
for name, cm in [("hints_on", 0), ("hints_off", 100), ("hints_on", 0), ("hints_off", 100)]:
    for g, vec, slabs in [(16,2,1),(8,2,2),(32,1,1)]:
        fs.check(fs.lib().fsb_tune_csr_algo(2, 0, cm)); fs.check(fs.lib().fsb_tune_csr_spmm(0, g, vec, slabs))
        for _ in range(3): A.spmm(X, R, out=Y)
        torch.cuda.synchronize(); e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(8): A.spmm(X, R, out=Y)
        e1.record(); torch.cuda.synchronize(); ms=e0.elapsed_time(e1)/8
        print(json.dumps(dict(l2=name, g=g, vec=vec, slabs=slabs, ms=ms, alg_gbs=alg_bytes(nrow,nnz,R)/ms/1e6)), flush=True)
PY
timeout 1500 python tools/bench_all.py --out gpurun_out/bench_all6.jsonl > gpurun_out/bench_all6.log 2>&1; echo "rc=$?" >> gpurun_out/bench_all6.log
echo done
