#!/bin/bash
# round 2, GPU session 35 (4 GPUs): the code as shipped (TMA-fed stream kernel) at N = 4 -- dist_check and the bench line
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29511 tools/dist_check.py > gpurun_out/r2z_dist_check_n4.json 2> gpurun_out/r2z_dist_check_n4.err; echo "dist_check rc=$?"
timeout 400 $TR --master-port 29514 bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/r2z_bench_c2_n4.json 2> gpurun_out/r2z_bench_n4.err; echo "bench rc=$?"
tail -c 300 gpurun_out/r2z_dist_check_n4.json
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2z_bench_c2_n4.json").read().strip().splitlines()[-1])
print({k:d.get(k) for k in ("value","ms_per_step","n_gpus","scaling","shard_parity")}, d["e2e"].get("ms_per_step"))
c=d.get("collectives",{})
for k in ("c3_At_mul_B_allreduce","c5_operator","c5_block_cg"):
    v=c.get(k,{}); print(k, v.get("ms") or v.get("ms_per_iteration"), v.get("speedup"), v.get("parity"))
PY
