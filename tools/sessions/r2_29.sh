#!/bin/bash
# round 2, GPU session 29 (2 GPUs): the code as shipped (TMA-fed stream kernel) through the multi-GPU paths -- dist_check and the bench line
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 400 $TR --master-port 29511 tools/dist_check.py > gpurun_out/r2z_dist_check_n2.json 2> gpurun_out/r2z_dist_check_n2.err; echo "dist_check rc=$?"
timeout 600 $TR --master-port 29514 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2z_bench_c2_n2.json 2> gpurun_out/r2z_bench_n2.err; echo "bench rc=$?"
tail -c 600 gpurun_out/r2z_dist_check_n2.json
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2z_bench_c2_n2.json").read().strip().splitlines()[-1])
print({k:d.get(k) for k in ("value","ms_per_step","n_gpus","scaling","shard_parity")})
print(json.dumps(d.get("collectives"))[:1500])
PY
