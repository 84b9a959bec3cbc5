#!/bin/bash
# round 2, GPU session 17 (2 GPUs): dist_check with the host-pointer sharded product, multi-GPU pytest, bench N = 2 on the final code
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29511 tools/dist_check.py > gpurun_out/r2q_dist_check_n2.json 2> gpurun_out/r2q_dist_check_n2.err; echo "rc=$?" >> gpurun_out/r2q_dist_check_n2.err
timeout 900 $TR --master-port 29514 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2q_bench_n2.json 2> gpurun_out/r2q_bench_n2.err; echo "rc=$?" >> gpurun_out/r2q_bench_n2.err
echo done
