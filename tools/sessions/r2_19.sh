#!/bin/bash
# round 2, GPU session 19 (2 GPUs): dist_check after factoring the overlapped allreduce (A'(A X) and A' X)
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29511 tools/dist_check.py > gpurun_out/r2s_dist_check_n2.json 2> gpurun_out/r2s_dist_check_n2.err; echo "rc=$?" >> gpurun_out/r2s_dist_check_n2.err
timeout 900 $TR --master-port 29514 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2s_bench_n2.json 2> gpurun_out/r2s_bench_n2.err; echo "rc=$?" >> gpurun_out/r2s_bench_n2.err
echo done
