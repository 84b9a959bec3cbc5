#!/bin/bash
# round 2, GPU session 20 (8 GPUs): the final code with its defaults at N = 8 -- correctness (dist_check) and the bench line
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 400 $TR --master-port 29511 tools/dist_check.py > gpurun_out/r2t_dist_check_n8.json 2> gpurun_out/r2t_dist_check_n8.err; echo "rc=$?" >> gpurun_out/r2t_dist_check_n8.err
timeout 600 $TR --master-port 29514 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2t_bench_n8.json 2> gpurun_out/r2t_bench_n8.err; echo "rc=$?" >> gpurun_out/r2t_bench_n8.err
FSB_CG_TRACE=2 timeout 300 $TR --master-port 29513 tools/bench_dist.py --only c5 > gpurun_out/r2t_bench_dist_n8.jsonl 2> gpurun_out/r2t_cg_n8.trace
echo done
