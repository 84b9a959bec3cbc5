#!/bin/bash
# round 2, GPU session 12 (2 GPUs): the unrolled peer push kernel -- correctness (dist_check) and all-gather time (phase trace)
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29511 tools/dist_check.py > gpurun_out/r2l_dist_check_n2.json 2> gpurun_out/r2l_dist_check_n2.err; echo "rc=$?" >> gpurun_out/r2l_dist_check_n2.err
FSB_CG_TRACE=2 timeout 600 $TR --master-port 29513 tools/bench_dist.py --only c5 > /dev/null 2> gpurun_out/r2l_cg_n2.trace
timeout 600 $TR --master-port 29512 tools/bench_dist.py > gpurun_out/r2l_bench_dist_n2.jsonl 2> gpurun_out/r2l_bench_dist_n2.err
echo done
