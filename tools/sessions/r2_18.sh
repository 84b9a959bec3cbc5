#!/bin/bash
# round 2, GPU session 18 (2 GPUs): chunked / overlapped allreduce of the sharded A'(A X) operator -- correctness (dist_check)
# and timing (bench collectives block, tools/bench_dist.py)
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29511 tools/dist_check.py > gpurun_out/r2r_dist_check_n2.json 2> gpurun_out/r2r_dist_check_n2.err; echo "rc=$?" >> gpurun_out/r2r_dist_check_n2.err
timeout 600 $TR --master-port 29512 tools/bench_dist.py --only c5 > gpurun_out/r2r_bench_dist_n2_overlap.jsonl 2> /dev/null
FSB_TUNE_ATA_OVERLAP=0 timeout 600 $TR --master-port 29513 tools/bench_dist.py --only c5 > gpurun_out/r2r_bench_dist_n2_plain.jsonl 2> /dev/null
echo done
