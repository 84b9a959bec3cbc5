#!/bin/bash
# round 2, GPU session 5 (2 GPUs): multi-GPU correctness with the peer-store all-gather, CG phase trace with / without
# it, the new strong-scaling bench line with its collectives block, host-copy ceiling at 2 GPUs
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29511 tools/dist_check.py > gpurun_out/r2e_dist_check_n2.json 2> gpurun_out/r2e_dist_check_n2.err; echo "rc=$?" >> gpurun_out/r2e_dist_check_n2.err
FSB_CG_TRACE=2 timeout 600 $TR --master-port 29512 tools/bench_dist.py --only c5 > gpurun_out/r2e_cg_n2_p2p.jsonl 2> gpurun_out/r2e_cg_n2_p2p.trace; echo "rc=$?" >> gpurun_out/r2e_cg_n2_p2p.trace
FSB_TUNE_CG_P2P=0 FSB_CG_TRACE=2 timeout 600 $TR --master-port 29513 tools/bench_dist.py --only c5 > gpurun_out/r2e_cg_n2_nccl.jsonl 2> gpurun_out/r2e_cg_n2_nccl.trace; echo "rc=$?" >> gpurun_out/r2e_cg_n2_nccl.trace
timeout 900 $TR --master-port 29514 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2e_bench_n2.json 2> gpurun_out/r2e_bench_n2.err; echo "rc=$?" >> gpurun_out/r2e_bench_n2.err
timeout 300 $TR --master-port 29515 tools/d2h_probe.py > gpurun_out/r2e_d2h_probe_n2.json 2> gpurun_out/r2e_d2h_probe_n2.err
timeout 300 python tools/d2h_probe.py > gpurun_out/r2e_d2h_probe_n1.json 2>> gpurun_out/r2e_d2h_probe_n2.err
timeout 300 python tools/bench_all.py --only c3 --out gpurun_out/r2e_c3.jsonl > /dev/null 2> gpurun_out/r2e_c3.err
echo done
