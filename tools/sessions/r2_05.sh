#!/bin/bash
# round 2, GPU session 5 (2 GPUs): multi-GPU correctness with / without the peer-memory paths, CG timing for each
# combination, the strong-scaling bench line with its collectives block, host-copy ceiling at 2 GPUs
mkdir -p gpurun_out
N=${NGPU:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
ON="FSB_TUNE_CG_P2P=1 FSB_TUNE_CG_P2P_RS=1 FSB_TUNE_CG_GRAPH=1 FSB_TUNE_HOST_X_ALLGATHER=1"
env $ON FSB_CG_TRACE=1 timeout 600 $TR --master-port 29511 tools/dist_check.py > gpurun_out/r2e_dist_check_n${N}_on.json 2> gpurun_out/r2e_dist_check_n${N}_on.err; echo "rc=$?" >> gpurun_out/r2e_dist_check_n${N}_on.err
timeout 600 $TR --master-port 29516 tools/dist_check.py > gpurun_out/r2e_dist_check_n${N}_off.json 2> gpurun_out/r2e_dist_check_n${N}_off.err; echo "rc=$?" >> gpurun_out/r2e_dist_check_n${N}_off.err
# CG: plain timing for every combination, then the device phase trace (trace level 2 disables the graph)
for combo in "0 0 0" "1 0 0" "1 1 0" "0 0 1" "1 0 1" "1 1 1"; do set -- $combo     # peer all-gather + Gram, peer reduce-scatter, graph
  FSB_TUNE_CG_P2P=$1 FSB_TUNE_CG_P2P_RS=$2 FSB_TUNE_CG_GRAPH=$3 timeout 600 $TR --master-port 29512 tools/bench_dist.py --only c5 > gpurun_out/r2e_cg_n${N}_p2p$1_rs$2_graph$3.jsonl 2> gpurun_out/r2e_cg_n${N}_p2p$1_rs$2_graph$3.err
done
for combo in "0 0" "1 0" "1 1"; do set -- $combo
  FSB_TUNE_CG_P2P=$1 FSB_TUNE_CG_P2P_RS=$2 FSB_CG_TRACE=2 timeout 600 $TR --master-port 29513 tools/bench_dist.py --only c5 > /dev/null 2> gpurun_out/r2e_cg_n${N}_p2p$1_rs$2.trace
done
env $ON timeout 900 $TR --master-port 29514 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2e_bench_n${N}_on.json 2> gpurun_out/r2e_bench_n${N}_on.err; echo "rc=$?" >> gpurun_out/r2e_bench_n${N}_on.err
timeout 900 $TR --master-port 29517 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2e_bench_n${N}_off.json 2> gpurun_out/r2e_bench_n${N}_off.err; echo "rc=$?" >> gpurun_out/r2e_bench_n${N}_off.err
timeout 300 $TR --master-port 29515 tools/d2h_probe.py > gpurun_out/r2e_d2h_probe_n${N}.json 2> gpurun_out/r2e_d2h_probe_n${N}.err
echo done
