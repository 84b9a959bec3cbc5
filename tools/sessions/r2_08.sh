#!/bin/bash
# round 2, GPU session 8 (8 GPUs, 8x cost: essentials only): correctness with the peer-memory paths, block-CG iteration
# for the baseline and the best combination, phase trace, the bench line at N = 8, host-copy ceiling
mkdir -p gpurun_out
N=${NGPU:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
ON="FSB_TUNE_CG_P2P=1 FSB_TUNE_CG_P2P_RS=${RS:-1} FSB_TUNE_CG_GRAPH=1 FSB_TUNE_HOST_X_ALLGATHER=1"
nvidia-smi topo -m > gpurun_out/r2h_topo_n${N}.txt 2>&1
env $ON timeout 400 $TR --master-port 29511 tools/dist_check.py > gpurun_out/r2h_dist_check_n${N}_on.json 2> gpurun_out/r2h_dist_check_n${N}_on.err; echo "rc=$?" >> gpurun_out/r2h_dist_check_n${N}_on.err
for combo in "0 0 0" "1 0 1" "1 1 1"; do set -- $combo
  FSB_TUNE_CG_P2P=$1 FSB_TUNE_CG_P2P_RS=$2 FSB_TUNE_CG_GRAPH=$3 timeout 400 $TR --master-port 29512 tools/bench_dist.py --only c5 > gpurun_out/r2h_cg_n${N}_p2p$1_rs$2_graph$3.jsonl 2> gpurun_out/r2h_cg_n${N}_p2p$1_rs$2_graph$3.err
done
FSB_TUNE_CG_P2P=1 FSB_TUNE_CG_P2P_RS=1 FSB_CG_TRACE=2 timeout 400 $TR --master-port 29513 tools/bench_dist.py --only c5 > /dev/null 2> gpurun_out/r2h_cg_n${N}_p2p1_rs1.trace
env $ON timeout 600 $TR --master-port 29514 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2h_bench_n${N}_on.json 2> gpurun_out/r2h_bench_n${N}_on.err; echo "rc=$?" >> gpurun_out/r2h_bench_n${N}_on.err
timeout 200 $TR --master-port 29515 tools/d2h_probe.py > gpurun_out/r2h_d2h_probe_n${N}.json 2> gpurun_out/r2h_d2h_probe_n${N}.err
echo done
