#!/bin/bash
# session 21 (N GPUs, N = $1): strong-scaling C3 / C5 on row shards + the weak-scaling bench line
N=${1:-8}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo21_n$N.txt 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29545 tools/bench_dist.py > gpurun_out/bench_dist21_n$N.jsonl 2> gpurun_out/bench_dist21_n$N.err; echo "rc=$?" >> gpurun_out/bench_dist21_n$N.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29546 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench21_n$N.json 2> gpurun_out/bench21_n$N.err; echo "rc=$?" >> gpurun_out/bench21_n$N.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29547 tools/dist_check.py > gpurun_out/dist_check21_n$N.json 2> gpurun_out/dist_check21_n$N.err; echo "rc=$?" >> gpurun_out/dist_check21_n$N.err
echo done
