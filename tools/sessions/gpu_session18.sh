#!/bin/bash
# session 18 (2 GPUs): sharded multi-GPU CG (reduce-scatter overlapped with the A' product, all-gather of P) vs the
# replicated-vector allreduce path, against single-GPU results; row-sharded C3 products; weak-scaling bench line
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo18.txt 2>&1
timeout 900 python -m pytest tests/test_multi_gpu.py -x -q > gpurun_out/pytest18_multi.log 2>&1; echo "rc=$?" >> gpurun_out/pytest18_multi.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544 tools/dist_check.py > gpurun_out/dist_check18.json 2> gpurun_out/dist_check18.err; echo "rc=$?" >> gpurun_out/dist_check18.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29545 tools/bench_dist.py > gpurun_out/bench_dist18_n2.jsonl 2> gpurun_out/bench_dist18_n2.err; echo "rc=$?" >> gpurun_out/bench_dist18_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29546 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench18_n2.json 2> gpurun_out/bench18_n2.err; echo "rc=$?" >> gpurun_out/bench18_n2.err
echo done
