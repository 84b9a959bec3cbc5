#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest9.log 2>&1; echo "rc=$?" >> gpurun_out/pytest9.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_r1_c.json 2> gpurun_out/bench_r1_c.err; echo "rc=$?" >> gpurun_out/bench_r1_c.err
timeout 300 python tools/sweep.py --workload c2 --R 1 --vals --reps 10 --combos "0,0,0,0,0,0;1,4,1,1,1,0" > gpurun_out/sweep9_spmv_dbl.log 2>&1
timeout 300 python tools/sweep.py --workload c2 --R 1 --reps 10 --combos "0,0,0,0,0,0;1,4,1,1,1,0" > gpurun_out/sweep9_spmv_bin.log 2>&1
timeout 300 python tools/sweep.py --workload c2 --R 1 --transpose --reps 10 --combos "0,0,0,0,0,0;1,32,1,1,1,0" > gpurun_out/sweep9_spmv_bin_t.log 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29555 tools/bench_dist.py > gpurun_out/bench_dist_n2.jsonl 2> gpurun_out/bench_dist_n2.err; echo "rc=$?" >> gpurun_out/bench_dist_n2.err
timeout 900 python -m pytest tests/test_multi_gpu.py -x -q > gpurun_out/pytest9_multi.log 2>&1; echo "rc=$?" >> gpurun_out/pytest9_multi.log
echo done
