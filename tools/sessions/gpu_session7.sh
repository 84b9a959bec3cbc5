#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest7.log 2>&1; echo "rc=$?" >> gpurun_out/pytest7.log
timeout 1500 python tools/bench_all.py --out gpurun_out/bench_all7.jsonl > gpurun_out/bench_all7.log 2>&1; echo "rc=$?" >> gpurun_out/bench_all7.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_r1_b.json 2> gpurun_out/bench_r1_b.err; echo "rc=$?" >> gpurun_out/bench_r1_b.err
timeout 300 python tools/sweep.py --workload c2 --R 1 --vals --reps 10 --combos "0,0,0,0,0,0;3,0,0,0,0,0;2,0,1,1,1,0;1,4,1,1,1,0;1,8,1,1,1,0" > gpurun_out/sweep7_spmv.log 2>&1
timeout 300 python tools/sweep.py --workload c2 --R 1 --vals --transpose --reps 10 --combos "0,0,0,0,0,0;3,0,0,0,0,0;1,8,1,1,1,0;1,32,1,1,1,0" > gpurun_out/sweep7_spmv_t.log 2>&1
timeout 300 python tools/sweep.py --workload c2 --R 4 --reps 10 --combos "0,0,0,0,0,0;3,0,0,0,0,0;2,0,2,2,1,0;2,0,1,4,1,0;1,8,2,2,1,0" > gpurun_out/sweep7_R4.log 2>&1
timeout 300 python tools/sweep.py --workload c2 --R 2 --reps 10 --combos "0,0,0,0,0,0;3,0,0,0,0,0;2,0,1,2,1,0;2,0,2,1,1,0;1,8,1,2,1,0" > gpurun_out/sweep7_R2.log 2>&1
echo done
