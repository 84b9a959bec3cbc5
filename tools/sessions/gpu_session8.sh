#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest8.log 2>&1; echo "rc=$?" >> gpurun_out/pytest8.log
timeout 300 python tools/sweep.py --workload c2 --R 1 --vals --reps 10 --combos "0,0,0,0,0,0;3,0,0,0,0,0;1,4,1,1,1,0" > gpurun_out/sweep8_spmv.log 2>&1
timeout 300 python tools/sweep.py --workload c2 --R 1 --vals --transpose --reps 10 --combos "0,0,0,0,0,0;1,32,1,1,1,0" > gpurun_out/sweep8_spmv_t.log 2>&1
timeout 300 python tools/sweep.py --workload c2 --R 1 --dist 1 --vals --transpose --reps 10 --combos "0,0,0,0,0,0;1,32,1,1,1,0;2,0,1,1,1,0" > gpurun_out/sweep8_spmv_t_powerlaw.log 2>&1
timeout 1500 python tools/bench_all.py --out gpurun_out/bench_all8.jsonl > gpurun_out/bench_all8.log 2>&1; echo "rc=$?" >> gpurun_out/bench_all8.log
CMD="python tools/sweep.py --workload c2 --R 1 --vals --reps 3 --combos 0,0,0,0,0,0"
$CMD > gpurun_out/plain8.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:csr_stream_kernel -s 2 -c 1 -o gpurun_out/prof_c3_stream $CMD > gpurun_out/ncu_stream.log 2>&1
echo done
