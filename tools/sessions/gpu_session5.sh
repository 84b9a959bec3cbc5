#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "device_ or ragged" > gpurun_out/pytest5.log 2>&1; echo "rc=$?" >> gpurun_out/pytest5.log
timeout 1500 python tools/bench_all.py --out gpurun_out/bench_all.jsonl > gpurun_out/bench_all.log 2>&1; echo "rc=$?" >> gpurun_out/bench_all.log
for R in 8 16; do
timeout 300 python tools/sweep.py --workload c2 --R $R --reps 5 > gpurun_out/sweep_c2_R$R.log 2>&1
done
timeout 300 python tools/sweep.py --workload c2 --R 1 --vals --reps 10 --combos "0,0,0,0,0,0;2,0,1,1,1,0;2,0,1,1,1,128;2,0,1,1,1,512;2,0,1,1,1,1024;1,4,1,1,1,0;1,8,1,1,1,0;1,16,1,1,1,0;1,32,1,1,1,0;1,2,1,1,1,0" > gpurun_out/sweep_c3_spmv.log 2>&1
timeout 300 python tools/sweep.py --workload c2 --R 1 --vals --transpose --reps 10 --combos "0,0,0,0,0,0;2,0,1,1,1,0;2,0,1,1,1,64;2,0,1,1,1,16;1,8,1,1,1,0;1,32,1,1,1,0" > gpurun_out/sweep_c3_spmv_t.log 2>&1
CMD="python bench.py --steps 2 --warmup 3 --no-cpu"
$CMD > gpurun_out/plain5.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:csr_spmm_staged -s 3 -c 1 -o gpurun_out/prof_c2_staged $CMD > gpurun_out/ncu_staged.log 2>&1
echo done
