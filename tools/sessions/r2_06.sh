#!/bin/bash
# round 2, GPU session 6 (1 GPU): staged kernel with fewer gathers in flight per lane and more resident warps
# (the probe's best point), C3 / C4 tables with the x-blocked transpose, view costs, preprocessing pipeline
mkdir -p gpurun_out
C="2,0,8,2,2,0,1;2,0,8,2,2,0,0;2,0,16,2,1,0,1;2,0,16,2,1,0,0"
for v in base u4m6 u4m8 u6m5; do
  if [ $v = base ]; then L=""; else L="$PWD/libfastsparse_b200/lib/libfastsparse_b200_$v.so"; fi
  FSB_LIB=$L timeout 400 python tools/sweep.py --workload c2 --reps 10 --combos "$C" --out gpurun_out/r2f_sweep_c2_$v.json > gpurun_out/r2f_sweep_c2_$v.log 2>&1
  FSB_LIB=$L timeout 400 python tools/sweep.py --workload c2 --dist 1 --reps 10 --combos "$C" --out gpurun_out/r2f_sweep_c4_$v.json > gpurun_out/r2f_sweep_c4_$v.log 2>&1
  FSB_LIB=$L timeout 400 python tools/sweep.py --workload c2 --transpose --reps 6 --combos "2,0,16,2,1,0,1;2,0,16,2,1,0,0" --out gpurun_out/r2f_sweep_c2t_$v.json > gpurun_out/r2f_sweep_c2t_$v.log 2>&1
done
timeout 900 python tools/bench_all.py --only c3,c4 --out gpurun_out/r2f_bench_all_c3c4.jsonl > /dev/null 2> gpurun_out/r2f_bench_all.err
echo done
