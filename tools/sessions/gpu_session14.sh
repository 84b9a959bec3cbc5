#!/bin/bash
# session 14 (1 GPU): gather-issue variants of the staged kernel on C2 / C4 / transpose, bench_all regression
mkdir -p gpurun_out
COMBOS="2,0,16,2,1,0;2,0,8,2,2,0"
for v in base b4 b4c b6; do
  if [ $v = base ]; then unset FSB_LIB; else export FSB_LIB=$PWD/libfastsparse_b200/lib/libfastsparse_b200_$v.so; fi
  timeout 600 python tools/sweep.py --workload c2 --reps 8 --combos "$COMBOS" > gpurun_out/sweep14_c2_$v.log 2>&1
  timeout 600 python tools/sweep.py --workload c2 --dist 1 --reps 8 --combos "$COMBOS" > gpurun_out/sweep14_c4_$v.log 2>&1
  timeout 600 python tools/sweep.py --workload c2 --transpose --reps 5 --combos "$COMBOS" > gpurun_out/sweep14_c2t_$v.log 2>&1
done
unset FSB_LIB
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest14.log 2>&1; echo "rc=$?" >> gpurun_out/pytest14.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_r1_f.json 2> gpurun_out/bench_r1_f.err; echo "rc=$?" >> gpurun_out/bench_r1_f.err
timeout 1500 python tools/bench_all.py --out gpurun_out/bench_all14.jsonl > gpurun_out/bench_all14.log 2>&1; echo "rc=$?" >> gpurun_out/bench_all14.log
echo done
