#!/bin/bash
# session 15 (1 GPU): gather-issue variants of the staged kernel (same box, with a copy-bandwidth calibration line), loader tests
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.mem,clocks.max.sm,clocks.max.mem,power.limit,temperature.gpu --format=csv > gpurun_out/gpu15.txt 2>&1
COMBOS="2,0,16,2,1,0;2,0,8,2,2,0"
for v in base b4 b4c b6 b8 base; do
  if [ $v = base ]; then unset FSB_LIB; else export FSB_LIB=$PWD/libfastsparse_b200/lib/libfastsparse_b200_$v.so; fi
  timeout 600 python tools/sweep.py --workload c2 --reps 8 --combos "$COMBOS" >> gpurun_out/sweep15_c2_$v.log 2>&1
  timeout 600 python tools/sweep.py --workload c2 --dist 1 --reps 8 --combos "$COMBOS" >> gpurun_out/sweep15_c4_$v.log 2>&1
done
unset FSB_LIB
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest15.log 2>&1; echo "rc=$?" >> gpurun_out/pytest15.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_r1_g.json 2> gpurun_out/bench_r1_g.err; echo "rc=$?" >> gpurun_out/bench_r1_g.err
echo done
