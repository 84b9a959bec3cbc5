#!/bin/bash
# session 16 (1 GPU): same-box comparison of the session-11 staged kernel ("old") with the current gather-issue variants
mkdir -p gpurun_out
COMBOS="2,0,16,2,1,0;2,0,8,2,2,0"
for v in base old b4 b4c b3 b5 old base; do
  if [ $v = base ]; then unset FSB_LIB; else export FSB_LIB=$PWD/libfastsparse_b200/lib/libfastsparse_b200_$v.so; fi
  timeout 600 python tools/sweep.py --workload c2 --reps 8 --combos "$COMBOS" >> gpurun_out/sweep16_c2_$v.log 2>&1
  timeout 600 python tools/sweep.py --workload c2 --dist 1 --reps 8 --combos "$COMBOS" >> gpurun_out/sweep16_c4_$v.log 2>&1
done
unset FSB_LIB
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest16.log 2>&1; echo "rc=$?" >> gpurun_out/pytest16.log
timeout 900 python tools/macau_loop.py --samples 3 > gpurun_out/macau16.json 2> gpurun_out/macau16.err
echo done
