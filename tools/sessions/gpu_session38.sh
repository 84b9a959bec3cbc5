#!/bin/bash
# session 38 (1 GPU): TMA bulk-copy staging of the index run (FSB_STAGED_TMA=1 variant) against per-thread staging, same box
mkdir -p gpurun_out
COMBOS="2,0,16,2,1,0,0;2,0,16,2,1,0,1;2,0,8,2,2,0,0;2,0,8,2,2,0,1"
for v in base tma base tma; do
  if [ $v = base ]; then unset FSB_LIB; else export FSB_LIB=$PWD/libfastsparse_b200/lib/libfastsparse_b200_$v.so; fi
  timeout 600 python tools/sweep.py --workload c2 --reps 8 --combos "$COMBOS" >> gpurun_out/sweep38_c2_$v.log 2>&1
done
for v in base tma; do
  if [ $v = base ]; then unset FSB_LIB; else export FSB_LIB=$PWD/libfastsparse_b200/lib/libfastsparse_b200_$v.so; fi
  timeout 600 python tools/sweep.py --workload c2 --dist 1 --reps 8 --combos "$COMBOS" >> gpurun_out/sweep38_c4_$v.log 2>&1
  timeout 600 python tools/sweep.py --workload c2 --vals --reps 5 --combos "$COMBOS" >> gpurun_out/sweep38_c2v_$v.log 2>&1
  timeout 600 python tools/sweep.py --workload c2 --R 8 --reps 5 --combos "2,0,4,2,1,0,0;2,0,4,2,1,0,1" >> gpurun_out/sweep38_R8_$v.log 2>&1
done
export FSB_LIB=$PWD/libfastsparse_b200/lib/libfastsparse_b200_tma.so
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q > gpurun_out/pytest38_tma.log 2>&1; echo "rc=$?" >> gpurun_out/pytest38_tma.log
echo done
