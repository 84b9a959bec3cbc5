#!/bin/bash
# session 17 (1 GPU): lean / deep builds of the staged kernel in one library + per-handle autotune, vs the session-11 kernel
mkdir -p gpurun_out
COMBOS="2,0,16,2,1,0,0;2,0,16,2,1,0,1;2,0,8,2,2,0,0;2,0,8,2,2,0,1;0,0,0,0,0,0,-1"
timeout 600 python tools/sweep.py --workload c2 --reps 8 --combos "$COMBOS" > gpurun_out/sweep17_c2.log 2>&1
timeout 600 python tools/sweep.py --workload c2 --dist 1 --reps 8 --combos "$COMBOS" > gpurun_out/sweep17_c4.log 2>&1
timeout 600 python tools/sweep.py --workload c2 --transpose --reps 5 --combos "$COMBOS" > gpurun_out/sweep17_c2t.log 2>&1
timeout 600 python tools/sweep.py --workload c2 --vals --reps 5 --combos "$COMBOS" > gpurun_out/sweep17_c2v.log 2>&1
timeout 600 python tools/sweep.py --workload c2 --R 16 --reps 5 --combos "2,0,8,2,1,0,0;2,0,8,2,1,0,1;0,0,0,0,0,0,-1" > gpurun_out/sweep17_c2_R16.log 2>&1
timeout 600 python tools/sweep.py --workload c2 --R 8 --reps 5 --combos "2,0,4,2,1,0,0;2,0,4,2,1,0,1;0,0,0,0,0,0,-1" > gpurun_out/sweep17_c2_R8.log 2>&1
FSB_LIB=$PWD/libfastsparse_b200/lib/libfastsparse_b200_old.so timeout 600 python tools/sweep.py --workload c2 --reps 8 --combos "2,0,16,2,1,0;2,0,8,2,2,0" > gpurun_out/sweep17_c2_old.log 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest17.log 2>&1; echo "rc=$?" >> gpurun_out/pytest17.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_r1_h.json 2> gpurun_out/bench_r1_h.err; echo "rc=$?" >> gpurun_out/bench_r1_h.err
timeout 1500 python tools/bench_all.py --out gpurun_out/bench_all17.jsonl > gpurun_out/bench_all17.log 2>&1; echo "rc=$?" >> gpurun_out/bench_all17.log
echo done
