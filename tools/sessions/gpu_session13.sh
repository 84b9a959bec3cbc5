#!/bin/bash
# session 13: batched-gather staged kernel (launch-bounds fix), DMMA dense CG kernels, fused lambda epilogue
mkdir -p gpurun_out
KREGEX='regex:csr_|blocked_spmm|cbcsr_spmm|gram_|cg_|small_solve|stream_fixup|axpy_lambda|max_row'
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest13.log 2>&1; echo "rc=$?" >> gpurun_out/pytest13.log
COMBOS="2,0,16,2,1,0;2,0,8,2,2,0;2,0,32,1,1,0;2,0,8,4,1,0;2,0,16,2,1,64;2,0,8,2,2,64;2,0,16,2,1,256"
timeout 600 python tools/sweep.py --workload c2 --reps 8 --combos "$COMBOS" > gpurun_out/sweep13_c2_base.log 2>&1
for v in u12m3 u16m2 u4m6; do
  FSB_LIB=$PWD/libfastsparse_b200/lib/libfastsparse_b200_$v.so timeout 600 python tools/sweep.py --workload c2 --reps 8 --combos "2,0,16,2,1,0;2,0,8,2,2,0" > gpurun_out/sweep13_c2_$v.log 2>&1
done
timeout 600 python tools/sweep.py --workload c2 --dist 1 --reps 8 --combos "2,0,16,2,1,0;2,0,8,2,2,0" > gpurun_out/sweep13_c4.log 2>&1
timeout 600 python tools/sweep.py --workload c2 --transpose --reps 5 --combos "2,0,16,2,1,0;2,0,8,2,2,0" > gpurun_out/sweep13_c2_t.log 2>&1
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_r1_e.json 2> gpurun_out/bench_r1_e.err; echo "rc=$?" >> gpurun_out/bench_r1_e.err
timeout 1500 python tools/bench_all.py --out gpurun_out/bench_all13.jsonl > gpurun_out/bench_all13.log 2>&1; echo "rc=$?" >> gpurun_out/bench_all13.log
CMDC="python tools/prof_kernels.py --only cg"
timeout 600 $CMDC > gpurun_out/plain13d.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KREGEX" --csv --log-file gpurun_out/r1f_launches_cg.csv $CMDC > gpurun_out/ncu13d.log 2>&1
echo done
