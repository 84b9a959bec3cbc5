#!/bin/bash
# round 2, GPU session 7 (1 GPU): stream-kernel row-reduction variants, full GPU suite, bench line, ncu of the final
# headline kernel (traffic for bench.py's roofline) + launch list of the same command
mkdir -p gpurun_out
for v in base sgl2 sgl4; do
  if [ $v = base ]; then L=""; else L="$PWD/libfastsparse_b200/lib/libfastsparse_b200_$v.so"; fi
  FSB_LIB=$L timeout 400 python tools/bench_all.py --only c3 --out gpurun_out/r2g_c3_$v.jsonl > /dev/null 2> gpurun_out/r2g_c3_$v.err
done
C="2,0,8,2,2,0,1;2,0,8,2,2,0,0;2,0,16,2,1,0,1;2,0,16,2,1,0,0"
for v in base lds128; do
  if [ $v = base ]; then L=""; else L="$PWD/libfastsparse_b200/lib/libfastsparse_b200_$v.so"; fi
  FSB_LIB=$L timeout 400 python tools/sweep.py --workload c2 --reps 10 --combos "$C" --out gpurun_out/r2g_sweep_c2_$v.json > gpurun_out/r2g_sweep_c2_$v.log 2>&1
  FSB_LIB=$L timeout 400 python tools/sweep.py --workload c2 --dist 1 --reps 10 --combos "$C" --out gpurun_out/r2g_sweep_c4_$v.json > gpurun_out/r2g_sweep_c4_$v.log 2>&1
  FSB_LIB=$L timeout 400 python tools/sweep.py --workload c2 --vals --reps 6 --combos "2,0,8,2,2,0,1;2,0,16,2,1,0,1" --out gpurun_out/r2g_sweep_c2v_$v.json > gpurun_out/r2g_sweep_c2v_$v.log 2>&1
  FSB_LIB=$L timeout 400 python tools/sweep.py --workload c2 --R 8 --reps 10 --combos "0,0,0,0,0,0" --out gpurun_out/r2g_sweep_c2_R8_$v.json > gpurun_out/r2g_sweep_c2_R8_$v.log 2>&1
done
FSB_LIB=$PWD/libfastsparse_b200/lib/libfastsparse_b200_lds128.so timeout 600 python -m pytest tests/test_gpu_parity.py -q -x > gpurun_out/r2g_pytest_lds128.log 2>&1; echo "rc=$?" >> gpurun_out/r2g_pytest_lds128.log
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2g_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r2g_pytest.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2g_bench_n1.json 2> gpurun_out/r2g_bench_n1.err; echo "rc=$?" >> gpurun_out/r2g_bench_n1.err
CMDB="python bench.py --steps 3 --warmup 3 --no-cpu --tune 2,0,8,2,2,0,1"
timeout 600 $CMDB > gpurun_out/r2g_plain_bench_tuned.json 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:csr_spmm_staged|repack" --csv --log-file gpurun_out/r2g_launches_bench_c2.csv $CMDB > gpurun_out/r2g_ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:csr_spmm_staged -s 10 -c 2 -o gpurun_out/r2g_prof_c2_staged $CMDB > gpurun_out/r2g_ncu_full.log 2>&1
ncu -i gpurun_out/r2g_prof_c2_staged.ncu-rep --page raw --csv > gpurun_out/r2g_ncu_c2_staged_raw.csv 2>/dev/null
echo done
