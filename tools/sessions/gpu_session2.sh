#!/bin/bash
# second GPU session: full-size parity, bench, variant sweeps, ncu captures
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py::test_full_size_c2_properties -x -q > gpurun_out/pytest_full.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_full.log
timeout 900 python bench.py --steps 20 --warmup 3 --tune 16,16,2,1 > gpurun_out/bench_r1_a.json 2> gpurun_out/bench_r1_a.err; echo "rc=$?" >> gpurun_out/bench_r1_a.err
COMBOS="32,8,4,1;16,8,4,1;8,8,4,1;16,16,2,1;32,16,2,1;32,32,1,1"
for v in noel u8 u2 mb3; do
  FSB_LIB=$PWD/libfastsparse_b200/lib/libfastsparse_b200_$v.so timeout 600 python tools/sweep.py --workload c2 --reps 5 --combos "$COMBOS" > gpurun_out/sweep_c2_$v.log 2>&1
done
CMD="python bench.py --steps 2 --warmup 3 --no-cpu --tune 16,16,2,1"
$CMD > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_bench.csv $CMD > gpurun_out/ncu_launches.log 2>&1
$CMD > gpurun_out/plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:csr_spmm -s 3 -c 1 -o gpurun_out/prof_c2_spmm $CMD > gpurun_out/ncu_full.log 2>&1
echo done
