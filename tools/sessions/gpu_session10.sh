#!/bin/bash
mkdir -p gpurun_out
for v in su8 su2; do
  FSB_LIB=$PWD/libfastsparse_b200/lib/libfastsparse_b200_$v.so timeout 600 python tools/sweep.py --workload c2 --reps 8 --combos "2,0,16,2,1,0;2,0,32,1,1,0;2,0,8,2,2,0;2,0,16,2,1,64;2,0,16,2,1,256" > gpurun_out/sweep10_c2_$v.log 2>&1
done
timeout 600 python tools/sweep.py --workload c2 --reps 8 --combos "2,0,16,2,1,0;2,0,32,1,1,0;2,0,8,2,2,0;2,0,16,2,1,64;2,0,16,2,1,256" > gpurun_out/sweep10_c2_base.log 2>&1
CMD="python tools/sweep.py --workload c2 --R 1 --reps 3 --combos 0,0,0,0,0,0"
$CMD > gpurun_out/plain10.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:csr_stream_kernel -s 2 -c 1 -o gpurun_out/prof_spmv_bin $CMD > gpurun_out/ncu_spmv_bin.log 2>&1
CMD2="python tools/sweep.py --workload c2 --R 1 --vals --reps 3 --combos 0,0,0,0,0,0"
$CMD2 > gpurun_out/plain10b.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:csr_stream_kernel -s 2 -c 1 -o gpurun_out/prof_spmv_dbl $CMD2 > gpurun_out/ncu_spmv_dbl.log 2>&1
echo done
