#!/bin/bash
# session 12: ncu coverage of every kernel family + refreshed bench launch list / staged-kernel capture + CG launch list
mkdir -p gpurun_out
KREGEX='regex:csr_|blocked_spmm|cbcsr_spmm|gram_|cg_|small_solve|stream_fixup|axpy_lambda|max_row'
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest12.log 2>&1; echo "rc=$?" >> gpurun_out/pytest12.log
CMD="python tools/prof_kernels.py --only spmv,ata,small_r,formats"
timeout 600 $CMD > gpurun_out/plain12a.log 2>&1 && \
timeout 1500 ncu --set full --clock-control none -k "$KREGEX" -o /tmp/prof_kernels $CMD > gpurun_out/ncu12a.log 2>&1
ncu -i /tmp/prof_kernels.ncu-rep --page raw --csv > gpurun_out/r1e_kernels_raw.csv 2>/dev/null
ncu -i /tmp/prof_kernels.ncu-rep --page details > gpurun_out/r1e_kernels_details.txt 2>/dev/null
CMDB="python bench.py --steps 5 --warmup 3 --no-cpu"
timeout 600 $CMDB > gpurun_out/plain12b.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KREGEX" --csv --log-file gpurun_out/r1e_launches_bench.csv $CMDB > gpurun_out/ncu12b.log 2>&1
timeout 600 $CMDB > gpurun_out/plain12c.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:csr_spmm_staged -s 8 -c 2 -o gpurun_out/prof_r1e_c2_staged $CMDB > gpurun_out/ncu12c.log 2>&1
ncu -i gpurun_out/prof_r1e_c2_staged.ncu-rep --page raw --csv > gpurun_out/r1e_c2_staged_raw.csv 2>/dev/null
ncu -i gpurun_out/prof_r1e_c2_staged.ncu-rep --page details > gpurun_out/r1e_c2_staged_details.txt 2>/dev/null
CMDC="python tools/prof_kernels.py --only cg"
timeout 600 $CMDC > gpurun_out/plain12d.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KREGEX" --csv --log-file gpurun_out/r1e_launches_cg.csv $CMDC > gpurun_out/ncu12d.log 2>&1
echo done
