#!/bin/bash
# round 2, GPU session 11 (1 GPU): ncu of the native blocked / column-blocked kernels at C4 (the evidence behind their
# retirement from the product path), refreshed gather probe on this box
mkdir -p gpurun_out
timeout 300 python tools/prof_kernels.py --only formats > gpurun_out/r2k_plain_formats.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:blocked_spmm|cbcsr_spmm" -c 4 -o /tmp/prof_c4_native \
    python tools/prof_kernels.py --only formats > gpurun_out/r2k_ncu_c4_native.log 2>&1
ncu -i /tmp/prof_c4_native.ncu-rep --page raw --csv > gpurun_out/r2k_ncu_c4_native_raw.csv 2>/dev/null
echo done
