#!/bin/bash
# round 2, GPU session 2: contiguous (repacked) column slabs of X for the C2 / C4 products
mkdir -p gpurun_out
C="2,0,8,2,2,0,1;2,0,8,2,2,0,0;2,0,0,0,1,0,1,2;2,0,0,0,1,0,0,2;2,0,0,0,1,0,1,4;2,0,0,0,1,0,0,4;2,0,16,2,1,0,0;2,0,16,2,1,0,1"
timeout 600 python tools/sweep.py --workload c2 --reps 10 --combos "$C" --out gpurun_out/r2b_sweep_c2_xslabs.json > gpurun_out/r2b_sweep_c2_xslabs.log 2>&1
timeout 600 python tools/sweep.py --workload c2 --dist 1 --reps 10 --combos "$C" --out gpurun_out/r2b_sweep_c4_xslabs.json > gpurun_out/r2b_sweep_c4_xslabs.log 2>&1
timeout 600 python tools/sweep.py --workload c2 --vals --reps 10 --combos "$C" --out gpurun_out/r2b_sweep_c2v_xslabs.json > gpurun_out/r2b_sweep_c2v_xslabs.log 2>&1
timeout 600 python tools/sweep.py --workload c2 --R 16 --reps 10 --combos "2,0,8,2,1,0,0;2,0,8,2,1,0,1;2,0,0,0,1,0,1,2;2,0,0,0,1,0,0,2" --out gpurun_out/r2b_sweep_c2_R16_xslabs.json > gpurun_out/r2b_sweep_c2_R16.log 2>&1
echo done
