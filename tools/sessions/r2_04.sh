#!/bin/bash
# round 2, GPU session 4: full GPU suite; pageable (malloc) drop-in timing vs bounce piece size / copy threads;
# x-blocked transposed SpMV (C3) timing + ncu
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2d_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r2d_pytest.log
for kb in 1024 4096 32768; do for th in 4 8 12; do
  echo "bounce_kb=$kb copy_threads=$th" >> gpurun_out/r2d_time_dropin.log
  FSB_BOUNCE_KB=$kb FSB_COPY_THREADS=$th timeout 300 tests/_build/time_dropin 10000000 1000000 200000000 32 4 >> gpurun_out/r2d_time_dropin.log 2>&1
done; done
echo "cache=fast bounce 4096 threads 8" >> gpurun_out/r2d_time_dropin.log
FSB_CACHE=fast timeout 300 tests/_build/time_dropin 10000000 1000000 200000000 32 4 >> gpurun_out/r2d_time_dropin.log 2>&1
timeout 600 python tools/bench_all.py --only c3 --out gpurun_out/r2d_c3_xblock.jsonl > /dev/null 2> gpurun_out/r2d_c3_xblock.err
FSB_TUNE_T_XBLOCK=0 timeout 600 python tools/bench_all.py --only c3 --out gpurun_out/r2d_c3_noxblock.jsonl > /dev/null 2> gpurun_out/r2d_c3_noxblock.err
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:csr_stream_kernel|fold_blocks" -o /tmp/prof_spmv_xb \
    python tools/prof_kernels.py --only spmv > gpurun_out/r2d_ncu_spmv_xb.log 2>&1
ncu -i /tmp/prof_spmv_xb.ncu-rep --page raw --csv > gpurun_out/r2d_ncu_spmv_xb_raw.csv 2>/dev/null
echo done
