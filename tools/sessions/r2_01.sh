#!/bin/bash
# round 2, GPU session 1: gather ceiling probe, parity after the thread-local tuning / stream-policy changes,
# stream kernel with and without L2 policies, ncu captures missing from round 1 (C3 transposed SpMV, C4, A' R=32)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,power.limit --format=csv > gpurun_out/r2a_smi.txt 2>&1
timeout 300 tools/_build/gather_probe 200 > gpurun_out/r2a_gather_probe.jsonl 2> gpurun_out/r2a_gather_probe.err
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q > gpurun_out/r2a_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r2a_pytest.log
FSB_TUNE_STREAM_POLICY=0 timeout 600 python tools/bench_all.py --only c3 --out gpurun_out/r2a_c3_nopol.jsonl > /dev/null 2> gpurun_out/r2a_c3_nopol.err
FSB_TUNE_STREAM_POLICY=1 timeout 600 python tools/bench_all.py --only c3 --out gpurun_out/r2a_c3_pol.jsonl > /dev/null 2> gpurun_out/r2a_c3_pol.err
# ncu (after the plain runs above exited): stream kernel, both policies
for pol in 0 1; do
  FSB_TUNE_STREAM_POLICY=$pol timeout 600 ncu --set full --clock-control none --import-source on -k regex:csr_stream_kernel -o /tmp/prof_spmv_pol$pol \
      python tools/prof_kernels.py --only spmv > gpurun_out/r2a_ncu_spmv_pol$pol.log 2>&1
  ncu -i /tmp/prof_spmv_pol$pol.ncu-rep --page raw --csv > gpurun_out/r2a_ncu_spmv_pol${pol}_raw.csv 2>/dev/null
done
# C4 (power-law columns): staged CSR kernel + native blocked / column-blocked kernels
timeout 300 python tools/prof_kernels.py --only formats > gpurun_out/r2a_plain_formats.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:csr_spmm_staged|blocked_spmm|cbcsr_spmm" -c 12 -o /tmp/prof_c4 \
    python tools/prof_kernels.py --only formats > gpurun_out/r2a_ncu_c4.log 2>&1
ncu -i /tmp/prof_c4.ncu-rep --page raw --csv > gpurun_out/r2a_ncu_c4_raw.csv 2>/dev/null
cp /tmp/prof_c4.ncu-rep gpurun_out/r2a_prof_c4.ncu-rep 2>/dev/null
# C5: the A'(AP) pass at R = 32 (staged kernel on the cached transpose) -- skip the autotune launches, take the iteration's two products
timeout 300 python tools/prof_kernels.py --only cg --cg-iters 2 > gpurun_out/r2a_plain_cg.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:csr_spmm_staged" -s 24 -c 8 -o /tmp/prof_cg_spmm \
    python tools/prof_kernels.py --only cg --cg-iters 3 > gpurun_out/r2a_ncu_cg_spmm.log 2>&1
ncu -i /tmp/prof_cg_spmm.ncu-rep --page raw --csv > gpurun_out/r2a_ncu_cg_spmm_raw.csv 2>/dev/null
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:csr_spmm_staged" --csv --log-file gpurun_out/r2a_launches_cg_spmm.csv \
    python tools/prof_kernels.py --only cg --cg-iters 3 > gpurun_out/r2a_ncu_cg_launches.log 2>&1
echo done
