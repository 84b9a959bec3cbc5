#!/bin/bash
# session 40 (1 GPU): final round-1 code (TMA index staging) -- tests, smoke, bench (both arms), all-config bench, sampler loop; ncu launch list and
# full capture of bench.py with the launch configuration pinned to what the plain run's autotune picks (two passes, deep build)
mkdir -p gpurun_out
KREGEX='regex:csr_|blocked_spmm|cbcsr_spmm|gram_|cg_|small_solve|stream_fixup|axpy_lambda|max_row|randn'
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest40.log 2>&1; echo "rc=$?" >> gpurun_out/pytest40.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke40.log 2>&1; echo "rc=$?" >> gpurun_out/smoke40.log
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_r1_l_ref.json 2> gpurun_out/bench_r1_l_ref.err; echo "rc=$?" >> gpurun_out/bench_r1_l_ref.err
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_r1_l.json 2> gpurun_out/bench_r1_l.err; echo "rc=$?" >> gpurun_out/bench_r1_l.err
timeout 1500 python tools/bench_all.py --out gpurun_out/bench_all40.jsonl > gpurun_out/bench_all40.log 2>&1; echo "rc=$?" >> gpurun_out/bench_all40.log
timeout 900 python tools/macau_loop.py --samples 5 > gpurun_out/macau40.json 2> gpurun_out/macau40.err
CMDB="python bench.py --steps 5 --warmup 3 --no-cpu --tune 2,0,8,2,2,0,1"
timeout 600 $CMDB > gpurun_out/plain40b.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KREGEX" --csv --log-file gpurun_out/r1l_launches_bench.csv $CMDB > gpurun_out/ncu40b.log 2>&1
timeout 600 $CMDB > gpurun_out/plain40c.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:csr_spmm_staged -s 10 -c 2 -o gpurun_out/prof_r1l_c2_staged $CMDB > gpurun_out/ncu40c.log 2>&1
ncu -i gpurun_out/prof_r1l_c2_staged.ncu-rep --page raw --csv > gpurun_out/r1l_c2_staged_raw.csv 2>/dev/null
ncu -i gpurun_out/prof_r1l_c2_staged.ncu-rep --page details > gpurun_out/r1l_c2_staged_details.txt 2>/dev/null
CMDC="python tools/prof_kernels.py --only cg --cg-iters 2"
timeout 600 $CMDC > gpurun_out/plain40d.log 2>&1 && \
timeout 900 ncu --set full --clock-control none -k "regex:gram_|cg_mix|small_solve" -o /tmp/prof_cg $CMDC > gpurun_out/ncu40d.log 2>&1
ncu -i /tmp/prof_cg.ncu-rep --page raw --csv > gpurun_out/r1l_cg_dense_raw.csv 2>/dev/null
echo done
