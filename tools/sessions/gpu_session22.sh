#!/bin/bash
# session 22 (1 GPU): final-code tests, bench, all-config bench, sampler loop; ncu: launch list of bench.py, full capture of
# the staged SpMM as bench.py runs it, full capture of the tensor-core CG passes
mkdir -p gpurun_out
KREGEX='regex:csr_|blocked_spmm|cbcsr_spmm|gram_|cg_|small_solve|stream_fixup|axpy_lambda|max_row|randn'
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest22.log 2>&1; echo "rc=$?" >> gpurun_out/pytest22.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_r1_i.json 2> gpurun_out/bench_r1_i.err; echo "rc=$?" >> gpurun_out/bench_r1_i.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_r1_i_ref.json 2> gpurun_out/bench_r1_i_ref.err; echo "rc=$?" >> gpurun_out/bench_r1_i_ref.err
timeout 1500 python tools/bench_all.py --out gpurun_out/bench_all22.jsonl > gpurun_out/bench_all22.log 2>&1; echo "rc=$?" >> gpurun_out/bench_all22.log
timeout 900 python tools/macau_loop.py --samples 5 > gpurun_out/macau22.json 2> gpurun_out/macau22.err
CMDB="python bench.py --steps 5 --warmup 3 --no-cpu"
timeout 600 $CMDB > gpurun_out/plain22b.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KREGEX" --csv --log-file gpurun_out/r1h_launches_bench.csv $CMDB > gpurun_out/ncu22b.log 2>&1
timeout 600 $CMDB > gpurun_out/plain22c.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:csr_spmm_staged -s 16 -c 2 -o gpurun_out/prof_r1h_c2_staged $CMDB > gpurun_out/ncu22c.log 2>&1
ncu -i gpurun_out/prof_r1h_c2_staged.ncu-rep --page raw --csv > gpurun_out/r1h_c2_staged_raw.csv 2>/dev/null
ncu -i gpurun_out/prof_r1h_c2_staged.ncu-rep --page details > gpurun_out/r1h_c2_staged_details.txt 2>/dev/null
CMDC="python tools/prof_kernels.py --only cg --cg-iters 2"
timeout 600 $CMDC > gpurun_out/plain22d.log 2>&1 && \
timeout 900 ncu --set full --clock-control none -k "regex:gram_|cg_mix|small_solve" -o /tmp/prof_cg $CMDC > gpurun_out/ncu22d.log 2>&1
ncu -i /tmp/prof_cg.ncu-rep --page raw --csv > gpurun_out/r1h_cg_dense_raw.csv 2>/dev/null
ncu -i /tmp/prof_cg.ncu-rep --page details > gpurun_out/r1h_cg_dense_details.txt 2>/dev/null
echo done
