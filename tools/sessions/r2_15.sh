#!/bin/bash
# round 2, GPU session 15 (1 GPU): packed host copies of blocked structures -- full suite, C4 table with the drop-in sort timing
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2o_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r2o_pytest.log
timeout 900 python tools/bench_all.py --only c4 --out gpurun_out/r2o_bench_all_c4.jsonl > /dev/null 2> gpurun_out/r2o_bench_all.err
echo done
