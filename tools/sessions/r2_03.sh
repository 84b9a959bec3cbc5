#!/bin/bash
# round 2, GPU session 3: full GPU test suite after the cache / host-copy / validation work, the new bench line (N = 1),
# the reference arm on the full matrix, refreshed gather probe
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2c_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r2c_pytest.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2c_bench_n1.json 2> gpurun_out/r2c_bench_n1.err; echo "rc=$?" >> gpurun_out/r2c_bench_n1.err
( time timeout 900 python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/r2c_bench_ref.json 2> gpurun_out/r2c_bench_ref.err ) 2> gpurun_out/r2c_bench_ref.time
timeout 300 tools/_build/gather_probe 200 > gpurun_out/r2c_gather_probe.jsonl 2> gpurun_out/r2c_gather_probe.err
nproc > gpurun_out/r2c_host.txt; lscpu | head -20 >> gpurun_out/r2c_host.txt; free -g >> gpurun_out/r2c_host.txt
echo done
