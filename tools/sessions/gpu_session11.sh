#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest11.log 2>&1; echo "rc=$?" >> gpurun_out/pytest11.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_r1_d.json 2> gpurun_out/bench_r1_d.err; echo "rc=$?" >> gpurun_out/bench_r1_d.err
timeout 1500 python tools/bench_all.py --out gpurun_out/bench_all11.jsonl > gpurun_out/bench_all11.log 2>&1; echo "rc=$?" >> gpurun_out/bench_all11.log
echo done
