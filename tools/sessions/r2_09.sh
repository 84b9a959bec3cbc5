#!/bin/bash
# round 2, GPU session 9 (1 GPU): the state the driver will see -- build check, full GPU suite, smoke, bench line
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2i_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r2i_pytest.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2i_smoke.log 2>&1; echo "rc=$?" >> gpurun_out/r2i_smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2i_bench_n1.json 2> gpurun_out/r2i_bench_n1.err; echo "rc=$?" >> gpurun_out/r2i_bench_n1.err
timeout 600 python tools/bench_all.py --only c3,c5 --out gpurun_out/r2i_bench_all_c3c5.jsonl > /dev/null 2> gpurun_out/r2i_bench_all.err
timeout 300 python tools/macau_loop.py > gpurun_out/r2i_macau_loop.json 2> gpurun_out/r2i_macau_loop.err
echo done
