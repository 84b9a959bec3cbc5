#!/bin/bash
# round 2, GPU session 10 (4 GPUs): the final code with its defaults at N = 4 and N = 2 (bench line incl. collectives
# block, dist_check through pytest), what the driver's scaling run will execute
mkdir -p gpurun_out
for N in 4 2; do
  TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
  timeout 900 $TR --master-port 2951$N bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2j_bench_n$N.json 2> gpurun_out/r2j_bench_n$N.err; echo "rc=$?" >> gpurun_out/r2j_bench_n$N.err
done
timeout 600 python -m pytest tests/test_multi_gpu.py -q > gpurun_out/r2j_pytest_multi.log 2>&1; echo "rc=$?" >> gpurun_out/r2j_pytest_multi.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29519 tools/d2h_probe.py > gpurun_out/r2j_d2h_probe_n4.json 2> /dev/null
timeout 600 python bench.py --impl reference --gpus 4 --steps 3 --warmup 1 > gpurun_out/r2j_bench_ref.json 2> gpurun_out/r2j_bench_ref.err
echo done
