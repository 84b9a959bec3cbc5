#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_multi_gpu.py -x -q > gpurun_out/pytest_multi.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_multi.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544 tools/dist_check.py > gpurun_out/dist_check.json 2> gpurun_out/dist_check.err; echo "rc=$?" >> gpurun_out/dist_check.err
timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q -k "ragged or staged or golden" > gpurun_out/pytest_staged.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_staged.log
timeout 600 python tools/sweep.py --workload c2 --reps 5 --out gpurun_out/sweep2_c2.json > gpurun_out/sweep2_c2.log 2>&1
echo done
