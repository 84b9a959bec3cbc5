#!/bin/bash
# round 2, GPU session 14 (1 GPU): compute-sanitizer (memcheck, then racecheck on the shared-memory kernels) over the small
# GPU tests -- new kernels of this round included (x-blocked fold, sort keys, validation kernels, repack, bounce copies)
mkdir -p gpurun_out
K="known_answers or xblocked or hilbert_sort or sort_of_host_blocked or residency_cache_sees or validate or blocked_upload or cg_matches or linalg or noise_rhs or invalid_indices or file_loaders"
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 99 --print-limit 20 python -m pytest tests/test_gpu_parity.py -q -x -k "$K" > gpurun_out/r2n_memcheck.log 2>&1; echo "rc=$?" >> gpurun_out/r2n_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 99 --print-limit 20 python -m pytest tests/test_gpu_parity.py -q -x -k "known_answers or xblocked or cg_matches" > gpurun_out/r2n_racecheck.log 2>&1; echo "rc=$?" >> gpurun_out/r2n_racecheck.log
echo done
