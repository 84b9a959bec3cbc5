#!/bin/bash
# round 2, GPU session 16 (1 GPU): operand size at which the x-blocked transposed SpMV starts to pay, and the block size
mkdir -p gpurun_out
timeout 900 python tools/xblock_threshold.py > gpurun_out/r2p_xblock_threshold.jsonl 2> gpurun_out/r2p_xblock_threshold.err
echo done
