#!/bin/bash
# round 2, GPU session 39 (1 GPU): texture-pipe gathers as shipped -- full GPU suite, C3 table, ncu capture of the R = 1 kernels, bench line
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2z2_pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/r2z2_pytest_gpu.log
tail -3 gpurun_out/r2z2_pytest_gpu.log
timeout 300 python tools/bench_all.py --only c3 --out gpurun_out/r2z2_bench_all_c3.jsonl > /dev/null 2> gpurun_out/r2z2_bench_all_c3.err
python - <<'PY'
import json
for l in open("gpurun_out/r2z2_bench_all_c3.jsonl"):
    d=json.loads(l); print("  %-70s %.3f ms" % (d["config"][:70], d["ms"]))
PY
timeout 900 python bench.py > gpurun_out/r2z2_bench_c2.json 2> gpurun_out/r2z2_bench.err; echo "bench rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"csr_stream|fold_blocks" -c 10 -o gpurun_out/r2z2_ncu_stream python tools/prof_kernels.py --only spmv > gpurun_out/r2z2_ncu_stream.log 2>&1
ncu -i gpurun_out/r2z2_ncu_stream.ncu-rep --page raw --csv > gpurun_out/r2z2_ncu_stream_raw.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/r2z2_ncu_stream_raw.csv "R = 1 at C3: TMA-fed merge-path stream kernel with texture-pipe gathers (as shipped): double SpMV, forced-stream runs, x-blocked transposed SpMV" > gpurun_out/r2z2_ncu_stream.md
tail -12 gpurun_out/r2z2_ncu_stream.md
