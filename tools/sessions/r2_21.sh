#!/bin/bash
# round 2, GPU session 21 (1 GPU): thread-serial row reduction of the merge-path stream kernel (knob stream_reduce) --
# full GPU test suite with the new default, then C3 with the old (0) and new (1) reduction on the same box
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2u_pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/r2u_pytest_gpu.log
FSB_TUNE_STREAM_REDUCE=0 timeout 300 python tools/bench_all.py --only c3 --out gpurun_out/r2u_c3_rowlanes.jsonl > /dev/null 2> gpurun_out/r2u_c3_rowlanes.err
FSB_TUNE_STREAM_REDUCE=1 timeout 300 python tools/bench_all.py --only c3 --out gpurun_out/r2u_c3_serial.jsonl > /dev/null 2> gpurun_out/r2u_c3_serial.err
FSB_TUNE_STREAM_REDUCE=0 timeout 300 python tools/bench_all.py --only c3 --out gpurun_out/r2u_c3_rowlanes_b.jsonl > /dev/null 2>&1
FSB_TUNE_STREAM_REDUCE=1 timeout 300 python tools/bench_all.py --only c3 --out gpurun_out/r2u_c3_serial_b.jsonl > /dev/null 2>&1
tail -3 gpurun_out/r2u_pytest_gpu.log
for f in gpurun_out/r2u_c3_*.jsonl; do echo $f; python - "$f" <<'PY'
import json,sys
for l in open(sys.argv[1]):
    d=json.loads(l); print("  %-70s %.3f ms" % (d["config"][:70], d["ms"]))
PY
done
