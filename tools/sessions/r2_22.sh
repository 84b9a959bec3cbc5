#!/bin/bash
# round 2, GPU session 22 (1 GPU): TMA-fed merge-path stream kernel (knob stream_tma) -- R = 1 parity tests with the new
# default, C3 with the per-thread-load form (0) and the TMA-fed form (1) on the same box, then an ncu capture of both
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "ragged or xblock or golden or stream or transposed or skew or alias or degenerate or empty" > gpurun_out/r2v_pytest_stream.log 2>&1; echo "rc=$?" >> gpurun_out/r2v_pytest_stream.log
tail -3 gpurun_out/r2v_pytest_stream.log
for rep in a b; do
FSB_TUNE_STREAM_TMA=0 timeout 300 python tools/bench_all.py --only c3 --out gpurun_out/r2v_c3_ldg_$rep.jsonl > /dev/null 2> gpurun_out/r2v_c3_ldg.err
FSB_TUNE_STREAM_TMA=1 timeout 300 python tools/bench_all.py --only c3 --out gpurun_out/r2v_c3_tma_$rep.jsonl > /dev/null 2> gpurun_out/r2v_c3_tma.err
done
for f in gpurun_out/r2v_c3_*.jsonl; do echo $f; python - "$f" <<'PY'
import json,sys
for l in list(open(sys.argv[1]))[:4]:
    d=json.loads(l); print("  %-70s %.3f ms" % (d["config"][:70], d["ms"]))
PY
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:csr_stream -c 8 -o gpurun_out/r2v_ncu_stream python tools/prof_kernels.py --only spmv > gpurun_out/r2v_ncu_stream.log 2>&1
ncu -i gpurun_out/r2v_ncu_stream.ncu-rep --page raw --csv > gpurun_out/r2v_ncu_stream_raw.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/r2v_ncu_stream_raw.csv "R = 1 at C3: TMA-fed merge-path stream kernel" > gpurun_out/r2v_ncu_stream.md
cat gpurun_out/r2v_ncu_stream.md | tail -12
