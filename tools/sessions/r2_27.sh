#!/bin/bash
# round 2, GPU session 27 (1 GPU): validation of the TMA-fed stream kernel as shipped -- full GPU test suite, smoke, the
# bench line (both arms), C3 table, ncu capture of the R = 1 kernels
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2x_pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/r2x_pytest_gpu.log
tail -3 gpurun_out/r2x_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2x_smoke.log 2>&1; tail -1 gpurun_out/r2x_smoke.log
timeout 300 python tools/bench_all.py --only c3 --out gpurun_out/r2x_bench_all_c3.jsonl > /dev/null 2> gpurun_out/r2x_bench_all_c3.err
python - <<'PY'
import json
for l in open("gpurun_out/r2x_bench_all_c3.jsonl"):
    d=json.loads(l); print("  %-70s %.3f ms" % (d["config"][:70], d["ms"]))
PY
timeout 900 python bench.py --impl reference > gpurun_out/r2x_bench_c2_reference_arm.json 2> gpurun_out/r2x_bench_ref.err; echo "ref rc=$?"
timeout 900 python bench.py > gpurun_out/r2x_bench_c2.json 2> gpurun_out/r2x_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2x_bench_c2.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step","gpu_launches","clocks")}, d["roofline"]["frac"], d["e2e"]["ms_per_step"], d["e2e"].get("pageable_ms"))
r=json.loads(open("gpurun_out/r2x_bench_c2_reference_arm.json").read().strip().splitlines()[-1])
print("reference arm", r.get("value"), r.get("ms_per_step"), r.get("cpu_baseline",{}).get("cores"))
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"csr_stream|fold_blocks" -c 10 -o gpurun_out/r2x_ncu_stream python tools/prof_kernels.py --only spmv > gpurun_out/r2x_ncu_stream.log 2>&1
ncu -i gpurun_out/r2x_ncu_stream.ncu-rep --page raw --csv > gpurun_out/r2x_ncu_stream_raw.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/r2x_ncu_stream_raw.csv "R = 1 at C3: TMA-fed merge-path stream kernel (final build, 6 CTAs per SM): double SpMV, forced-stream runs, x-blocked transposed SpMV" > gpurun_out/r2x_ncu_stream.md
tail -12 gpurun_out/r2x_ncu_stream.md
