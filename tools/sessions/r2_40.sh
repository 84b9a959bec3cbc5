#!/bin/bash
# round 2, GPU session 40 (1 GPU): the final build -- GPU suite, smoke, both bench arms
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2z3_pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/r2z3_pytest_gpu.log
tail -3 gpurun_out/r2z3_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2z3_smoke.log 2>&1; tail -1 gpurun_out/r2z3_smoke.log
timeout 900 python bench.py --impl reference > gpurun_out/r2z3_bench_c2_reference_arm.json 2> gpurun_out/r2z3_bench_ref.err; echo "ref rc=$?"
timeout 900 python bench.py > gpurun_out/r2z3_bench_c2.json 2> gpurun_out/r2z3_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2z3_bench_c2.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step","gpu_launches","clocks")}, d["roofline"]["frac"], d["e2e"]["ms_per_step"], d["e2e"].get("pageable_ms"))
r=json.loads(open("gpurun_out/r2z3_bench_c2_reference_arm.json").read().strip().splitlines()[-1])
print("reference arm", r.get("value"), r.get("ms_per_step"), r.get("cpu_baseline",{}).get("cores"))
PY
