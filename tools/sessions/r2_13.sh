#!/bin/bash
# round 2, GPU session 13 (1 GPU): device sort of host blocked structures + the full suite on the final code
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2m_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r2m_pytest.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2m_smoke.log 2>&1; echo "rc=$?" >> gpurun_out/r2m_smoke.log
echo done
