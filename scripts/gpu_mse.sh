#!/bin/bash
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
echo "== tests"; timeout 600 python -m pytest tests/test_gpu_observers.py tests/test_gpu_extra.py -q -m gpu --timeout 300 > gpurun_out/test_mse.log 2>&1; echo "rc=$?"
grep -E "passed|failed|error|^E " gpurun_out/test_mse.log | tail -20
timeout 300 python scripts/bench_mse_tensor.py 2>&1 | tail -30
