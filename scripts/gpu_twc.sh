#!/bin/bash
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
echo "== twc test"; timeout 600 python -m pytest tests/test_gpu_twc.py tests/test_gpu_observers.py -q -m gpu --timeout 500 -x > gpurun_out/test_twc.log 2>&1; echo "rc=$?"
grep -E "passed|failed|error|^E |Error" gpurun_out/test_twc.log | tail -20
