#!/bin/bash
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
echo "== test_gpu_observers" ; timeout 900 python -m pytest tests/test_gpu_observers.py -q -m gpu --timeout 600 -x > gpurun_out/test_gpu_observers.log 2>&1 ; echo "rc=$?"
grep -E "passed|failed|error|^E " gpurun_out/test_gpu_observers.log | tail -12
echo "== observer call timing (4 CTAs/SM)"
timeout 300 python scripts/time_observer_call.py 2>&1 | tail -30
cp gpurun_out/observer_call.json gpurun_out/observer_call_4.json
echo "== observer call timing (2 CTAs/SM)"
OSQ_OBS_CTAS_PER_SM=2 timeout 300 python scripts/time_observer_call.py 2>&1 | tail -30
cp gpurun_out/observer_call.json gpurun_out/observer_call_2.json
