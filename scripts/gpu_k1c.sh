#!/bin/bash
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
echo "== tests"; timeout 900 python -m pytest tests/test_gpu_fq.py tests/test_gpu_reference_model.py tests/test_gpu_model_chain.py -q -m gpu --timeout 600 > gpurun_out/test_k1c.log 2>&1; echo "rc=$?"
grep -E "passed|failed|error|^E " gpurun_out/test_k1c.log | tail -20
timeout 300 python scripts/time_output_stage.py
bash scripts/gpu_e2e.sh
