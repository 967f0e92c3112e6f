#!/bin/bash
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
echo "== epi tests"; timeout 900 python -m pytest tests/test_gpu_fused_linear.py -q -m gpu --timeout 600 -k "output_stage" > gpurun_out/test_epi.log 2>&1; echo "rc=$?"
grep -E "passed|failed|error|^E |Mismatch" gpurun_out/test_epi.log | tail -30
echo "== model tests"; timeout 900 python -m pytest tests/test_gpu_reference_model.py tests/test_gpu_model_chain.py -q -m gpu --timeout 600 > gpurun_out/test_model.log 2>&1; echo "rc=$?"
grep -E "passed|failed|error|^E " gpurun_out/test_model.log | tail -20
