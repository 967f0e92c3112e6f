#!/bin/bash
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
echo "== tests"; timeout 1200 python -m pytest tests/test_gpu_fq.py tests/test_gpu_model_chain.py tests/test_gpu_reference_model.py tests/test_gpu_extra.py -q -m gpu --timeout 600 > gpurun_out/test_k1.log 2>&1; echo "rc=$?"
grep -E "passed|failed|error|^E " gpurun_out/test_k1.log | tail -20
timeout 300 python scripts/time_output_stage.py
echo "== bench"; timeout 400 python bench.py --steps 20 --warmup 3 --no-cpu --no-sweep > gpurun_out/bench_ab.json 2> gpurun_out/bench.err; echo "rc=$?"
python - <<PY
import json
r=json.load(open("gpurun_out/bench_ab.json"))
print("value %.0f ms %.4f frac %.4f e2e %.0f host_issue %.2f" % (r["value"], r["ms_per_step"], r["roofline"]["frac"], r["e2e"]["value"], r["e2e"]["host_issue_ms_per_step"]))
PY
