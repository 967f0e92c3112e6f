#!/bin/bash
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
echo "== fused tests"; timeout 1200 python -m pytest tests/test_gpu_fused_linear.py -q -m gpu --timeout 600 -x > gpurun_out/test_gpu_fused_linear.log 2>&1; echo "rc=$?"
grep -E "passed|failed|error|^E " gpurun_out/test_gpu_fused_linear.log | tail -12
for v in ${VARIANTS:-"OSQ_FUSED_STORE3D=1" "OSQ_FUSED_STORE3D=0"}; do
echo "== bench $v"; env $v timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu --no-sweep > gpurun_out/bench_$v.json 2> gpurun_out/bench.err; echo "rc=$?"
python - <<PY
import json
r=json.load(open("gpurun_out/bench_$v.json"))
print("value %.0f ms %.4f frac %.4f e2e %.0f" % (r["value"], r["ms_per_step"], r["roofline"]["frac"], r["e2e"]["value"]))
print({k:(round(x["us"],2), round(x["frac"],3)) for k,x in r["roofline"]["sites"].items()})
PY
done
