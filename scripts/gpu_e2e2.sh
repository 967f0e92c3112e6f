#!/bin/bash
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
for d in 2 3; do
echo "== bench depth $d"; timeout 400 python bench.py --steps 20 --warmup 3 --no-cpu --no-sweep --e2e-depth $d > gpurun_out/bench_ab.json 2> gpurun_out/bench.err; echo "rc=$?"
python - <<PY
import json
r=json.load(open("gpurun_out/bench_ab.json"))
print("value %.0f ms %.4f frac %.4f" % (r["value"], r["ms_per_step"], r["roofline"]["frac"]))
print(json.dumps({k:v for k,v in r["e2e"].items() if k!="api"}))
PY
tail -3 gpurun_out/bench.err
done
