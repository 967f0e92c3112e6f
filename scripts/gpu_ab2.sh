#!/bin/bash
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
run() {
  echo "== bench $*"; env "$@" timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu --no-sweep > gpurun_out/bench_ab.json 2> gpurun_out/bench.err; echo "rc=$?"
  python - <<PY
import json
r=json.load(open("gpurun_out/bench_ab.json"))
print("value %.0f ms %.4f frac %.4f e2e %.0f step-frac %.4f" % (r["value"], r["ms_per_step"], r["roofline"]["frac"], r["e2e"]["value"], 12*812.5e6/(r["ms_per_step"]*1e-3)/6534.5e9))
print({k:(round(x["us"],2), round(x["frac"],3)) for k,x in r["roofline"]["sites"].items()})
PY
  tail -3 gpurun_out/bench.err
}
for v in "$@"; do run $v; done
