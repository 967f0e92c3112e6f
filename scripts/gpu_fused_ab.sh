#!/bin/bash
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
run() {
  echo "== bench $*"; env "$@" timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu --no-sweep > gpurun_out/bench_ab.json 2> gpurun_out/bench.err; echo "rc=$?"
  python - <<PY
import json
r=json.load(open("gpurun_out/bench_ab.json"))
print("value %.0f ms %.4f frac %.4f e2e %.0f" % (r["value"], r["ms_per_step"], r["roofline"]["frac"], r["e2e"]["value"]))
print({k:(round(x["us"],2), round(x["frac"],3)) for k,x in r["roofline"]["sites"].items()})
PY
}
for v in "$@"; do run $v; done
