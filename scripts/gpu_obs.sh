#!/bin/bash
# observers + extra tests + bench (with sweep) + per-call timing of the observer at the C2 shape
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
for f in tests/test_gpu_observers.py tests/test_gpu_extra.py; do
  n=$(basename $f .py)
  echo "== $n" ; timeout 1500 python -m pytest $f -q -m gpu --timeout 900 -x > gpurun_out/$n.log 2>&1 ; echo "rc=$?"
  grep -E "passed|failed|error|^E " gpurun_out/$n.log | tail -12
done
echo "== observer call timing"
timeout 300 python scripts/time_observer_call.py 2>&1 | tail -12
echo "== bench" ; timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err ; echo "rc=$?"
python - <<'PY'
import json
r=json.load(open("gpurun_out/bench.json"))
print(json.dumps({k:r[k] for k in ("value","ms_per_step","e2e","observer_sweep","config4_bart_large")}, indent=1)[:5000])
print(r["roofline"]["frac"])
PY
tail -5 gpurun_out/bench.err
