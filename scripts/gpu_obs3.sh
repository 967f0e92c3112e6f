#!/bin/bash
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
echo "== timing carveout 50"; timeout 300 python scripts/time_observer_call.py 2>&1 | grep -E "x|prune_observe_us|token_minmax_only|avg_minmax"
echo "== timing carveout off"; OSQ_OBS_CARVEOUT=-1 timeout 300 python scripts/time_observer_call.py 2>&1 | grep -E "x|prune_observe_us"
echo "== timing carveout 100"; OSQ_OBS_CARVEOUT=100 timeout 300 python scripts/time_observer_call.py 2>&1 | grep -E "x|prune_observe_us"
echo "== ncu durations"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:prune_select_tail|token_minmax" -c 60 --csv --log-file gpurun_out/obs_launches.csv python scripts/time_observer_call.py > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open("gpurun_out/obs_launches.csv")) if len(r)>5]
hdr=rows[0]; ki=hdr.index("Kernel Name"); vi=hdr.index("Metric Value")
for r in rows[1:41]: print(r[ki][:40], r[vi])
PY
