#!/bin/bash
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
N=${1:-8}
echo "== bench N=$N"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "rc=$?"
python - <<PY
import json
r=json.loads([l for l in open("gpurun_out/bench_n$N.json") if l.startswith("{")][-1])
print("value %.0f ms %.4f n_gpus %d" % (r["value"], r["ms_per_step"], r["n_gpus"]))
print(json.dumps({k:v for k,v in r["e2e"].items() if k!="api"}))
print(json.dumps({k:v for k,v in r["observer_sweep"].items() if k not in ("workload","collective","unit")}))
PY
tail -3 gpurun_out/bench_n$N.err
