#!/bin/bash
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
echo "== bench N=2"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "rc=$?"
python - <<PY
import json
r=json.loads([l for l in open("gpurun_out/bench_n2.json") if l.startswith("{")][-1])
print("value %.0f ms %.4f n_gpus %d" % (r["value"], r["ms_per_step"], r["n_gpus"]))
print(json.dumps({k:v for k,v in r["e2e"].items() if k!="api"}))
print(json.dumps(r["observer_sweep"]))
PY
tail -5 gpurun_out/bench_n2.err
echo "== reference arm under torchrun"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 2>/dev/null | tail -2 | cut -c1-400
