#!/bin/bash
# One gpurun call: smoke + the whole -m gpu suite (per-file logs, no -x) + both bench arms.
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
echo "== smoke" ; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1 ; echo "smoke rc=$?"
tail -2 gpurun_out/smoke.log
for f in ${TEST_FILES:-tests/test_gpu_fq.py tests/test_gpu_observers.py tests/test_gpu_extra.py tests/test_gpu_model_chain.py tests/test_gpu_reference_model.py tests/test_gpu_fused_linear.py}; do
  n=$(basename $f .py)
  echo "== $n" ; timeout 1500 python -m pytest $f -q -m gpu --timeout 600 -s > gpurun_out/$n.log 2>&1 ; echo "rc=$?"
  grep -E "passed|failed|error|replayed" gpurun_out/$n.log | tail -8
done
if [ "${SKIP_BENCH:-0}" != "1" ]; then
echo "== bench ref" ; timeout 900 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err ; echo "rc=$?"
tail -c 1500 gpurun_out/bench_ref.json
echo "== bench" ; timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err ; echo "rc=$?"
tail -c 4000 gpurun_out/bench.json ; tail -5 gpurun_out/bench.err
fi
