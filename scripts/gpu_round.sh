#!/bin/bash
# One gpurun call: smoke, GPU parity tests, bench, ncu launch list + full capture of the fused kernel.
# Every stage has its own timeout so a hung kernel cannot eat the box lease.
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
echo "== smoke" ; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1 ; echo "smoke rc=$?"
tail -3 gpurun_out/smoke.log
echo "== tests fq/observers" ; timeout 600 python -m pytest tests/test_gpu_fq.py tests/test_gpu_observers.py -q -m gpu -x --timeout 300 > gpurun_out/test_fq_obs.log 2>&1 ; echo "rc=$?"
tail -15 gpurun_out/test_fq_obs.log
echo "== tests fused" ; timeout 900 python -m pytest tests/test_gpu_fused_linear.py -q -m gpu --timeout 300 > gpurun_out/test_fused.log 2>&1 ; echo "rc=$?"
tail -25 gpurun_out/test_fused.log
if [ "${SKIP_BENCH:-0}" != "1" ]; then
echo "== bench" ; timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err ; echo "rc=$?"
tail -c 3000 gpurun_out/bench.json ; tail -5 gpurun_out/bench.err
fi
if [ "${SKIP_NCU:-0}" != "1" ]; then
echo "== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:fused_fq|pack_weight|minmax|prune_select|fq_per" -c 700 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --only-value --no-graph > gpurun_out/ncu_list.log 2>&1 ; echo "rc=$?"
echo "== ncu full (fused kernel, the 4 launches of one encoder layer: qkv, attn_out, ffn_up, ffn_down)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fused_fq_linear -s 144 -c 4 -o gpurun_out/prof_fused \
    python bench.py --steps 2 --warmup 3 --only-value --no-graph > gpurun_out/ncu_full.log 2>&1 ; echo "rc=$?"
fi
ls -la gpurun_out | head -30
