#!/bin/bash
# Round-end evidence in one gpurun call: the driver's own GPU tier (pytest -m gpu, smoke), the bench line, the ncu launch list and
# full captures (fused kernel: one encoder layer; f3 kernels K7 / K8 / K9; observer / fake-quant kernels).
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
echo "== smoke" ; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1 ; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
echo "== pytest -m gpu" ; timeout 1500 python -m pytest tests/ -x -q -m gpu --timeout 600 > gpurun_out/test_gpu_all.log 2>&1 ; echo "rc=$?"; tail -4 gpurun_out/test_gpu_all.log
echo "== bench" ; timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err ; echo "rc=$?"; tail -c 600 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
echo "== bench reference arm" ; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2>> gpurun_out/bench.err ; echo "rc=$?"; tail -c 400 gpurun_out/bench_ref.json
if [ "${SKIP_NCU:-0}" != "1" ]; then
echo "== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:fused_fq|pack_weight|minmax|prune_select|fq_per|layernorm|attn_" -c 700 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --only-value --no-graph > gpurun_out/ncu_list.log 2>&1 ; echo "rc=$?"
echo "== ncu full (fused kernel, one encoder layer)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fused_fq_linear -s 144 -c 4 -o gpurun_out/prof_fused \
    python bench.py --steps 2 --warmup 3 --only-value --no-graph > gpurun_out/ncu_full.log 2>&1 ; echo "rc=$?"
echo "== ncu full (K7 / K8 / K9)"
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:layernorm_fq|attn_scores|attn_context" -c 6 -o gpurun_out/prof_f3 \
    python scripts/profile_f3_kernels.py > gpurun_out/ncu_f3.log 2>&1 ; echo "rc=$?"
fi
ls -la gpurun_out | head -40
