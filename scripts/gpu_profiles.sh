#!/bin/bash
# Round evidence in one gpurun call: bench line (driver's flags), both arms, ncu launch list, ncu --set full of one encoder
# layer's fused launches, ncu --set full of the observer / fake-quant / MSE kernels.
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
echo "== bench ref" ; timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err ; echo "rc=$?"
echo "== bench" ; timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err ; echo "rc=$?"
tail -c 600 gpurun_out/bench.json ; tail -3 gpurun_out/bench.err
echo "== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:fused_fq|pack_weight|minmax|prune_select|fq_per" -c 700 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --only-value --no-graph > gpurun_out/ncu_list.log 2>&1 ; echo "rc=$?"
echo "== ncu full (fused kernel, the 4 launches of one encoder layer: qkv, attn_out, ffn_up, ffn_down)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fused_fq_linear -s 144 -c 4 -o gpurun_out/prof_fused \
    python bench.py --steps 2 --warmup 3 --only-value --no-graph > gpurun_out/ncu_full.log 2>&1 ; echo "rc=$?"
echo "== ncu full (observer / fake-quant kernels at the C2 shape)"
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:token_minmax|prune_select_tail|minmax_masked|fq_per_tensor|abs_hist|mse_multi" -c 14 -o gpurun_out/prof_obs \
    python scripts/profile_small_kernels.py > gpurun_out/ncu_obs.log 2>&1 ; echo "rc=$?"
ls -la gpurun_out | head -40
