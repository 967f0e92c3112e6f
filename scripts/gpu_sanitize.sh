#!/bin/bash
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
S=/usr/local/cuda/bin/compute-sanitizer
for tool in memcheck synccheck; do
  for part in ${SAN_PARTS:-observers mse fq fused f3}; do
    echo "== $tool $part"; timeout 900 $S --tool $tool --error-exitcode 9 python scripts/sanitize_round2.py $part > gpurun_out/san_${tool}_${part}.log 2>&1; echo "rc=$?"
    grep -E "ERROR SUMMARY|ok |Error|error:" gpurun_out/san_${tool}_${part}.log | tail -6
  done
done
for part in ${SAN_RACE_PARTS:-observers f3}; do
  echo "== racecheck $part"; timeout 900 $S --tool racecheck --error-exitcode 9 python scripts/sanitize_round2.py $part > gpurun_out/san_racecheck_${part}.log 2>&1; echo "rc=$?"
  grep -E "RACECHECK SUMMARY|ok |hazard" gpurun_out/san_racecheck_${part}.log | tail -8
done
