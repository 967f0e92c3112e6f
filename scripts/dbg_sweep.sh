#!/bin/bash
for d in 0 8 16 32 48 56; do
  echo "## OSQ_FUSED_DBG=$d (8=no fence 16=no reload 32=no quant math)"
  OSQ_FUSED_DBG=$d TRACE_SHAPES="((768,768),)" timeout 120 python scripts/trace_fused.py 2>&1 | grep -E "conv|ctas=" | cut -c1-200
done
