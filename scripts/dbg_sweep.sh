#!/bin/bash
for d in 0 1 2 4 3 5 6 7; do
  echo "## OSQ_FUSED_DBG=$d (1=W pinned 2=no Y stores 4=A rows pinned)"
  OSQ_FUSED_DBG=$d OSQ_FUSED_CLUSTER=1 timeout 120 python scripts/trace_fused.py 2>&1 | grep -E "===|ctas=" | sed -n 'N;s/\n/ /;p' | cut -c1-200
done
