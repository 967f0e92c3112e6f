#!/bin/bash
# quick iteration loop: fused tests, CTA-0 timeline, short bench summary (per cluster size)
for c in ${CLUSTERS:-2}; do
export OSQ_FUSED_CLUSTER=$c
echo "######## cluster=$c"
timeout 600 python -m pytest tests/test_gpu_fused_linear.py -q -m gpu --timeout 300 2>&1 | tail -4
timeout 120 python scripts/trace_fused.py 2>&1 | grep -E "===|ctas=|mma|epi " | cut -c1-420
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l[:300]); continue
    print('value',d['value'],'ms',d['ms_per_step'],'frac',d['roofline']['frac'],'e2e',d['e2e']['value'])
    for k,v in d['roofline']['sites'].items(): print(k,round(v['us'],1),round(v['frac'],3))
"
done
