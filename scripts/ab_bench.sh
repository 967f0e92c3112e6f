#!/bin/bash
# A/B of one environment knob of the fused kernel: scripts/ab_bench.sh VAR v1 v2 ...   (tests run once, first)
var=$1; shift
timeout 600 python -m pytest tests/test_gpu_fused_linear.py tests/test_gpu_model_chain.py -q -m gpu --timeout 300 2>&1 | tail -3
for v in "$@"; do
  echo "######## $var=$v"
  env $var=$v timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l[:300]); continue
    print('value',round(d['value']),'ms',round(d['ms_per_step'],4),'frac',round(d['roofline']['frac'],4),'e2e',round(d['e2e']['value']))
    print(' '.join('%s %.1f' % (k, v['us']) for k,v in d['roofline']['sites'].items()))
"
done
