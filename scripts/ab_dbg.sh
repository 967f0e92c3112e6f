for v in 0 128 256 384 386; do
  echo "######## DBG=$v"
  OSQ_FUSED_DBG=$v timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu 2>&1 | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print(' '.join('%s %.1f' % (k, v['us']) for k,v in d['roofline']['sites'].items()))
"
done
