#!/bin/bash
for f in 0 1 2 3; do
FNR_DEBUG_BWD=$f timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu 2>/dev/null | python -c "import sys,json; j=json.loads(sys.stdin.read()); print('flags $f', {k:round(j[k],3) for k in ('ms_per_step','fwd_ms','bwd_ms')})"
done
