#!/bin/bash
for f in 0 1 4 5; do
  echo "== FNR_DEBUG_BWD=$f"
  FNR_DEBUG_BWD=$f timeout 300 python bench.py --variant big --steps 10 --warmup 3 --no-cpu 2>/dev/null | python -c "import sys,json; j=json.loads(sys.stdin.read()); print({k:round(j[k],4) for k in ('ms_per_step','fwd_ms','bwd_ms')})"
done
