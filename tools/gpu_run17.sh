#!/bin/bash
for lib in "" tools/bin/libfnr_agg6.so tools/bin/libfnr_agg8.so; do
  echo "== lib=${lib:-default(agg4)}"
  FNR_LIB=${lib:+$PWD/$lib} timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu 2>/dev/null | python -c "import sys,json; j=json.loads(sys.stdin.read()); print({k:round(j[k],4) for k in ('ms_per_step','fwd_ms','bwd_ms')}, int(j['value']))"
done
