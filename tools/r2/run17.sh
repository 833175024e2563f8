#!/bin/bash
# repeatability of the small-backward scatter schedules (alternating runs)
mkdir -p gpurun_out
T="timeout -s KILL"
for v in default L3 L9 L10 L11 default L3 L11; do
  if [ "$v" = default ]; then unset FNR_LIB; else export FNR_LIB=$PWD/tools/bin/libfnr_$v.so; fi
  $T 100 python bench.py --steps 30 --warmup 3 --no-cpu --no-train --no-variants > gpurun_out/r2_bench_$v.json 2> gpurun_out/r2_bench_$v.err; rc=$?
  python - <<PY
import json
try:
    j=json.loads(open('gpurun_out/r2_bench_$v.json').read())
    print('small $v rc=$rc', {k:round(j[k],4) for k in ('value','ms_per_step','fwd_ms','bwd_ms')}, 'e2e', int(j['e2e']['value']))
except Exception as e:
    print('$v rc=$rc parse failed', e); print(open('gpurun_out/r2_bench_$v.err').read()[-600:])
PY
done
