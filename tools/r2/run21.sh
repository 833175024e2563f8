#!/bin/bash
mkdir -p gpurun_out
T="timeout -s KILL"
for v in O6 O8 O6m4 O6v O8m6 O6m10 O6 O8; do
  export FNR_LIB=$PWD/tools/bin/libfnr_$v.so
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
for v in O8; do
FNR_LIB=$PWD/tools/bin/libfnr_$v.so $T 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_golden.py -m gpu -x -q --timeout 120 -k "(backward or gradients) and not big" > gpurun_out/r2_pytest_$v.log 2>&1; echo "$v parity rc=$?"; tail -2 gpurun_out/r2_pytest_$v.log | cut -c1-300
done
