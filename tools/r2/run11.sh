#!/bin/bash
# red-request diet: run-length aggregation depth of the table scatter, small backward (in-chain scatter warps) and big (level-major in the dW kernel)
mkdir -p gpurun_out
T="timeout -s KILL"
for v in default agg6 agg8; do
  if [ "$v" = default ]; then unset FNR_LIB; else export FNR_LIB=$PWD/tools/bin/libfnr_$v.so; fi
  $T 100 python bench.py --steps 20 --warmup 3 --no-cpu --no-train --no-variants > gpurun_out/r2_bench_$v.json 2> gpurun_out/r2_bench_$v.err; rc=$?
  python - <<PY
import json
try:
    j=json.loads(open('gpurun_out/r2_bench_$v.json').read())
    print('small $v rc=$rc', {k:round(j[k],4) for k in ('ms_per_step','fwd_ms','bwd_ms')})
except Exception as e:
    print('$v rc=$rc parse failed', e); print(open('gpurun_out/r2_bench_$v.err').read()[-600:])
PY
done
for v in default dwagg7 dwagg4; do
  if [ "$v" = default ]; then unset FNR_LIB; else export FNR_LIB=$PWD/tools/bin/libfnr_$v.so; fi
  $T 100 python bench.py --variant big --steps 10 --warmup 3 --no-cpu --no-train --no-variants > gpurun_out/r2_bench_big_$v.json 2> gpurun_out/r2_bench_big_$v.err; rc=$?
  python - <<PY
import json
try:
    j=json.loads(open('gpurun_out/r2_bench_big_$v.json').read())
    print('big $v rc=$rc', {k:round(j[k],4) for k in ('ms_per_step','fwd_ms','bwd_ms')})
except Exception as e:
    print('$v rc=$rc parse failed', e); print(open('gpurun_out/r2_bench_big_$v.err').read()[-600:])
PY
done
unset FNR_LIB
$T 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -x -q --timeout 120 -k "backward or gradients" > gpurun_out/r2_pytest_bwd.log 2>&1; echo "bwd parity rc=$?"; tail -3 gpurun_out/r2_pytest_bwd.log | cut -c1-300
FNR_LIB=$PWD/tools/bin/libfnr_agg8.so $T 200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout 120 -k "backward and small" > gpurun_out/r2_pytest_bwd8.log 2>&1; echo "agg8 parity rc=$?"; tail -2 gpurun_out/r2_pytest_bwd8.log | cut -c1-300
