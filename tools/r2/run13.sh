#!/bin/bash
# red-throughput microbenchmark + A/B of the small-backward variants (prefetched loads; aggregation depth; single-round merge)
mkdir -p gpurun_out
T="timeout -s KILL"
$T 60 tools/bin/red_probe > gpurun_out/r2_red_probe.log 2>&1; echo "red_probe rc=$?"; cat gpurun_out/r2_red_probe.log | cut -c1-160
for v in default old agg3 agg5 m6 m8 a2m8; do
  if [ "$v" = default ]; then unset FNR_LIB; else export FNR_LIB=$PWD/tools/bin/libfnr_$v.so; fi
  $T 100 python bench.py --steps 20 --warmup 3 --no-cpu --no-train --no-variants > gpurun_out/r2_bench_$v.json 2> gpurun_out/r2_bench_$v.err; rc=$?
  python - <<PY
import json
try:
    j=json.loads(open('gpurun_out/r2_bench_$v.json').read())
    print('small $v rc=$rc', {k:round(j[k],4) for k in ('value','ms_per_step','fwd_ms','bwd_ms')}, 'e2e', int(j['e2e']['value']))
except Exception as e:
    print('$v rc=$rc parse failed', e); print(open('gpurun_out/r2_bench_$v.err').read()[-600:])
PY
done
unset FNR_LIB
$T 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_golden.py -m gpu -x -q --timeout 120 -k "(backward or gradients) and not big" > gpurun_out/r2_pytest_bwd.log 2>&1; echo "default bwd parity rc=$?"; tail -2 gpurun_out/r2_pytest_bwd.log | cut -c1-300
for v in m8 a2m8; do
FNR_LIB=$PWD/tools/bin/libfnr_$v.so $T 200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -x -q --timeout 120 -k "(backward or gradients) and small" > gpurun_out/r2_pytest_$v.log 2>&1; echo "$v parity rc=$?"; tail -2 gpurun_out/r2_pytest_$v.log | cut -c1-300
done
$T 300 python -m pytest tests/test_gpu_training.py -m gpu -x -q --timeout 200 -k "trained_model_export_128 or sampler_training_flag or reduces_loss" > gpurun_out/r2_pytest_new.log 2>&1; echo "new tests rc=$?"; tail -3 gpurun_out/r2_pytest_new.log | cut -c1-400
