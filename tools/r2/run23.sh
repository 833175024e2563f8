#!/bin/bash
# big backward: hybrid placement of the table scatter (coarse levels in the chain kernel's idle scatter warps)
mkdir -p gpurun_out
T="timeout -s KILL"
for cfg in "0 0" "2 0" "4 0" "6 0" "8 0" "4 1" "6 1" "8 1" "0 0"; do
  set -- $cfg
  FNR_BIG_CHAIN_LEVELS=$1 FNR_BIG_CHAIN_MERGE=$2 $T 100 python bench.py --variant big --steps 15 --warmup 3 --no-cpu --no-train --no-variants > gpurun_out/r2_bench_big_h$1_$2.json 2> gpurun_out/r2_bench_big_h$1_$2.err; rc=$?
  python - <<PY
import json
try:
    j=json.loads(open('gpurun_out/r2_bench_big_h$1_$2.json').read())
    print('big chain_levels=$1 merge=$2 rc=$rc', {k:round(j[k],4) for k in ('value','ms_per_step','fwd_ms','bwd_ms')})
except Exception as e:
    print('$cfg rc=$rc parse failed', e); print(open('gpurun_out/r2_bench_big_h$1_$2.err').read()[-600:])
PY
done
for cfg in "6 0" "8 1"; do
  set -- $cfg
  FNR_BIG_CHAIN_LEVELS=$1 FNR_BIG_CHAIN_MERGE=$2 $T 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -x -q --timeout 120 -k "(backward or gradients) and (big or ragged)" > gpurun_out/r2_pytest_h$1_$2.log 2>&1; echo "chain_levels=$1 merge=$2 parity rc=$?"; tail -2 gpurun_out/r2_pytest_h$1_$2.log | cut -c1-300
done
