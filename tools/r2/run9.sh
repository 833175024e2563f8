#!/bin/bash
# big-family backward with the level-major table scatter fused into the dW kernel: parity, then A/B against the in-chain scatter
mkdir -p gpurun_out
T="timeout -s KILL"
$T 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_golden.py -m gpu -x -q --timeout 120 -k "big or golden" > gpurun_out/r2_pytest_big.log 2>&1; echo "big parity rc=$?"; tail -4 gpurun_out/r2_pytest_big.log | cut -c1-300
for mode in 1 0; do
  FNR_BIG_BWD_MODE=$mode $T 120 python bench.py --variant big --steps 10 --warmup 3 --no-cpu --no-train --no-variants > gpurun_out/r2_bench_big_mode$mode.json 2> gpurun_out/r2_bench_big_mode$mode.err; rc=$?
  python - <<PY
import json
try:
    j=json.loads(open('gpurun_out/r2_bench_big_mode$mode.json').read())
    print('mode $mode rc=$rc', {k:round(j[k],4) for k in ('value','ms_per_step','fwd_ms','fwd_loss_ms','bwd_ms')}, j['gpu_launches_per_step'])
except Exception as e:
    print('mode $mode rc=$rc parse failed', e); print(open('gpurun_out/r2_bench_big_mode$mode.err').read()[-800:])
PY
done
ncu --metrics gpu__time_duration.sum --clock-control none --csv -s 40 -c 60 --log-file gpurun_out/r2_launches_big_step.csv python tools/profile_driver.py big 4 > /dev/null 2>&1; echo "list rc=$?"
grep -E "tc_big|big_fold|forward_big" gpurun_out/r2_launches_big_step.csv | awk -F'","' '{print $5, $(NF)}' | cut -c1-120 | tail -8
