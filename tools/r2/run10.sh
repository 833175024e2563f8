#!/bin/bash
mkdir -p gpurun_out
T="timeout -s KILL"
$T 120 python bench.py --variant big --steps 10 --warmup 3 --no-cpu --no-train --no-variants > gpurun_out/r2_bench_big_dw14.json 2> gpurun_out/r2_bench_big_dw14.err; rc=$?
python - <<PY
import json
try:
    j=json.loads(open('gpurun_out/r2_bench_big_dw14.json').read())
    print('big rc=$rc', {k:round(j[k],4) for k in ('value','ms_per_step','fwd_ms','fwd_loss_ms','bwd_ms')}, j['gpu_launches_per_step'])
except Exception as e:
    print('big rc=$rc parse failed', e); print(open('gpurun_out/r2_bench_big_dw14.err').read()[-800:])
PY
$T 700 python -m pytest tests -m gpu -q --timeout 150 -x > gpurun_out/r2_pytest.log 2>&1; echo "pytest rc=$?"; grep -E "^FAILED|^ERROR|passed|failed|Error" gpurun_out/r2_pytest.log | cut -c1-250 | tail -8
ncu --metrics gpu__time_duration.sum --clock-control none --csv -s 40 -c 30 --log-file gpurun_out/r2_launches_big_step.csv python tools/profile_driver.py big 4 > /dev/null 2>&1; echo "list rc=$?"
grep -E "tc_big|big_fold|forward_big" gpurun_out/r2_launches_big_step.csv | awk -F'","' '{print $5, $(NF)}' | cut -c1-120 | tail -4
