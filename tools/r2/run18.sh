#!/bin/bash
# big-family dW kernel: scatter schedules (full scan vs one merge round) + parity of the chosen small backward (default build)
mkdir -p gpurun_out
T="timeout -s KILL"
for v in default D1 D2 D3 D4 default D1; do
  if [ "$v" = default ]; then unset FNR_LIB; else export FNR_LIB=$PWD/tools/bin/libfnr_$v.so; fi
  $T 100 python bench.py --variant big --steps 15 --warmup 3 --no-cpu --no-train --no-variants > gpurun_out/r2_bench_big_$v.json 2> gpurun_out/r2_bench_big_$v.err; rc=$?
  python - <<PY
import json
try:
    j=json.loads(open('gpurun_out/r2_bench_big_$v.json').read())
    print('big $v rc=$rc', {k:round(j[k],4) for k in ('value','ms_per_step','fwd_ms','bwd_ms')}, 'e2e', int(j['e2e']['value']))
except Exception as e:
    print('$v rc=$rc parse failed', e); print(open('gpurun_out/r2_bench_big_$v.err').read()[-600:])
PY
done
unset FNR_LIB
$T 100 python bench.py --steps 30 --warmup 3 --no-cpu --no-train --no-variants > gpurun_out/r2_bench_small_head.json 2> gpurun_out/r2_bench_small_head.err; python -c "
import json; j=json.loads(open('gpurun_out/r2_bench_small_head.json').read()); print('small HEAD', {k:round(j[k],4) for k in ('value','ms_per_step','fwd_ms','bwd_ms')})"
$T 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_golden.py -m gpu -x -q --timeout 120 -k "(backward or gradients) and not big" > gpurun_out/r2_pytest_bwd.log 2>&1; echo "default small bwd parity rc=$?"; tail -2 gpurun_out/r2_pytest_bwd.log | cut -c1-300
for v in D1 D2; do
FNR_LIB=$PWD/tools/bin/libfnr_$v.so $T 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -x -q --timeout 120 -k "(backward or gradients) and big" > gpurun_out/r2_pytest_$v.log 2>&1; echo "$v big parity rc=$?"; tail -2 gpurun_out/r2_pytest_$v.log | cut -c1-300
done
FNR_LIB=$PWD/tools/bin/libfnr_bwdprof.so $T 120 python tools/profile_driver.py small 3 > gpurun_out/r2_bwd_prof_final.log 2>&1; echo "prof rc=$?"; grep "^bwd" gpurun_out/r2_bwd_prof_final.log | tail -3
FNR_DEBUG_BWD=1 FNR_LIB=$PWD/tools/bin/libfnr_bwdprof.so $T 120 python tools/profile_driver.py small 3 > gpurun_out/r2_bwd_prof_final_noscatter.log 2>&1; grep "^bwd" gpurun_out/r2_bwd_prof_final_noscatter.log | tail -2
