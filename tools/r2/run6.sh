#!/bin/bash
# round-2 checkpoint: big dW kernel parity (short leash), forward depth A/B, whole GPU suite, full bench line
mkdir -p gpurun_out
T="timeout -s KILL"
$T 200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout 90 -k "backward and big" > gpurun_out/r2_pytest_bigbwd.log 2>&1; echo "big backward parity rc=$?"; tail -4 gpurun_out/r2_pytest_bigbwd.log | cut -c1-300
for v in default d3 d4; do
  if [ "$v" = default ]; then unset FNR_LIB; else export FNR_LIB=$PWD/tools/bin/libfnr_$v.so; fi
  FNR_BENCH_DEBUG=1 $T 100 python bench.py --steps 20 --warmup 3 --no-cpu --no-train --no-variants > gpurun_out/r2_bench_$v.json 2> gpurun_out/r2_bench_$v.err; rc=$?
  python - <<PY
import json
try:
    j=json.loads(open('gpurun_out/r2_bench_$v.json').read())
    print('$v rc=$rc', {k:round(j[k],4) for k in ('ms_per_step','fwd_ms','fwd_loss_ms','bwd_ms')}, round(j['roofline_forward']['frac'],4))
except Exception as e:
    print('$v rc=$rc bench parse failed', e); print(open('gpurun_out/r2_bench_$v.err').read()[-800:])
PY
done
unset FNR_LIB
$T 900 python -m pytest tests -m gpu -q --timeout 200 > gpurun_out/r2_pytest.log 2>&1; echo "pytest rc=$?"; grep -E "^FAILED|^ERROR|passed|failed" gpurun_out/r2_pytest.log | cut -c1-250 | tail -15
FNR_BENCH_DEBUG=1 $T 400 python bench.py > gpurun_out/r2_bench_full.json 2> gpurun_out/r2_bench_full.err; echo "full bench rc=$?"; grep "^\[bench" gpurun_out/r2_bench_full.err | tail -3
python - <<'PY'
import json
try:
    j=json.loads(open('gpurun_out/r2_bench_full.json').read())
    print({k:(round(v,4) if isinstance(v,float) else v) for k,v in j.items() if k in ('value','ms_per_step','fwd_ms','bwd_ms','gpu_launches')})
    print('big', {k:(round(v,4) if isinstance(v,float) else v) for k,v in j.get('variants',{}).get('big',{}).items() if k in ('value','ms_per_step','fwd_ms','bwd_ms','error','gpu_launches_per_step')})
    print('export', {k:v for k,v in (j.get('export_512') or {}).items() if k in ('ms','counts','error','keys_unique_and_nested')})
    print('cpu', j.get('cpu_baseline')); print('train', j.get('train_iteration'))
except Exception as e:
    print('full bench parse failed', e); print(open('gpurun_out/r2_bench_full.err').read()[-1500:])
PY
