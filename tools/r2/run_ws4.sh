#!/bin/bash
mkdir -p gpurun_out
for v in default g8 g16d2; do
  if [ "$v" = default ]; then unset FNR_LIB; else export FNR_LIB=$PWD/tools/bin/libfnr_$v.so; fi
  echo "=== $v"
  for f in 1 15; do FNR_DEBUG_FWD=$f timeout 120 python tools/profile_driver.py small 2 2>&1 | grep "^fwd" | tail -4 | cut -c1-230; done
  timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu --no-train --no-variants > gpurun_out/r2_bench_$v.json 2> gpurun_out/r2_bench_$v.err; echo "bench rc=$?"
  python - <<PY
import json
try:
    j=json.loads(open('gpurun_out/r2_bench_$v.json').read())
    print({k:round(j[k],4) for k in ('ms_per_step','fwd_ms','fwd_loss_ms','bwd_ms')}, round(j['roofline_forward']['frac'],4), j['gpu_launches_per_step'])
except Exception as e:
    print('bench parse failed', e); print(open('gpurun_out/r2_bench_$v.err').read()[-1500:])
PY
done
unset FNR_LIB
timeout 600 python -m pytest tests -m gpu -x -q --timeout 300 > gpurun_out/r2_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_pytest.log | cut -c1-200
timeout 900 python bench.py > gpurun_out/r2_bench_full.json 2> gpurun_out/r2_bench_full.err; echo "full bench rc=$?"; tail -c 600 gpurun_out/r2_bench_full.err
python - <<'PY'
import json
j=json.loads(open('gpurun_out/r2_bench_full.json').read())
print({k:(round(v,4) if isinstance(v,float) else v) for k,v in j.items() if k in ('value','ms_per_step','fwd_ms','bwd_ms','gpu_launches')})
print('big', {k:(round(v,4) if isinstance(v,float) else v) for k,v in j.get('variants',{}).get('big',{}).items() if k in ('value','ms_per_step','fwd_ms','bwd_ms','error')})
print('export', {k:v for k,v in (j.get('export_512') or {}).items() if k in ('ms','counts','error','keys_unique_and_nested')})
print('cpu', j.get('cpu_baseline')); print('train', j.get('train_iteration'))
PY
