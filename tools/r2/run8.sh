#!/bin/bash
mkdir -p gpurun_out
T="timeout -s KILL"
$T 700 python -m pytest tests -m gpu -q --timeout 150 -x > gpurun_out/r2_pytest.log 2>&1; echo "pytest rc=$?"; grep -E "^FAILED|^ERROR|passed|failed|Error" gpurun_out/r2_pytest.log | cut -c1-250 | tail -12
FNR_BENCH_DEBUG=1 FNR_BENCH_WATCHDOG=200 $T 420 python bench.py > gpurun_out/r2_bench_full.json 2> gpurun_out/r2_bench_full.err; echo "full bench rc=$?"; grep "^\[bench" gpurun_out/r2_bench_full.err | tail -2
python - <<'PY'
import json
try:
    j=json.loads(open('gpurun_out/r2_bench_full.json').read())
    print({k:(round(v,4) if isinstance(v,float) else v) for k,v in j.items() if k in ('value','ms_per_step','fwd_ms','bwd_ms','gpu_launches')}, round(j['roofline_forward']['frac'],4), round(j['roofline']['frac'],4))
    print('big', {k:(round(v,4) if isinstance(v,float) else v) for k,v in j.get('variants',{}).get('big',{}).items() if k in ('value','ms_per_step','fwd_ms','bwd_ms','error','gpu_launches_per_step')})
    print('export', {k:v for k,v in (j.get('export_512') or {}).items() if k in ('ms','counts','error','keys_unique_and_nested')})
    print('cpu', j.get('cpu_baseline')); print('train', j.get('train_iteration'))
except Exception as e:
    print('full bench parse failed', e); print(open('gpurun_out/r2_bench_full.err').read()[-1500:])
PY
$T 120 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_bench_ref.json 2> gpurun_out/r2_bench_ref.err; echo "ref rc=$?"; head -c 700 gpurun_out/r2_bench_ref.json
