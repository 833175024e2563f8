#!/bin/bash
# round-2 evidence: ncu captures of the restructured big backward + export, launch lists, training runs, the bench lines
mkdir -p gpurun_out
T="timeout -s KILL"
N="ncu --set full --clock-control none --import-source on"
L="ncu --metrics gpu__time_duration.sum --clock-control none --csv"
$T 200 $L -s 30 -c 40 --log-file gpurun_out/r2_launches_small_step.csv python tools/profile_driver.py small 4 > /dev/null 2>&1; echo "list small rc=$?"
$T 200 $L -s 40 -c 60 --log-file gpurun_out/r2_launches_big_step.csv python tools/profile_driver.py big 4 > /dev/null 2>&1; echo "list big rc=$?"
$T 300 $L -s 1500 -c 400 --log-file gpurun_out/r2_launches_train_iteration.csv python tools/train_driver.py 12 > /dev/null 2>&1; echo "list train rc=$?"
$T 240 $N -k regex:tc_big_backward_chain -s 2 -c 1 -o gpurun_out/prof_bigchain_r2 -f python tools/profile_driver.py big 3 > gpurun_out/ncu_r2_3.log 2>&1; echo "bigchain rc=$?"
$T 240 $N -k regex:tc_big_dw -s 2 -c 1 -o gpurun_out/prof_bigdw_r2 -f python tools/profile_driver.py big 3 > gpurun_out/ncu_r2_4.log 2>&1; echo "bigdw rc=$?"
for seed in 0 1 2; do
  $T 150 python -m fruitnerf_b200.scripts.train --steps 3000 --seed $seed --json gpurun_out/r2_train_synthetic_seed$seed.json > gpurun_out/r2_train_seed$seed.log 2>&1; echo "train seed $seed rc=$?"; tail -c 700 gpurun_out/r2_train_seed$seed.log | cut -c1-700
done
$T 420 python bench.py > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; echo "bench rc=$?"
$T 150 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_reference.json 2> gpurun_out/r2_bench_reference.err; echo "ref rc=$?"
python - <<'PY'
import json
j=json.loads(open('gpurun_out/r2_bench_n1.json').read())
print({k:(round(v,4) if isinstance(v,float) else v) for k,v in j.items() if k in ('value','ms_per_step','fwd_ms','bwd_ms','gpu_launches')}, round(j['roofline_forward']['frac'],4), round(j['roofline']['frac'],4), int(j['e2e']['value']))
print('big', {k:(round(v,4) if isinstance(v,float) else v) for k,v in j.get('variants',{}).get('big',{}).items() if k in ('value','ms_per_step','fwd_ms','bwd_ms','error')})
print('export', {k:v for k,v in (j.get('export_512') or {}).items() if k in ('ms','counts','error')})
print('cpu', j.get('cpu_baseline',{}).get('value'), 'train', j.get('train_iteration',{}).get('ms_per_iteration'))
PY
