#!/bin/bash
mkdir -p gpurun_out
R="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
timeout -k 10 150 $R --master-port 29511 tools/r2/nvls_test.py > gpurun_out/r2_nvls_test_n4.log 2>&1; echo "nvls test rc=$?"; grep "^nccl\|^nvls" gpurun_out/r2_nvls_test_n4.log | cut -c1-150
timeout -k 10 200 $R --master-port 29512 bench.py --gpus 4 --steps 20 --warmup 3 > gpurun_out/r2_bench_n4_auto.json 2> gpurun_out/r2_bench_n4_auto.err; echo "bench auto rc=$?"
python - <<PY
import json
try:
    j=json.loads(open('gpurun_out/r2_bench_n4_auto.json').read().strip().splitlines()[-1])
    b=j.get('variants',{}).get('big',{})
    print('auto', {k:round(j[k],4) for k in ('value','ms_per_step','comm_ms')}, j.get('exchange',{}).get('kind'), 'big', {k:round(b[k],4) for k in ('value','ms_per_step','comm_ms') if k in b}, b.get('exchange',{}).get('kind'), b.get('error'))
except Exception as e:
    print('auto parse failed', e); print(open('gpurun_out/r2_bench_n4_auto.err').read()[-1500:])
PY
