#!/bin/bash
mkdir -p gpurun_out
echo "== pytest backward"
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "backward" --timeout 300 > gpurun_out/pytest_bwd.log 2>&1; echo "rc=$?"; tail -40 gpurun_out/pytest_bwd.log
echo "== host overhead auto"
timeout 300 python tools/host_overhead.py auto > gpurun_out/host_auto.log 2>&1; grep -E "trial|per-step|forward-only|Error|error" gpurun_out/host_auto.log | head
echo "== bench"
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu 2>/dev/null | python -c "import sys,json; j=json.loads(sys.stdin.read()); print('ms/step', round(j['ms_per_step'],3), 'fwd', round(j['fwd_ms'],3), 'bwd', round(j['bwd_ms'],3), 'rays/s', int(j['value']), 'e2e rays/s', int(j['e2e']['value']), j['clocks'])"
