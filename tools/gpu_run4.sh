#!/bin/bash
mkdir -p gpurun_out
echo "== probe MN-major"
for cfg in "3 64 64" "3 16 64" "3 64 32" "4 64 64" "4 64 16" "4 64 32"; do
  timeout 30 tools/bin/tc_probe $cfg 2>&1 || echo "PROBE $cfg : exit $?"
done
echo "== bench (nvml sampler + flush)"
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu 2>/dev/null | python -c "import sys,json; j=json.loads(sys.stdin.read()); print('ms/step', round(j['ms_per_step'],3), 'fwd', round(j['fwd_ms'],3), 'bwd', round(j['bwd_ms'],3), 'e2e rays/s', int(j['e2e']['value']), j['clocks'])"
