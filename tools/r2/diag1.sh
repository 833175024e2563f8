#!/bin/bash
# round-2 baseline diagnostics: forward phase split (FNR_DEBUG_FWD), bench line of the round-1 kernels
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2_gpu.txt
FNR_DEBUG_FWD=1 timeout 300 python tools/profile_driver.py small 2 > gpurun_out/r2_fwd_phase.log 2>&1; echo "phase rc=$?"; grep "fwd slot" gpurun_out/r2_fwd_phase.log | tail -4
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu --no-train > gpurun_out/r2_bench0.json 2> gpurun_out/r2_bench0.err; echo "bench rc=$?"
python - <<'PY'
import json
j=json.loads(open('gpurun_out/r2_bench0.json').read())
print({k:round(j[k],4) for k in ('ms_per_step','fwd_ms','bwd_ms')}, j['roofline_forward']['frac'])
PY
