#!/bin/bash
# warp-specialised forward: parity tests, role wait split (FNR_DEBUG_FWD), bench line, A/B against the round-1 kernel
mkdir -p gpurun_out
timeout 120 python __graft_entry__.py smoke > gpurun_out/r2_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2_smoke.log
timeout 600 python -m pytest tests -m gpu -x -q --timeout 300 > gpurun_out/r2_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2_pytest.log | cut -c1-300
FNR_DEBUG_FWD=1 timeout 120 python tools/profile_driver.py small 2 > gpurun_out/r2_fwd_phase.log 2>&1; echo "phase rc=$?"; grep "^fwd" gpurun_out/r2_fwd_phase.log | tail -3
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu --no-train > gpurun_out/r2_bench1.json 2> gpurun_out/r2_bench1.err; echo "bench rc=$?"
python - <<'PY'
import json
j=json.loads(open('gpurun_out/r2_bench1.json').read())
print({k:round(j[k],4) for k in ('ms_per_step','fwd_ms','bwd_ms')}, j['roofline_forward']['frac'])
PY
