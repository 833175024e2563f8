#!/bin/bash
# A/B of gather variants (coarse LDG + fine cp.async): role wait split + bench line per library build
mkdir -p gpurun_out
for v in default d2 d8 d8cg; do
  if [ "$v" = default ]; then unset FNR_LIB; else export FNR_LIB=$PWD/tools/bin/libfnr_$v.so; fi
  echo "=== $v"
  FNR_DEBUG_FWD=1 timeout 120 python tools/profile_driver.py small 2 2>&1 | grep "^fwd" | tail -4
  timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu --no-train > gpurun_out/r2_bench_$v.json 2> gpurun_out/r2_bench_$v.err; echo "bench rc=$?"
  python - <<PY
import json
j=json.loads(open('gpurun_out/r2_bench_$v.json').read())
print({k:round(j[k],4) for k in ('ms_per_step','fwd_ms','bwd_ms')}, round(j['roofline_forward']['frac'],4))
PY
done
unset FNR_LIB
timeout 600 python -m pytest tests -m gpu -x -q --timeout 300 > gpurun_out/r2_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_pytest.log | cut -c1-200
