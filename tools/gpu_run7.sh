#!/bin/bash
mkdir -p gpurun_out
echo "== pytest gpu full"
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"; grep -E "AssertionError|^FAILED|passed|failed|Error" gpurun_out/pytest_gpu.log | head -20
echo "== bench"
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_r1.json 2> gpurun_out/bench_r1.err; cat gpurun_out/bench_r1.json | python -c "import sys,json; j=json.loads(sys.stdin.read()); print({k:j[k] for k in ('value','ms_per_step','fwd_ms','bwd_ms','e2e','roofline_forward','cpu_baseline','clocks')})"
