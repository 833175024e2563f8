#!/bin/bash
# quick regression: tc-related tests + bench line
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "${1:-backward or full_size or render_forward or field_forward}" --timeout 300 > gpurun_out/pytest_sub.log 2>&1; echo "pytest rc=$?"; grep -E "AssertionError|^FAILED|passed|failed|Error" gpurun_out/pytest_sub.log | head
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu 2>/dev/null | python -c "import sys,json; j=json.loads(sys.stdin.read()); print('bench', {k:round(j[k],3) for k in ('ms_per_step','fwd_ms','bwd_ms')}, int(j['value']), int(j['e2e']['value']))"
