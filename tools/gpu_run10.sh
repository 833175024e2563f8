#!/bin/bash
mkdir -p gpurun_out
echo "== pytest backward + full-size"
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "backward or full_size" --timeout 300 > gpurun_out/pytest_bwd.log 2>&1; echo "rc=$?"; grep -E "AssertionError|^FAILED|passed|failed|Error" gpurun_out/pytest_bwd.log | head
echo "== bench"
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/b.json 2> gpurun_out/b.err; echo "rc=$?"
python -c "import sys,json; j=json.loads(open('gpurun_out/b.json').read()); print({k:j[k] for k in ('value','ms_per_step','fwd_ms','bwd_ms','e2e')}); print(j['roofline_forward']['frac'], j['roofline_step']['frac'])"
