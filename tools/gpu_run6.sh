#!/bin/bash
mkdir -p gpurun_out
echo "== debug bwd"
timeout 300 python tools/debug_bwd.py > gpurun_out/debug_bwd.log 2>&1; echo "rc=$?"; cat gpurun_out/debug_bwd.log | tail -80
echo "== pytest backward"
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "backward" --timeout 300 > gpurun_out/pytest_bwd.log 2>&1; echo "rc=$?"; grep -E "AssertionError: grad|^FAILED|passed|failed" gpurun_out/pytest_bwd.log | head
