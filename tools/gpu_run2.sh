#!/bin/bash
# GPU call 2: tcgen05 fused forward bring-up (tests), host-overhead profile, bench both kernels.
mkdir -p gpurun_out
echo "== pytest gpu (tcgen05 subset first)"
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "tcgen05" --timeout 300 > gpurun_out/pytest_tc.log 2>&1; echo "pytest tc rc=$?"; tail -30 gpurun_out/pytest_tc.log
echo "== host overhead simt"
timeout 300 python tools/host_overhead.py simt > gpurun_out/host_simt.log 2>&1; tail -40 gpurun_out/host_simt.log
echo "== host overhead tcgen05"
timeout 300 python tools/host_overhead.py tcgen05 > gpurun_out/host_tc.log 2>&1; grep -E "trial|per-step|forward-only|Error|error" gpurun_out/host_tc.log | head
echo "== bench auto"
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_auto.json 2> gpurun_out/bench_auto.err; echo "bench rc=$?"; cat gpurun_out/bench_auto.json; tail -3 gpurun_out/bench_auto.err
