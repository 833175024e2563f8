#!/bin/bash
# GPU call 1: smoke, tcgen05 probe matrix, gpu tests, first bench line, ncu launch list.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
echo "== smoke" ; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke.log
echo "== probe"
: > gpurun_out/probe.log
for swap in 0 1; do
  for cfg in "0 64 32" "0 16 64" "0 128 128" "0 64 80" "1 64 32" "1 64 64" "1 128 128" "2 64 32" "2 64 64"; do
    timeout 30 tools/bin/tc_probe $cfg $swap >> gpurun_out/probe.log 2>&1 || echo "PROBE $cfg $swap : exit $?" >> gpurun_out/probe.log
  done
done
cat gpurun_out/probe.log
echo "== pytest gpu"
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu.log
echo "== bench"
timeout 600 python bench.py --steps 10 --warmup 3 --kernel simt > gpurun_out/bench_simt.json 2> gpurun_out/bench_simt.err; echo "bench rc=$?"; cat gpurun_out/bench_simt.json; tail -3 gpurun_out/bench_simt.err
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 2 --warmup 3 --kernel simt --no-cpu > gpurun_out/bench_ncu.log 2>&1; echo "ncu rc=$?"
grep -E "simt_|hash_" gpurun_out/launches_r1.csv | awk -F'","' '{print $5, $NF}' | tail -12
