#!/bin/bash
# red microbenchmark (conflict patterns, fewer SMs) + role timers of the small backward (prof build)
mkdir -p gpurun_out
T="timeout -s KILL"
$T 60 tools/bin/red_probe > gpurun_out/r2_red_probe.log 2>&1; echo "red_probe rc=$?"; grep -E "pattern [345]|^---|warps/SM 16" gpurun_out/r2_red_probe.log | cut -c1-150
FNR_LIB=$PWD/tools/bin/libfnr_prof.so $T 120 python tools/profile_driver.py small 3 > gpurun_out/r2_bwd_prof.log 2>&1; echo "prof rc=$?"; grep "^bwd\|done" gpurun_out/r2_bwd_prof.log | tail -8
FNR_DEBUG_BWD=1 FNR_LIB=$PWD/tools/bin/libfnr_prof.so $T 120 python tools/profile_driver.py small 3 > gpurun_out/r2_bwd_prof_noscatter.log 2>&1; echo "prof (no scatter) rc=$?"; grep "^bwd\|done" gpurun_out/r2_bwd_prof_noscatter.log | tail -4
FNR_DEBUG_BWD=2 FNR_LIB=$PWD/tools/bin/libfnr_prof.so $T 120 python tools/profile_driver.py small 3 > gpurun_out/r2_bwd_prof_nodw.log 2>&1; echo "prof (no dW) rc=$?"; grep "^bwd\|done" gpurun_out/r2_bwd_prof_nodw.log | tail -4
