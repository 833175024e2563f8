#!/bin/bash
# round-2 ncu evidence: launch lists of the bench step (small / big) + `ncu --set full` of the hot kernels
mkdir -p gpurun_out
T="timeout -s KILL"
N="ncu --set full --clock-control none --import-source on"
L="ncu --metrics gpu__time_duration.sum --clock-control none --csv"
$T 200 $L -s 30 -c 40 --log-file gpurun_out/r2_launches_small_step.csv python tools/profile_driver.py small 4 > /dev/null 2>&1; echo "list small rc=$?"
$T 200 $L -s 40 -c 60 --log-file gpurun_out/r2_launches_big_step.csv python tools/profile_driver.py big 4 > /dev/null 2>&1; echo "list big rc=$?"
$T 240 $N -k regex:tc_render_forward_ws -s 2 -c 1 -o gpurun_out/prof_fwd_ws_r2 -f python tools/profile_driver.py small 3 > gpurun_out/ncu_r2_1.log 2>&1; echo "fwd rc=$?"
$T 240 $N -k regex:tc_field_backward -s 2 -c 1 -o gpurun_out/prof_bwd_r2 -f python tools/profile_driver.py small 3 > gpurun_out/ncu_r2_2.log 2>&1; echo "bwd rc=$?"
$T 240 $N -k regex:tc_big_backward_chain -s 2 -c 1 -o gpurun_out/prof_bigchain_r2 -f python tools/profile_driver.py big 3 > gpurun_out/ncu_r2_3.log 2>&1; echo "bigchain rc=$?"
$T 240 $N -k regex:tc_big_dw -s 2 -c 1 -o gpurun_out/prof_bigdw_r2 -f python tools/profile_driver.py big 3 > gpurun_out/ncu_r2_4.log 2>&1; echo "bigdw rc=$?"
$T 240 $N -k regex:tc_render_forward_big -s 2 -c 1 -o gpurun_out/prof_bigfwd_r2 -f python tools/profile_driver.py big 3 > gpurun_out/ncu_r2_5.log 2>&1; echo "bigfwd rc=$?"
ls -la gpurun_out/*_r2.ncu-rep gpurun_out/r2_launches_*.csv
