#!/bin/bash
mkdir -p gpurun_out
echo "== ncu full: tc backward"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tc_field_backward -s 2 -c 1 -o gpurun_out/prof_tc_bwd_r1 -f python tools/profile_driver.py small 3 > gpurun_out/ncu_tcb.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/ncu_tcb.log
echo "== ncu launch list (eager step, auto kernels)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 80 -c 60 --csv --log-file gpurun_out/launches_r1_tc.csv python tools/profile_driver.py small 5 > /dev/null 2>&1; echo "rc=$?"
