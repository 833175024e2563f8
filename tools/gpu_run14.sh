#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tc_field_backward -s 2 -c 1 -o gpurun_out/prof_tc_bwd_r1b -f python tools/profile_driver.py small 3 > gpurun_out/ncu_tcb.log 2>&1; echo "rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tc_render_forward -s 2 -c 1 -o gpurun_out/prof_tc_fwd_r1b -f python tools/profile_driver.py small 3 > gpurun_out/ncu_tcf.log 2>&1; echo "rc=$?"
