#!/bin/bash
mkdir -p gpurun_out
N="ncu --set full --clock-control none --import-source on"
timeout 400 $N -k regex:tc_render_forward_big -s 2 -c 1 -o gpurun_out/prof_bigfwd_r1f -f python tools/profile_driver.py big 3 > gpurun_out/ncu7.log 2>&1; echo "bigfwd rc=$?"
timeout 400 $N -k regex:tc_big_backward_chain -s 2 -c 1 -o gpurun_out/prof_bigbwd_r1f -f python tools/profile_driver.py big 3 > gpurun_out/ncu8.log 2>&1; echo "bigbwd rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 80 --csv --log-file gpurun_out/launches_big_r1f.csv python tools/profile_driver.py big 4 > /dev/null 2>&1; echo "list rc=$?"
timeout 600 python bench.py --variant big --steps 20 --warmup 5 > gpurun_out/bench_big_r1f.json 2> /dev/null; echo "bench rc=$?"; tail -c 900 gpurun_out/bench_big_r1f.json
