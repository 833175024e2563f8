#!/bin/bash
# round-1 final profiles: ncu --set full on the hot kernels, launch lists, bench line
mkdir -p gpurun_out
N="ncu --set full --clock-control none --import-source on"
timeout 400 $N -k regex:tc_render_forward -s 2 -c 1 -o gpurun_out/prof_fwd_r1f -f python tools/profile_driver.py small 3 > gpurun_out/ncu1.log 2>&1; echo "fwd rc=$?"
timeout 400 $N -k regex:tc_field_backward -s 2 -c 1 -o gpurun_out/prof_bwd_r1f -f python tools/profile_driver.py small 3 > gpurun_out/ncu2.log 2>&1; echo "bwd rc=$?"
timeout 400 $N -k regex:simt_composite_backward -s 2 -c 1 -o gpurun_out/prof_cbwd_r1f -f python tools/profile_driver.py small 3 > gpurun_out/ncu3.log 2>&1; echo "cbwd rc=$?"
timeout 400 $N -k regex:adam_kernel -s 4 -c 1 -o gpurun_out/prof_adam_r1f -f python tools/train_driver.py 6 > gpurun_out/ncu4.log 2>&1; echo "adam rc=$?"
timeout 400 $N -k regex:proposal_weights_backward -s 2 -c 1 -o gpurun_out/prof_pbwd_r1f -f python tools/train_driver.py 6 > gpurun_out/ncu5.log 2>&1; echo "pbwd rc=$?"
timeout 400 $N -k regex:proposal_weights_forward -s 6 -c 1 -o gpurun_out/prof_pfwd_r1f -f python tools/train_driver.py 6 > gpurun_out/ncu6.log 2>&1; echo "pfwd rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 60 --csv --log-file gpurun_out/launches_bench_r1f.csv python tools/profile_driver.py small 4 > /dev/null 2>&1; echo "list1 rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 1500 -c 700 --csv --log-file gpurun_out/launches_train_r1f.csv python tools/train_driver.py 12 > /dev/null 2>&1; echo "list2 rc=$?"
timeout 600 python bench.py --steps 30 --warmup 5 > gpurun_out/bench_r1f.json 2> gpurun_out/bench_r1f.err; echo "bench rc=$?"; tail -c 1500 gpurun_out/bench_r1f.json
ls -la gpurun_out/*r1f*
