#!/bin/bash
# isolate the gather costs of the warp-specialised forward (debug bits) + one ncu --set full capture of it
mkdir -p gpurun_out
for f in 1 3 5 9 7 15; do
  echo "=== FNR_DEBUG_FWD=$f"
  FNR_DEBUG_FWD=$f timeout 120 python tools/profile_driver.py small 2 2>&1 | grep "^fwd" | tail -4
done
timeout 400 ncu --set full --clock-control none --import-source on -k regex:tc_render_forward_ws -s 2 -c 1 -o gpurun_out/prof_fwd_ws_r2a -f python tools/profile_driver.py small 3 > gpurun_out/ncu_ws.log 2>&1; echo "ncu rc=$?"
ls -la gpurun_out/prof_fwd_ws_r2a.ncu-rep
