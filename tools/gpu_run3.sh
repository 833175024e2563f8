#!/bin/bash
mkdir -p gpurun_out
echo "== bench timing variants"
for flags in "--no-flush --no-clocks" "--no-clocks" "--no-flush" ""; do
  timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu $flags 2>/dev/null | python -c "import sys,json; j=json.loads(sys.stdin.read()); print('$flags', '| ms/step', round(j['ms_per_step'],3), 'fwd', round(j['fwd_ms'],3), 'bwd', round(j['bwd_ms'],3), 'e2e rays/s', int(j['e2e']['value']), j['clocks'])"
done
echo "== ncu full: tc forward"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_render_forward -s 2 -c 1 -o gpurun_out/prof_tc_fwd_r1 -f python tools/profile_driver.py small 3 > gpurun_out/ncu_tc.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/ncu_tc.log
echo "== ncu full: simt backward"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:simt_field_backward -s 2 -c 1 -o gpurun_out/prof_simt_bwd_r1 -f python tools/profile_driver.py small 3 > gpurun_out/ncu_bwd.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/ncu_bwd.log
echo "== ncu launch list (auto kernel)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 70 --csv --log-file gpurun_out/launches_r1.csv python tools/profile_driver.py small 4 > /dev/null 2>&1; echo "rc=$?"
ls -la gpurun_out/*.ncu-rep
