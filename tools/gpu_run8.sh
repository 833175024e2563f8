#!/bin/bash
mkdir -p gpurun_out
for flags in "" "--no-graph" "--no-flush"; do
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu $flags > gpurun_out/b.json 2> gpurun_out/b.err; echo "flags[$flags] rc=$?"; tail -2 gpurun_out/b.err | cut -c1-300
python -c "import sys,json; j=json.loads(open('gpurun_out/b.json').read()); print({k:j[k] for k in ('value','ms_per_step','fwd_ms','bwd_ms','e2e','clocks')}); print(j['roofline_forward']['frac'], j['roofline_step']['frac'])"
done
