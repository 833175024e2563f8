#!/bin/bash
mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 80 --csv --log-file gpurun_out/launches_big_r1f.csv python tools/profile_driver.py big 4 > /dev/null 2>&1; echo "list rc=$?"
python - <<'PY'
import csv, collections
rows=list(csv.reader(open('gpurun_out/launches_big_r1f.csv')))
hi=next(i for i,r in enumerate(rows) if 'Kernel Name' in r)
hdr=rows[hi]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
data=[(r[ki], float(r[vi].replace(',',''))) for r in rows[hi+1:] if len(r)>vi and r[vi]]
idx=[i for i,(k,_) in enumerate(data) if 'tc_render_forward_big' in k]
seg=data[idx[-2]:idx[-1]] if len(idx)>=2 else data
for k,v in seg: print(f'{v/1e3:9.1f} us  {k[:110]}')
PY
