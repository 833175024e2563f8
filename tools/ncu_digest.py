"""Digest an .ncu-rep: key metrics, stall mix, hottest SASS lines and hottest source lines."""
import csv, subprocess, sys, collections
rep = sys.argv[1]
out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, unit, vals = rows[0], rows[1], rows[-1]
keys = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'sm__cycles_elapsed.max', 'smsp__inst_executed.sum', 'launch__registers_per_thread',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum', 'lts__t_requests_srcunit_tex_op_red.sum',
        'l1tex__data_pipe_lsu_wavefronts.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'sm__inst_executed_pipe_lsu.sum']
for i, h in enumerate(hdr):
    if h in keys:
        print(f'{h:75s} {vals[i]} {unit[i]}')
st = {h.replace('smsp__pcsamp_warps_issue_stalled_', ''): float(vals[i]) for i, h in enumerate(hdr)
      if h.startswith('smsp__pcsamp_warps_issue_stalled_') and not h.endswith('_not_issued')}
tot = sum(st.values())
print('stalls:', ', '.join(f'{k} {100*v/tot:.1f}%' for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:8]))
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
si, wi = hdr.index('Source'), hdr.index('Warp Stall Sampling (All Samples)')
data = [(int(r[wi] or 0), r[si].strip(), idx) for idx, r in enumerate(rows[2:]) if len(r) > wi]
tot = sum(d[0] for d in data)
print('total samples', tot, 'instructions', len(data))
for s_, src, idx in sorted(data, reverse=True)[:int(sys.argv[2]) if len(sys.argv) > 2 else 14]:
    print(f'{100*s_/tot:5.1f}%  #{idx:5d}  {src[:90]}')
# CUDA-source view
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
try:
    hi = next(i for i, r in enumerate(rows) if 'Source' in r and '#' in r)
    hdr = rows[hi]
    li, si, wi = hdr.index('#'), hdr.index('Source'), hdr.index('Warp Stall Sampling (All Samples)')
    agg = [(int(r[wi] or 0), r[li], r[si].strip()) for r in rows[hi + 1:] if len(r) > wi and r[wi]]
    t2 = sum(a[0] for a in agg) or 1
    print('--- hottest CUDA source lines')
    for s_, ln, src in sorted(agg, reverse=True)[:18]:
        print(f'{100*s_/t2:5.1f}%  L{ln:>5}  {src[:100]}')
except StopIteration:
    pass
