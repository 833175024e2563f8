"""profiles/r2_ncu_summary.md + profiles/r2_ncu_traffic.json from the round-2 `ncu --set full` captures in gpurun_out/."""
import csv, json, subprocess, sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent.parent
REPS = [("tc_render_forward_ws_kernel", "prof_fwd_ws_r2", "fused warp-specialised forward, fruit_nerf 4096x192 (training forward: writes the encoding stash)"),
        ("tc_field_backward_kernel", "prof_bwd_r2", "tensor-core backward, fruit_nerf 4096x192"),
        ("tc_render_forward_big_kernel", "prof_bigfwd_r2", "fused forward, fruit_nerf_big 4096x192"),
        ("tc_big_backward_chain_kernel", "prof_bigchain_r2", "big-family backward chain (recompute + dX + table scatter + operand tiles to scratch)"),
        ("tc_big_dw_kernel", "prof_bigdw_r2", "big-family weight gradients: TMA-fed tcgen05, accumulators resident in TMEM")]
KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'sm__cycles_elapsed.max', 'smsp__inst_executed.sum', 'launch__registers_per_thread',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum', 'lts__t_requests_srcunit_tex_op_red.sum']


def bytes_of(v, unit):
    f = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
    return float(v.replace(",", "")) * f


out = ["# Round 2 - ncu summaries (B200, `--set full --clock-control none`; numbers under ncu are cold-cache and serialised)", "",
       "Captured with `tools/r2/profile_all.sh` (`ncu ... -k regex:<kernel> -s 2 -c 1 python tools/profile_driver.py <variant> 3`).", ""]
traffic = {}
for kernel, stem, what in REPS:
    rep = ROOT / "gpurun_out" / f"{stem}.ncu-rep"
    if not rep.exists():
        continue
    raw = subprocess.run(["ncu", "-i", str(rep), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, unit, vals = rows[0], rows[1], rows[-1]
    out += [f"## {kernel} ({what})", "", "| metric | value | unit |", "|---|---|---|"]
    rd = wr = 0.0
    for i, h in enumerate(hdr):
        if h in KEYS:
            out.append(f"| `{h}` | {vals[i]} | {unit[i]} |")
        if h == "dram__bytes_read.sum":
            rd = bytes_of(vals[i], unit[i])
        if h == "dram__bytes_write.sum":
            wr = bytes_of(vals[i], unit[i])
    st = {h.replace("smsp__pcsamp_warps_issue_stalled_", ""): float(vals[i]) for i, h in enumerate(hdr)
          if h.startswith("smsp__pcsamp_warps_issue_stalled_") and not h.endswith("_not_issued")}
    tot = sum(st.values()) or 1.0
    out += ["", "Stall samples: " + ", ".join(f"{k} {100 * v / tot:.1f} %" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:7]), ""]
    traffic[kernel] = {"dram_bytes": int(rd + wr), "dram_read": int(rd), "dram_write": int(wr), "capture": f"profiles/r2_ncu_summary.md ({stem}.ncu-rep, tools/r2/profile_all.sh)"}
(ROOT / "profiles" / "r2_ncu_summary.md").write_text("\n".join(out) + "\n")
(ROOT / "profiles" / "r2_ncu_traffic.json").write_text(json.dumps(traffic, indent=1) + "\n")
print("\n".join(out[:8]))
print(json.dumps(traffic, indent=1))
