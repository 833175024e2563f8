"""Lay a kernel-trained run (python -m fruitnerf_b200.scripts.train --json ...) next to an oracle-trained run
(tools/oracle_train.py --json ...) of the same scene / seed / schedule: markdown table for DESIGN.md section 7.

    python tools/compare_training_runs.py profiles/r2_train_synthetic_seed0.json profiles/r2_oracle_train_seed0.json
"""
import json
import sys


def main(gpu_path, ref_path):
    g, r = json.load(open(gpu_path)), json.load(open(ref_path))
    gh = {row["step"]: row for row in g["history"]}
    rh = {row["step"]: row for row in r["history"]}
    print(f"| | kernels (`{gpu_path}`) | CPU oracle + torch.optim.Adam (`{ref_path}`) |")
    print("|---|---|---|")
    print(f"| steps x rays, schedule | {g['steps']} x {g['rays_per_batch']}, {g['lr_schedule']} | {r['steps']} x {r['rays_per_batch']}, {r['lr_schedule']} |")
    print(f"| wall time | {g['train_seconds']:.1f} s ({g['train_rays_per_s'] / 1e6:.2f} M rays/s) | {r['train_seconds'] / 3600:.2f} h ({r['train_rays_per_s']:.0f} rays/s, {r['threads']} threads) |")
    for step in sorted(set(gh) & set(rh)):
        if step % 500 == 0 or step == max(set(gh) & set(rh)):
            print(f"| train-batch PSNR / loss @ {step} | {gh[step]['psnr']:.2f} dB / {gh[step]['loss']:.5f} | {rh[step]['psnr']:.2f} dB / {rh[step]['loss']:.5f} |")
    print(f"| held-out PSNR (mean appearance) | {g['eval']['psnr']:.2f} dB | {r['eval']['psnr']:.2f} dB |")
    print(f"| held-out fruit IoU | {g['eval']['fruit_iou']:.3f} | {r['eval']['fruit_iou']:.3f} |")
    if "export" in g and "export" in r:
        ge, re_ = g["export"], r["export"]
        for k in ("semantic_colormap", "semantic", "density"):
            print(f"| export cloud `{k}` (points of {ge['export_points']}) | {ge['cloud_sizes'][k]} | {re_['cloud_sizes'][k]} |")
        print(f"| fruit count (truth {ge.get('fruit_count_gt')}) | {ge['fruit_count']} ({ge['fruit_count_before_merge']} before merging) | {re_['fruit_count']} ({re_['fruit_count_before_merge']} before merging) |")
        print(f"| true centres matched / mean centre error | {ge.get('matched_within_radius')} / {ge.get('mean_center_error', float('nan')):.2e} | "
              f"{re_.get('matched_within_radius')} / {re_.get('mean_center_error', float('nan')):.2e} |")
        print(f"| export time | {ge['export_seconds'] * 1e3:.0f} ms | {re_['export_seconds']:.0f} s |")


if __name__ == "__main__":
    main(*sys.argv[1:3])
