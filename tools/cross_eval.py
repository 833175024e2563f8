"""Cross-evaluation of trained weights: separates "what training produced" from "how it is rendered" (DESIGN.md section 7.1).

    # CPU, no GPU needed: held-out metrics of ANY pipeline state dict (a kernel-trained Trainer checkpoint, or the state
    # tools/oracle_train.py --state wrote) through the ORACLE's evaluation
    python tools/cross_eval.py --ckpt RUN/nerfstudio_models/step-000003000.ckpt --seed 0 --through oracle
    # GPU: the same weights through the KERNELS' evaluation (FruitPipeline.get_eval_image_metrics_and_images)
    python tools/cross_eval.py --ckpt /tmp/oracle_runs/A_state.pt --seed 0 --through kernels

Both print one row per held-out view (PSNR, fruit IoU, mean accumulation).  Kernel-trained weights that look bad through BOTH
paths were trained differently; weights that look good through the oracle and bad through the kernels expose an evaluation bug.
"""
from __future__ import annotations

import argparse
import json
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tools"))

from fruitnerf_b200.scripts.train import synthetic_spec  # noqa: E402


def _load(pipeline, path):
    state = torch.load(path, map_location="cpu", weights_only=False)
    pipeline.load_state_dict(state["pipeline"], strict=True)
    return int(state.get("step", 0))


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--ckpt", required=True)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--through", choices=["oracle", "kernels"], default="oracle")
    ap.add_argument("--near", type=float, default=0.0, help="near plane of the evaluation rays (nerfstudio's collider uses 0 outside training)")
    ap.add_argument("--anneal", type=float, default=1.0, help="proposal-weight annealing exponent at the checkpoint's step (1.0 after proposal_weights_anneal_max_num_iters)")
    ap.add_argument("--threads", type=int, default=0)
    ap.add_argument("--image-size", type=int, default=160)
    ap.add_argument("--num-images", type=int, default=40)
    ap.add_argument("--num-fruits", type=int, default=12)
    a = ap.parse_args(argv)
    if a.threads:
        torch.set_num_threads(a.threads)
    torch.manual_seed(a.seed)
    spec = synthetic_spec("fruit_nerf", a.num_images, a.image_size, a.num_fruits, a.seed)
    rows = []
    if a.through == "oracle":
        import oracle_train as ot

        pipeline = spec.pipeline.setup(device="cpu", test_mode="val")
        step = _load(pipeline, a.ckpt)
        pipeline.eval()
        model, dm, cfg = pipeline.model, pipeline.datamanager, pipeline.model.config
        fp, fspec, pp, ps = ot._param_dicts(model)
        ds = dm.eval_dataset
        for view in range(len(ds)):
            b = ds.cameras.generate_rays(view)
            o, d = b.origins.reshape(-1, 3), b.directions.reshape(-1, 3)
            rgb, sem, acc = [], [], []
            with torch.no_grad():
                for s in range(0, o.shape[0], 6400):
                    oo, dd = o[s:s + 6400].contiguous(), d[s:s + 6400].contiguous()
                    n = oo.shape[0]
                    out = ot._forward(model, fp, fspec, pp, ps, oo, dd, None, torch.full((n, 1), a.near), torch.full((n, 1), cfg.far_plane), False, False,
                                      a.anneal)
                    rgb.append(out["rgb"]), sem.append(out["semantics"]), acc.append(out["accumulation"])
            rgb, sem, acc = torch.cat(rgb), torch.cat(sem), torch.cat(acc)
            image, gt = ds.images[view].reshape(-1, 3), ds.fruit_masks[view].reshape(-1, 1)
            pred = (torch.sigmoid(sem) > 0.9).float()
            inter, union = float((pred * gt).sum()), float(((pred + gt) > 0).float().sum())
            rows.append({"view": view, "psnr": float(-10 * torch.log10(torch.mean((rgb - image) ** 2))), "fruit_iou": inter / union if union else 1.0,
                         "mean_accumulation": float(acc.mean())})
    else:
        from fruitnerf_b200.trainer import Trainer

        trainer = Trainer(spec, device="cuda:0")
        step = _load(trainer.pipeline, a.ckpt)
        trainer.pipeline.model.proposal_sampler.set_anneal(a.anneal)
        for view in range(len(trainer.pipeline.datamanager.eval_dataset)):
            m, _ = trainer.pipeline.get_eval_image_metrics_and_images(step)
            rows.append({"view": m["image_idx"], "psnr": m["psnr"], "fruit_iou": m["fruit_iou"], "ssim": m["ssim"]})
    for r in rows:
        print(json.dumps(r))
    print(json.dumps({"ckpt": a.ckpt, "step": step, "through": a.through, "near": a.near,
                      "mean": {k: sum(r[k] for r in rows) / len(rows) for k in ("psnr", "fruit_iou")}}))


if __name__ == "__main__":
    main()
