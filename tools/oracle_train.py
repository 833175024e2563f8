"""Train the CPU ORACLE (oracle/: the restatement of the reference's nerfstudio torch path) on the synthetic apple
scene with stock ``torch.optim.Adam`` -- the "reference side" of BASELINE.json's last clause (rendered PSNR and
exported fruit count matching the reference on the synthetic apple scene), produced the only way this environment
allows: neither nerfstudio nor the GPU are needed, everything is torch autograd in fp32 on the host cores.

The run mirrors ``python -m fruitnerf_b200.scripts.train`` step for step (same scene, same method config, same
schedules, same update / anneal callbacks, same evaluation, same export + clustering) with the kernels replaced by the
oracle's functions and FusedAdam by torch.optim.Adam, so its JSON can be laid next to ``profiles/r2_train_synthetic_seed*.json``.
It is test / evidence infrastructure: nothing in the product imports it.

    python tools/oracle_train.py --steps 3000 --seed 0 --json profiles/r2_oracle_train_seed0.json
    python tools/oracle_train.py --steps 3000 --stock-schedule ...   # the stability experiment of DESIGN.md section 7

Parameters are initialised by building the SAME FruitPipeline on the CPU (the parameter holders construct without a
GPU), so the initial state equals the GPU run's for the same seed; ray batches come from the data manager's CPU path.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

from fruitnerf_b200.clustering import count_fruits  # noqa: E402
from fruitnerf_b200.optim import ExponentialDecay  # noqa: E402
from fruitnerf_b200.scripts.train import synthetic_spec  # noqa: E402
from oracle import fruit_ref as fr  # noqa: E402
from oracle import ns_torch as ns  # noqa: E402


def _param_dicts(model):
    cfg = model.config
    fparams = dict(model.field.state_dict(keep_vars=True))  # nn.Parameters (autograd leaves) + buffers, reference names
    pparams, pspecs = [], []
    for i, net in enumerate(model.proposal_networks):
        args = cfg.proposal_net_args_list[min(i, len(cfg.proposal_net_args_list) - 1)]
        pparams.append(dict(net.state_dict(keep_vars=True)))
        pspecs.append(fr.DensitySpec(num_levels=args["num_levels"], max_res=args["max_res"], log2_hashmap_size=args["log2_hashmap_size"]))
    spec = fr.FieldSpec(max_res=cfg.max_res, log2_hashmap_size=cfg.log2_hashmap_size, geo_feat_dim=cfg.geo_feat_dim)
    return fparams, spec, pparams, pspecs


def _forward(model, fparams, spec, pparams, pspecs, o, d, cam, nears, fars, training: bool, updated: bool, anneal: float):
    """FruitModel.get_outputs (fruit_nerf.py:316-357) on the oracle."""
    cfg = model.config
    R = o.shape[0]
    t0 = torch.rand(R, 1) if training else None  # UniformLinDispPiecewiseSampler, single_jitter
    us = [torch.rand(R, 1) for _ in pparams] if training else None  # PDFSampler, single_jitter
    with torch.set_grad_enabled(training and updated):
        starts, ends, bins, wl, sl = fr.proposal_sampler(pparams, pspecs, o, d, nears, fars, tuple(cfg.num_proposal_samples_per_ray),
                                                         cfg.num_nerf_samples_per_ray, fparams["aabb"], t_rand0=t0, u_rands=us, anneal=anneal)
    with torch.set_grad_enabled(training):
        f = fr.field_forward(fparams, spec, o[:, None, :], d[:, None, :], starts[..., None], ends[..., None], cam, True,
                             "train" if training else "mean")
        out = fr.render(f, starts[..., None], ends[..., None], training=training)
    out["weights_list"] = list(wl) + [out["weights"][..., 0]]
    out["sdist_list"] = list(sl) + [bins]
    return out


def evaluate(model, fparams, spec, pparams, pspecs, dm, anneal: float, chunk: int = 8192):
    """FruitPipeline.get_average_eval_image_metrics: held-out views, eval-mode sampler, mean appearance, rays from the
    camera centre (NearFarCollider in eval mode), rgb clamped."""
    cfg = model.config
    ds = dm.eval_dataset
    rows = []
    for i in range(len(ds)):
        b = ds.cameras.generate_rays(i)
        o, d = b.origins.reshape(-1, 3), b.directions.reshape(-1, 3)
        rgb, sem = [], []
        with torch.no_grad():
            for a in range(0, o.shape[0], chunk):
                oo, dd = o[a:a + chunk].contiguous(), d[a:a + chunk].contiguous()
                n = oo.shape[0]
                out = _forward(model, fparams, spec, pparams, pspecs, oo, dd, None, torch.zeros(n, 1), torch.full((n, 1), cfg.far_plane), False,
                               False, anneal)
                rgb.append(out["rgb"])
                sem.append(out["semantics"])
        rgb, sem = torch.cat(rgb), torch.cat(sem)
        image, gt = ds.images[i].reshape(-1, 3), ds.fruit_masks[i].reshape(-1, 1)
        mse = torch.mean((rgb - image) ** 2)
        pred = (torch.sigmoid(sem) > 0.9).float()
        inter, union = float((pred * gt).sum()), float(((pred + gt) > 0).float().sum())
        rows.append({"psnr": float(-10.0 * torch.log10(mse)), "fruit_iou": inter / union if union > 0 else 1.0})
    return {k: float(sum(r[k] for r in rows) / len(rows)) for k in ("psnr", "fruit_iou")}, rows


def export_and_count(fparams, spec, dm, points_per_side: int = 256, half_extent: float = 0.3, rays_per_batch: int = 4096, jitter: bool = True):
    """scripts/train.py:export_and_count on the oracle: uniform volume (ns.surface_points / orthographic rays), the sampler's
    per-sample jitter (the state the reference exporter runs in), thresholds 3 / 70 / 0.9, scale(2 / dataparser_scale), DBSCAN."""
    lo, hi = (-half_extent,) * 3, (half_extent,) * 3
    pts, plane = ns.surface_points((lo, hi), points_per_side)
    n_rays = pts.shape[0]
    S = points_per_side
    clouds = {k: [] for k in ("semantic_colormap", "semantic", "density")}
    t0 = time.time()
    with torch.no_grad():
        for count in range(1, (n_rays + rays_per_batch - 1) // rays_per_batch + 1):
            o, dirs, nears, fars = ns.orthographic_rays(pts, plane, batch=rays_per_batch, count=count)
            if o.shape[0] == 0:
                break
            t_rand = torch.rand(o.shape[0], S + 1) if jitter else None
            out = fr.export_outputs(fparams, spec, o, dirs, nears, fars, S, chunk=1 << 17, t_rand=t_rand)
            sel = fr.export_select(out)
            for k in clouds:
                clouds[k].append(sel[k]["points"])
    export_s = time.time() - t0
    scale = 2.0 / float(dm.train_dataset.dataparser_scale)
    clouds = {k: torch.cat(v).double().numpy() * scale for k, v in clouds.items()}
    h = 2.0 * (2 * half_extent) / (points_per_side - 1)
    geom = dm.train_dataset.geometry
    res = count_fruits(clouds["semantic_colormap"], eps=2.5 * h, min_samples=8, cluster_merge_distance=geom.fruit_radius if geom is not None else 0.04)
    out = {"export_seconds": export_s, "export_points": int(n_rays * S), "cloud_sizes": {k: int(v.shape[0]) for k, v in clouds.items()},
           "fruit_count": res["count"], "fruit_count_before_merge": res["count_before_merge"]}
    if geom is not None:
        gt = geom.fruit_centers.numpy()
        out["fruit_count_gt"] = int(gt.shape[0])
        if res["count"]:
            dist = np.linalg.norm(res["centers"][:, None, :] - gt[None], axis=-1)
            out["matched_within_radius"] = int((dist.min(axis=0) < 1.5 * geom.fruit_radius).sum())
            out["mean_center_error"] = float(dist.min(axis=0).mean())
    return out


def train(steps: int, seed: int, short_schedule: bool, log_every: int, method: str = "fruit_nerf", image_size: int = 160, num_images: int = 40,
          num_fruits: int = 12, rays_per_batch=None, points_per_side: int = 256, state_path=None, do_export: bool = True, log=print, stream: int = 0):
    torch.manual_seed(seed)
    tspec = synthetic_spec(method, num_images, image_size, num_fruits, seed, rays_per_batch=rays_per_batch,
                           schedule_steps=steps if short_schedule else None)
    pipeline = tspec.pipeline.setup(device="cpu", test_mode="val")
    pipeline.train()
    model, dm, cfg = pipeline.model, pipeline.datamanager, pipeline.model.config
    if stream:
        # same scene and same initial parameters, a different stream of pixel batches and sampler jitter: run-to-run variance
        torch.manual_seed(seed * 1000003 + stream)
        dm._generator().manual_seed(seed * 7919 + 104729 * stream)
    fparams, spec, pparams, pspecs = _param_dicts(model)
    groups = model.get_param_groups()
    opts, scheds = {}, {}
    for name, params in groups.items():
        oc, sc = tspec.optimizers[name]["optimizer"], tspec.optimizers[name]["scheduler"]
        assert oc["type"] == "Adam", "this tool restates the fruit_nerf method (Adam on both groups)"
        opts[name] = torch.optim.Adam(params, lr=oc["lr"], eps=oc["eps"])
        scheds[name] = ExponentialDecay(oc["lr"], sc["lr_final"] if sc else None, sc["max_steps"] if sc else None)
    R = dm.config.train_num_rays_per_batch

    def update_sched(step):  # fruit_nerf.py:131-136
        return np.clip(np.interp(step, [0, cfg.proposal_warmup], [0, cfg.proposal_update_every]), 1, cfg.proposal_update_every)

    sampler_step, since_update = 0, 0  # ProposalNetworkSampler._step / _steps_since_update (step_cb runs AFTER each iteration)
    history, events = [], []
    anneal = 1.0
    t_start = time.time()
    for step in range(steps):
        # BEFORE_TRAIN_ITERATION: proposal-weight annealing (fruit_nerf.py:199-211)
        tf = float(np.clip(step / cfg.proposal_weights_anneal_max_num_iters, 0, 1))
        b = cfg.proposal_weights_anneal_slope
        anneal = b * tf / ((b - 1) * tf + 1)
        updated = bool(since_update > update_sched(sampler_step) or sampler_step < 10)
        for name, opt in opts.items():
            for g in opt.param_groups:
                g["lr"] = scheds[name].lr(step)  # nerfstudio steps every scheduler every iteration
            opt.zero_grad(set_to_none=True)
        bundle, batch = dm.next_train(step)
        o, d, cam = bundle.origins, bundle.directions, bundle.camera_indices[:, 0]
        nears, fars = torch.full((R, 1), cfg.near_plane), torch.full((R, 1), cfg.far_plane)
        out = _forward(model, fparams, spec, pparams, pspecs, o, d, cam, nears, fars, True, updated, anneal)
        if updated:
            since_update = 0
        losses = fr.loss_dict(out, batch["image"], batch["fruit_mask"], cfg.semantic_loss_weight)
        losses["interlevel_loss"] = cfg.interlevel_loss_mult * ns.interlevel_loss(out["weights_list"], out["sdist_list"])
        loss = sum(losses.values())
        loss.backward()
        opts["fields"].step()
        if updated:  # torch optimisers skip parameters without a gradient: same effect
            opts["proposal_networks"].step()
        # AFTER_TRAIN_ITERATION: sampler.step_cb
        sampler_step = step
        since_update += 1
        lv = float(loss.detach())
        if not math.isfinite(lv):
            events.append({"step": step + 1, "event": "non-finite loss"})
            log(f"step {step + 1}: non-finite loss, stopping")
            break
        if (step + 1) % log_every == 0 or step == 0:
            with torch.no_grad():
                psnr = float(-10.0 * torch.log10(torch.mean((out["rgb"] - batch["image"]) ** 2)))
                acc = float(out["accumulation"].mean())
                gnorm = float(torch.sqrt(sum((p.grad ** 2).sum() for p in groups["fields"] if p.grad is not None)))
            row = {"step": step + 1, "loss": lv, "psnr": psnr, "mean_accumulation": acc, "fields_grad_norm": gnorm, "elapsed_s": time.time() - t_start,
                   **{k: float(v) for k, v in losses.items()}}
            history.append(row)
            log(json.dumps(row))
            if state_path:
                torch.save({"step": step + 1, "pipeline": pipeline.state_dict()}, state_path)
    train_s = time.time() - t_start
    res = {"what": "CPU oracle (oracle/fruit_ref.py + oracle/ns_torch.py) trained with torch.optim.Adam; mirrors fruitnerf_b200.scripts.train",
           "method": method, "steps": steps, "seed": seed, "stream": stream, "lr_schedule": "1e-2 -> 1e-4 over the run" if short_schedule else "stock (200k steps)",
           "rays_per_batch": R, "threads": torch.get_num_threads(), "train_seconds": train_s, "train_rays_per_s": len(history) and history[-1]["step"] * R / train_s,
           "history": history, "events": events, "scene": {"images": num_images, "size": image_size, "fruits": num_fruits}}
    pipeline.eval()
    res["eval"], res["eval_rows"] = evaluate(model, fparams, spec, pparams, pspecs, dm, anneal)
    log(json.dumps({"eval": res["eval"]}))
    if do_export:
        res["export"] = export_and_count(fparams, spec, dm, points_per_side)
        log(json.dumps({"export": res["export"]}))
    return res


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=3000)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--stock-schedule", action="store_true")
    ap.add_argument("--threads", type=int, default=0)
    ap.add_argument("--log-every", type=int, default=250)
    ap.add_argument("--image-size", type=int, default=160)
    ap.add_argument("--num-images", type=int, default=40)
    ap.add_argument("--num-fruits", type=int, default=12)
    ap.add_argument("--rays-per-batch", type=int, default=None)
    ap.add_argument("--points-per-side", type=int, default=256)
    ap.add_argument("--no-export", action="store_true")
    ap.add_argument("--stream", type=int, default=0, help="non-zero: same scene / initial parameters, another stream of batches and jitter")
    ap.add_argument("--state", default=None, help="path of a periodically rewritten pipeline state dict")
    ap.add_argument("--json", default=None)
    a = ap.parse_args(argv)
    if a.threads:
        torch.set_num_threads(a.threads)
    res = train(a.steps, a.seed, not a.stock_schedule, a.log_every, image_size=a.image_size, num_images=a.num_images, num_fruits=a.num_fruits,
                rays_per_batch=a.rays_per_batch, points_per_side=a.points_per_side, state_path=a.state, do_export=not a.no_export, stream=a.stream)
    if a.json:
        os.makedirs(os.path.dirname(a.json) or ".", exist_ok=True)
        with open(a.json, "w") as f:
            json.dump(res, f, indent=1)
    print(json.dumps({k: v for k, v in res.items() if k not in ("history", "eval_rows")}))


if __name__ == "__main__":
    main()
