"""Diagnostic run of the synthetic training: history with held-out metrics, parameter health, kernel profile."""
import json, sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
from fruitnerf_b200.scripts.train import synthetic_spec, phase_timing_ms
from fruitnerf_b200.trainer import Trainer

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
prof_at = int(sys.argv[2]) if len(sys.argv) > 2 else steps
torch.manual_seed(0)
spec = synthetic_spec("fruit_nerf", num_images=int(os.environ.get("NIMG", 40)))
if os.environ.get("NOISE"):
    spec.pipeline.datamanager.synthetic_scene["noise_std"] = float(os.environ["NOISE"])
if os.environ.get("ELEV"):
    spec.pipeline.datamanager.synthetic_scene["elevations"] = tuple(float(v) for v in os.environ["ELEV"].split(","))
tr = Trainer(spec, device="cuda:0", use_cuda_graph=os.environ.get("NO_GRAPH") is None)
model = tr.pipeline.model


def health():
    out = {}
    for name, p in model.named_parameters():
        if p.numel() == 0 or name.startswith("field.mlp_base.") or ".mlp_base.0." in name:
            continue
        if not torch.isfinite(p).all():
            out[name] = "NONFINITE"
    f = model.field
    out["table_absmax"] = float(f.mlp_base_grid.hash_table.abs().max())
    out["base_w1_absmax"] = float(f.mlp_base_mlp.layers[1].weight.abs().max())
    out["base_b1_0"] = float(f.mlp_base_mlp.layers[1].bias[0])
    out["prop0_table_absmax"] = float(model.proposal_networks[0].encoding.hash_table.abs().max())
    # one training batch through the model, eagerly
    rb, batch = tr.pipeline.datamanager.next_train(0)
    with torch.no_grad():
        o = model(rb)
        w = o["weights_list"][-1][..., 0]
        out["acc_mean"] = float(o["accumulation"].mean())
        out["w_nan"] = int(torch.isnan(w).sum())
        out["w_first"] = float(w[:, 0].mean())
        out["w_last"] = float(w[:, -1].mean())
        out["sem_absmax"] = float(o["semantics"].abs().max())
        rs = o["ray_samples_list"][-1].frustums
        out["t_first"] = float(rs.starts[:, 0].mean())
        out["t_last"] = float(rs.ends[:, -1].mean())
        pw = o["weights_list"][0][..., 0]
        out["prop0_acc"] = float(pw.sum(-1).mean())
        f = model.field
        from fruitnerf_b200 import ops, _lib as L
        shape, oo, dd, ss, ee, cam = type(f).ray_tensors(o["ray_samples_list"][-1])
        r = ops.render(f.kernel_shape(), f.kernel_params(), oo, dd, ss, ee, cam, f.position_mode(), f.appearance_mode())
        dens = r["sample_density"]
        out["dens_nan"] = int(torch.isnan(dens).sum()); out["dens_inf"] = int(torch.isinf(dens).sum())
        out["dens_max"] = float(torch.nan_to_num(dens, posinf=0).max()); out["dens_median"] = float(dens.median())
        out["rgb_nan"] = int(torch.isnan(r["sample_rgb"]).sum())
    return out


def profile(n=4):
    from torch.profiler import profile as tprof, ProfilerActivity
    with tprof(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as p:
        for _ in range(n):
            tr.train_iteration(tr.step)
            tr.step += 1
        torch.cuda.synchronize()
    rows = sorted(p.key_averages(), key=lambda e: -e.device_time_total)[:14]
    for e in rows:
        print(f"   {e.key[:70]:70s} n={e.count:4d} cuda_total={e.device_time_total/1e3/n:8.3f} ms/iter")


done = 0
while done < steps:
    chunk = min(250, steps - done)
    hist = tr.train(chunk, log_every=chunk, eval_every=chunk)
    done += chunk
    row = hist[-1]
    print({k: (round(v, 4) if isinstance(v, float) else v) for k, v in row.items() if k != "eval"}, "eval_psnr", round(row["eval"]["psnr"], 2), health(), flush=True)
    if done == prof_at:
        profile()
