"""Per-iteration gradient / parameter magnitudes around the training collapse (diagnostic)."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
from fruitnerf_b200.scripts.train import synthetic_spec
from fruitnerf_b200.trainer import Trainer

torch.manual_seed(0)
tr = Trainer(synthetic_spec("fruit_nerf"), device="cuda:0", use_cuda_graph=True)
model = tr.pipeline.model
f = model.field
watch = {"b1": f.mlp_base_mlp.layers[1].bias, "w1": f.mlp_base_mlp.layers[1].weight, "w0": f.mlp_base_mlp.layers[0].weight,
         "table": f.mlp_base_grid.hash_table, "col0": f.mlp_head.layers[0].weight, "p0tab": model.proposal_networks[0].encoding.hash_table,
         "p0w1": model.proposal_networks[0].mlp_base[1].layers[1].weight}
rows = []
for step in range(int(sys.argv[1]) if len(sys.argv) > 1 else 4000):
    loss, ld, m = tr.train_iteration(tr.step)
    tr.step += 1
    if tr.step < 1200:
        continue
    rec = {"step": tr.step, "loss": float(loss), "sem": float(ld["semantics_loss"])}
    for k, p in watch.items():
        g = p.grad
        rec["g_" + k] = float(g.abs().max()) if g is not None else None
    rec["b1_0"] = float(f.mlp_base_mlp.layers[1].bias[0])
    rec["v_b1_0"] = float(tr.optimizers["fields"].exp_avg_sq[[id(q) for q in tr.optimizers["fields"].params].index(id(f.mlp_base_mlp.layers[1].bias))][0])
    rows.append(rec)
    if rec["sem"] > 0.69 and tr.step > 1300:
        break
for r in rows[-45:]:
    print({k: (f"{v:.3g}" if isinstance(v, float) else v) for k, v in r.items()})
