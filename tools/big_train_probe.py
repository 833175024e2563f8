import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
from fruitnerf_b200.scripts.train import synthetic_spec
from fruitnerf_b200.trainer import Trainer
graph = sys.argv[1] == "graph"
steps = int(sys.argv[2])
torch.manual_seed(0)
spec = synthetic_spec("fruit_nerf_big", schedule_steps=int(sys.argv[3]) if len(sys.argv) > 3 else None)
tr = Trainer(spec, device="cuda:0", use_cuda_graph=graph)
h = tr.train(steps, log_every=max(steps // 6, 1), eval_every=10**9)
for r in h:
    print(graph, {k: (round(v, 4) if isinstance(v, float) else v) for k, v in r.items()})
print(tr.pipeline.get_average_eval_image_metrics())
