"""Tiny driver for ncu: a few op-by-op training iterations on the synthetic scene (proposal + field + optimiser kernels)."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
from fruitnerf_b200.scripts.train import synthetic_spec
from fruitnerf_b200.trainer import Trainer

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 6
torch.manual_seed(0)
tr = Trainer(synthetic_spec("fruit_nerf", num_images=10, image_size=64), device="cuda:0", use_cuda_graph=False)
tr.train(steps)
torch.cuda.synchronize()
print("done")
