"""Tiny driver for ncu: a few fwd+bwd steps of the bench workload (no timing, no sampling)."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch

import bench
from fruitnerf_b200 import _lib as L
from fruitnerf_b200 import synthetic as syn

variant = sys.argv[1] if len(sys.argv) > 1 else "small"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dev = torch.device("cuda:0")
field = bench.build_field(variant, dev)
o, d, s, e, cam = syn.ray_batch(4096, 192, num_images=100)
img, mask = syn.targets(4096)
batch = [t.to(dev) for t in (o, d, s, e, cam.to(torch.int32), img, mask)]
for _ in range(steps):
    for p in field.kernel_params():
        p.grad = None
    out, loss = bench.step_fn(field, batch, 1, L.FNR_IMPL_AUTO)
    loss.backward()
torch.cuda.synchronize()
print("done", float(loss))
