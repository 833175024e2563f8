"""Where does a step's time go?  Host enqueue time vs GPU time per phase, plus a cProfile of the
un-synchronised step loop.  Run on the GPU box: python tools/host_overhead.py [simt|tcgen05|auto]"""
import cProfile
import pstats
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch

import bench
from fruitnerf_b200 import _lib as L
from fruitnerf_b200 import ops
from fruitnerf_b200 import synthetic as syn

kernel = sys.argv[1] if len(sys.argv) > 1 else "auto"
impl = {"auto": L.FNR_IMPL_AUTO, "simt": L.FNR_IMPL_SIMT, "tcgen05": L.FNR_IMPL_TCGEN05}[kernel]
dev = torch.device("cuda:0")
field = bench.build_field("small", dev)
params = field.kernel_params()
o, d, s, e, cam = syn.ray_batch(4096, 192, num_images=100)
img, mask = syn.targets(4096)
batch = [t.to(dev) for t in (o, d, s, e, cam.to(torch.int32), img, mask)]


def step(rec=None):
    for p in params:
        p.grad = None
    t0 = time.perf_counter()
    if rec: rec[0].record()
    out, loss = bench.step_fn(field, batch, 1, impl)
    if rec: rec[1].record()
    t1 = time.perf_counter()
    loss.backward()
    if rec: rec[2].record()
    t2 = time.perf_counter()
    return t1 - t0, t2 - t1


for _ in range(5):
    step()
torch.cuda.synchronize()
print(f"kernel={kernel}")
for trial in range(2):
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(10)]
    host = []
    t0 = time.perf_counter()
    for i in range(10):
        host.append(step(rec=evs[i]))
    t_enq = time.perf_counter() - t0
    torch.cuda.synchronize()
    t_all = time.perf_counter() - t0
    f = [ev[0].elapsed_time(ev[1]) for ev in evs]
    b = [ev[1].elapsed_time(ev[2]) for ev in evs]
    print(f"trial {trial}: wall {t_all*100:.2f} ms/step, host enqueue {t_enq*100:.2f} ms/step; host fwd {sum(h[0] for h in host)*100:.3f} "
          f"bwd {sum(h[1] for h in host)*100:.3f} ms/step; gpu(events) fwd {sum(f)/10:.3f} bwd {sum(b)/10:.3f} ms")
    print("   per-step gpu fwd:", " ".join(f"{x:.2f}" for x in f))
    print("   per-step gpu bwd:", " ".join(f"{x:.2f}" for x in b))

with torch.no_grad():
    for _ in range(3):
        bench.step_fn(field, batch, 1, impl)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        ops.render(field.kernel_shape(), params, *batch[:5], field.position_mode(), field.appearance_mode(), impl=impl)
    e1.record()
    torch.cuda.synchronize()
    print(f"forward-only (no stash), 10 back-to-back renders: {e0.elapsed_time(e1)/10:.3f} ms each")

pr = cProfile.Profile()
pr.enable()
for i in range(10):
    step()
pr.disable()
torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("cumulative").print_stats(14)
