"""Multi-GPU check of the NVLS gradient exchange kernel against NCCL (run under torchrun, N >= 2):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/r2/nvls_test.py
Correctness: mean over ranks of seeded fp32 buffers (bit-comparison against the fp64 mean within fp32 rounding; bf16 wire within
bf16 rounding), timing: CUDA events, max over ranks, for the two flat-gradient sizes of the bench (fruit_nerf 67 MB, _big 269 MB)."""
import os
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent.parent))
import torch
import torch.distributed as dist

from fruitnerf_b200.grad_exchange import make_gradient_exchange

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)


def fill(t, r):
    g = torch.Generator(device=dev).manual_seed(1234 + r)
    t.copy_(torch.randn(t.shape, device=dev, generator=g) * 1e-2)


def timed(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    dist.barrier()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev[0].record()
    for _ in range(iters):
        fn()
    ev[1].record()
    torch.cuda.synchronize()
    t = torch.tensor([ev[0].elapsed_time(ev[1]) / iters], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)


for numel in (16_807_312, 67_262_416):
    for kind in ("nccl", "nvls", "nvls_bf16"):
        try:
            ex = make_gradient_exchange(numel, world, dev, kind=kind)
        except Exception as e:  # noqa: BLE001
            if rank == 0:
                print(f"{kind} n={numel}: unavailable: {type(e).__name__}: {str(e)[:200]}", flush=True)
            continue
        fill(ex.flat, rank)
        # expected mean in fp64 from the same seeds
        want = torch.zeros(numel, device=dev, dtype=torch.float64)
        tmp = torch.empty(numel, device=dev, dtype=torch.float32)
        for r in range(world):
            fill(tmp, r)
            want += tmp.double()
        want /= world
        ex()
        torch.cuda.synchronize()
        err = float((ex.flat.double() - want).abs().max())
        scale = float(want.abs().max())
        tol = (2e-7 if kind != "nvls_bf16" else 1.2e-2) * scale * max(1, world // 2)
        ok = err <= tol
        ms = timed(ex)
        nbytes = numel * 4
        if rank == 0:
            print(f"{ex.kind:10s} n={numel} ({nbytes / 1e6:.0f} MB): max err {err:.3e} (tol {tol:.1e}) {'OK' if ok else 'MISMATCH'}  {ms * 1e3:.1f} us  "
                  f"algbw {nbytes / ms / 1e6:.0f} GB/s  busbw {nbytes / ms / 1e6 * 2 * (world - 1) / world:.0f} GB/s  {ex.describe()['note'][:90]}", flush=True)
        del ex, want, tmp
        torch.cuda.empty_cache()
dist.barrier()
dist.destroy_process_group()
