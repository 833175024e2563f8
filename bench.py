#!/usr/bin/env python
"""bench.py -- rays/s of the FruitNeRF hot path (fused field + compositing, forward + backward) on
synthetic 4096-ray x 192-sample batches (BASELINE.json metric), N GPUs of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--variant small|big] [--impl ours|reference]

One "step" = one pass of the hot path over one batch: render forward (hash encode -> MLPs ->
composite), MSE + BCE loss, backward into the flat gradient buffer (and, for N > 1, one NCCL
all-reduce of that buffer -- the reference's DDP exchange, fruit_pipeline.py:117).  Prints ONE JSON
line on rank 0.  See DESIGN.md "Measurement" for the definitions of every field.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

import torch  # noqa: E402

R_RAYS, S_SAMPLES = 4096, 192
NUM_IMAGES = 100
HASH_BYTES_PER_POINT_FWD = 16 * 8 * 2 * 4  # L levels x 8 corners x F=2 x fp32
HASH_BYTES_PER_POINT_BWD = 2 * HASH_BYTES_PER_POINT_FWD  # read-modify-write scatter


def _peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        j = json.loads(p.read_text())
        return float(j["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock / throttle-reason sampling during the timed region (B200_PROFILING.md clocks line),
    through NVML in-process (an `nvidia-smi -lms` poller next to the 256 MiB L2-flush fills stalled
    the GPU for ~10 ms per step on this pool; NVML queries from a thread do not)."""

    def __init__(self, index: int, period_s: float = 0.02):
        self.index, self.period = index, period_s
        self.samples, self.reasons = [], set()
        self.max_mhz = None
        self._stop = threading.Event()
        self._thread = None
        self._nv = None

    def start(self):
        try:
            import pynvml as nv

            nv.nvmlInit()
            self._nv = nv
            self._h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(self._h, nv.NVML_CLOCK_SM)
        except Exception as e:  # noqa: BLE001
            self._nv = None
            self._err = repr(e)
            return
        self._thread = threading.Thread(target=self._run, daemon=True)
        self._thread.start()

    def _run(self):
        nv = self._nv
        names = {
            "hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
            "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
            "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
            "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4),
        }
        get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM))
                mask = get_reasons(self._h)
                for n, bit in names.items():
                    if mask & bit:
                        self.reasons.add(n)
            except Exception:  # noqa: BLE001
                pass
            self._stop.wait(self.period)

    def stop(self):
        if self._nv is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [f"nvml unavailable: {getattr(self, '_err', '')}"]}
        self._stop.set()
        self._thread.join(timeout=2)
        sm = sorted(self.samples)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(sm), "source": "nvml"}


def build_field(variant: str, device, table_scale: float = 1e-1):
    """Random-init FruitField of the named variant, parameters from torch's RNG."""
    from fruitnerf_b200.fruit_field import FruitField, SceneContraction

    torch.manual_seed(0)
    kw = dict(geo_feat_dim=15, max_res=2048, log2_hashmap_size=19, num_layers_semantic=2, hidden_dim_semantics=64)
    if variant == "big":
        kw = dict(geo_feat_dim=30, max_res=4096, log2_hashmap_size=21, num_layers_semantic=3, hidden_dim_semantics=128)
    aabb = torch.tensor([[-1.0, -1, -1], [1, 1, 1]])
    f = FruitField(aabb, num_images=NUM_IMAGES, use_semantics=True, num_semantic_classes=1,
                   spatial_distortion=SceneContraction(order=float("inf")), **kw)
    with torch.no_grad():
        f.mlp_base_grid.hash_table.mul_(table_scale / 1e-3)  # U(-1,1) * table_scale
    return f.to(device).train()


def step_fn(field, batch, world, impl_id):
    """One eager step through the public op (the call FruitModel.get_outputs makes) -- used by tools/."""
    from fruitnerf_b200 import ops

    o, d, s, e, cam, img, mask = batch
    out = ops.render(field.kernel_shape(), field.kernel_params(), o, d, s, e, cam, field.position_mode(), field.appearance_mode(),
                     impl=impl_id)
    loss = torch.nn.functional.mse_loss(img, out["rgb"]) + torch.nn.functional.binary_cross_entropy_with_logits(
        out["semantics"][:, None], mask)
    return out, loss


def run_ours(args):
    from fruitnerf_b200 import _lib as L
    from fruitnerf_b200 import ops
    from fruitnerf_b200 import synthetic as syn
    from fruitnerf_b200.engine import GraphedTrainStep

    L.load()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs CUDA devices (no CPU fallback in the product path)")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world == 1 and args.gpus > 1:
        raise SystemExit("launch with torch.distributed.run --nproc-per-node N for --gpus N > 1")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=dev)
    impl_id = {"auto": L.FNR_IMPL_AUTO, "simt": L.FNR_IMPL_SIMT, "tcgen05": L.FNR_IMPL_TCGEN05}[args.kernel]
    field = build_field(args.variant, dev)
    N_pts = R_RAYS * S_SAMPLES

    # per-rank batch (weak scaling: each rank draws its own 4096 rays, fruit_pipeline.py:97-99)
    o, d, s, e, cam = syn.ray_batch(R_RAYS, S_SAMPLES, salt=rank, num_images=NUM_IMAGES)
    img, mask = syn.targets(R_RAYS, salt=rank)
    host = [t.pin_memory() for t in (o, d, s, e, cam.to(torch.int32), img, mask)]
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)  # > 126 MB L2

    # the public training-step API: render fwd + loss + bwd captured in one CUDA graph
    step = GraphedTrainStep(field, R_RAYS, S_SAMPLES, impl=impl_id, use_graph=not args.no_graph)
    h2d = step.load_batch(*host)
    packed = step.pack_batch(*host)  # the same batch as one pinned byte buffer: one H2D copy per step in the e2e loop
    step.capture(warmup=max(args.warmup, 3))

    def one_step():
        loss = step()
        if world > 1:
            import torch.distributed as dist

            dist.all_reduce(step.flat_grad, op=dist.ReduceOp.AVG)  # the reference's DDP exchange
        return loss

    for _ in range(max(args.warmup, 3)):
        one_step()
    torch.cuda.synchronize()

    sampler = ClockSampler(local) if (rank == 0 and not args.no_clocks) else None
    if sampler:
        sampler.start()  # before the barrier: host work on rank 0 between the barrier and the first timed step would show up
                         # as a long first step on the other ranks (they wait in the all-reduce)
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(2)] for _ in range(args.steps)]
    if world > 1:
        import torch.distributed as dist

        dist.barrier()
    torch.cuda.synchronize()
    for i in range(args.steps):
        if not args.no_flush:
            flush.fill_(float(i))  # evict the table / weights from L2 between timed steps
        evs[i][0].record()
        one_step()
        evs[i][1].record()
    torch.cuda.synchronize()
    if world > 1:
        import torch.distributed as dist

        dist.barrier()
    clocks = sampler.stop() if sampler else None
    step_ms = [ev[0].elapsed_time(ev[1]) for ev in evs]
    if os.environ.get("FNR_BENCH_DEBUG"):
        print(f"rank {rank} step_ms " + " ".join(f"{v:.3f}" for v in step_ms), file=sys.stderr)
    total_ms = torch.tensor([sum(step_ms)], device=dev, dtype=torch.float64)
    if world > 1:
        import torch.distributed as dist

        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
    total_ms = float(total_ms)

    # forward kernel alone (training forward: writes the encoding stash), for the roofline of the fused forward
    st = step.static
    fwd_graph = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())

    def fwd_only():
        return ops.render(field.kernel_shape(), field.kernel_params(), st["origins"], st["directions"], st["starts"], st["ends"],
                          st["camera_indices"], field.position_mode(), field.appearance_mode(), impl=impl_id)

    with torch.cuda.stream(side):
        fwd_only()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    with torch.cuda.graph(fwd_graph):
        fwd_out = fwd_only()
    fev = [[torch.cuda.Event(enable_timing=True) for _ in range(2)] for _ in range(args.steps)]
    for i in range(args.steps):
        if not args.no_flush:
            flush.fill_(float(i))
        fev[i][0].record()
        fwd_graph.replay()
        fev[i][1].record()
    torch.cuda.synchronize()
    fwd_ms = [ev[0].elapsed_time(ev[1]) for ev in fev]
    del fwd_out

    # end-to-end through the public API with HOST buffers: per step H2D of the rays/targets from pinned
    # memory, one graph replay, D2H read of the loss -- all inside the timed region
    e2e_steps = max(args.steps, 20)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    step.load_packed(packed)
    for i in range(e2e_steps):
        # H2D of the NEXT step's rays / bins / targets (one packed pinned buffer) overlaps this step, as a prefetching
        # data loader does; every step still moves one full batch host->device and one loss device->host
        step.prefetch_packed(packed)
        loss = one_step()
        _ = float(loss)  # D2H + sync
        step.commit_prefetched()
    torch.cuda.synchronize()
    e2e_s = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
    if world > 1:
        import torch.distributed as dist

        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_s = float(e2e_s)

    if world > 1:
        import torch.distributed as dist

        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return
    peak, peak_src = _peaks()
    mean_step = total_ms / args.steps
    mean_fwd = sum(fwd_ms) / len(fwd_ms)
    mean_bwd = max(mean_step - mean_fwd, 1e-6)
    fwd_bytes = N_pts * HASH_BYTES_PER_POINT_FWD
    bwd_bytes = N_pts * HASH_BYTES_PER_POINT_BWD
    fwd_ach = fwd_bytes / (mean_fwd * 1e-3) / 1e9
    bwd_ach = bwd_bytes / (mean_bwd * 1e-3) / 1e9
    dominant_is_bwd = mean_bwd >= mean_fwd
    on_tc = args.variant == "small" and args.kernel != "simt"  # the families the tcgen05 kernels serve (fnr_api.cu dispatch)
    fwd_kernel = ("tc_render_forward_kernel" if on_tc else
                  "tc_render_forward_big_kernel" if args.kernel != "simt" else "simt_field_forward_kernel + simt_composite_kernel")
    bwd_kernel = ("tc_field_backward_kernel" if on_tc else
                  "tc_big_backward_chain_kernel + cuBLAS dW GEMMs" if args.kernel != "simt" else "simt_field_backward_kernel")
    line = {
        "metric": "rays/sec (4096 rays x 192 samples) fused fwd+bwd",
        "value": world * R_RAYS * args.steps / (total_ms * 1e-3),
        "unit": "rays/s",
        "n_gpus": world,
        "steps": args.steps,
        "warmup": max(args.warmup, 3),
        "ms_per_step": mean_step,
        "higher_is_better": True,
        "scaling": "weak",
        "vs_baseline": None,
        "dtype": "f32",
        "data": "synthetic",
        "config": {
            "workload": f"fruit_nerf{'_big' if args.variant == 'big' else ''} field ({args.variant}): {R_RAYS} rays x {S_SAMPLES} samples "
                        "per GPU, render fwd + MSE/BCE loss + bwd" + (" + NCCL grad all-reduce" if world > 1 else ""),
            "variant": args.variant,
            "kernel": args.kernel,
            "rays_per_gpu": R_RAYS,
            "samples_per_ray": S_SAMPLES,
            "execution": "eager" if args.no_graph else "one CUDA graph per step (fruitnerf_b200.engine.GraphedTrainStep)",
            "l2": "flushed between timed steps (256 MiB fill); per-step CUDA-event durations summed",
            "parallelism": f"dp{world}",
        },
        "fwd_ms": mean_fwd,
        "bwd_ms": mean_bwd,
        "fwd_rays_per_s": R_RAYS / (mean_fwd * 1e-3),
        "roofline": {
            "kernel": (f"render backward ({bwd_kernel} + simt_composite_backward_kernel + loss)" if dominant_is_bwd
                       else f"fused render forward ({fwd_kernel})"),
            "bound": "hbm",
            "achieved": bwd_ach if dominant_is_bwd else fwd_ach,
            "peak": peak,
            "unit": "GB/s",
            "frac": (bwd_ach if dominant_is_bwd else fwd_ach) / peak,
            # dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the committed
            # `ncu --set full` capture (profiles/r1_ncu_summary.md, final round-1 tables); small variant only
            "traffic": (NCU_DRAM_BYTES["bwd" if dominant_is_bwd else "fwd"] if on_tc else None),
            "peak_source": peak_src,
            "algorithmic_bytes_per_launch": bwd_bytes if dominant_is_bwd else fwd_bytes,
            "launch_ms": mean_bwd if dominant_is_bwd else mean_fwd,
        },
        "roofline_forward": {"kernel": fwd_kernel, "bound": "hbm", "achieved": fwd_ach, "peak": peak, "unit": "GB/s",
                             "frac": fwd_ach / peak, "launch_ms": mean_fwd, "algorithmic_bytes_per_launch": fwd_bytes},
        "roofline_step": {"bound": "hbm", "achieved": (fwd_bytes + bwd_bytes) / (mean_step * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                          "frac": (fwd_bytes + bwd_bytes) / (mean_step * 1e-3) / 1e9 / peak},
        "e2e": {"value": world * R_RAYS * e2e_steps / e2e_s, "unit": "rays/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                "steps": e2e_steps},
        "gpu_launches": 3 * args.steps,
        "clocks": clocks,
    }
    if world == 1 and not args.no_cpu:
        line["cpu_baseline"] = cpu_baseline(args.variant, sample_rays=args.cpu_rays, repeats=1)
    if world == 1 and not args.no_train:
        line["train_iteration"] = train_iteration_rate(args.variant, dev)
    print(json.dumps(line))


def train_iteration_rate(variant: str, dev, iterations: int = 300):
    """Supplementary number (BASELINE.json configs[1]/[2]): a WHOLE training iteration of the method on the synthetic
    apple scene -- pixel batch, proposal stage, field forward / backward, losses, Adam -- replayed as CUDA graphs by
    fruitnerf_b200.trainer.Trainer.  Never allowed to break the headline line: any failure is reported as a string."""
    try:
        from fruitnerf_b200.scripts.train import synthetic_spec
        from fruitnerf_b200.trainer import Trainer

        method = "fruit_nerf" if variant == "small" else "fruit_nerf_big"
        trainer = Trainer(synthetic_spec(method, schedule_steps=3000), device=dev, use_cuda_graph=True)
        trainer.train(40)  # warm-up: eager iterations + capture of both schedule branches
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        trainer.train(iterations)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        rays = trainer.spec.pipeline.datamanager.train_num_rays_per_batch
        return {"value": rays * iterations / dt, "unit": "rays/s", "ms_per_iteration": 1e3 * dt / iterations, "rays_per_iteration": rays,
                "iterations": iterations, "method": method,
                "what": "data + proposal sampling + field fwd/bwd + losses + optimiser, synthetic apple scene, CUDA-graph replay"}
    except Exception as ex:  # noqa: BLE001
        return {"error": f"{type(ex).__name__}: {ex}"[:300]}


# per-launch DRAM traffic measured by ncu (profiles/r1_ncu_summary.md): tc_render_forward_kernel 57.9 MB read + 87.2 MB
# written; tc_field_backward_kernel 176.9 MB read + 19.7 MB written.  Far below the algorithmic hash bytes because the
# 64 MiB fruit_nerf table is L2-resident.
NCU_DRAM_BYTES = {"fwd": 57_898_240 + 87_247_360, "bwd": 176_867_840 + 19_650_816}


def cpu_baseline(variant: str, sample_rays: int, repeats: int):
    """The oracle (a port: pure-PyTorch restatement of the reference's CPU-runnable torch path) on
    the host cores, fwd+bwd on a bounded sample of the same workload."""
    from fruitnerf_b200 import synthetic as syn
    from oracle import fruit_ref as fr

    v = syn.SMALL if variant == "small" else syn.BIG
    sd = syn.field_state(geo=v["geo"], sem_dims=v["sem_dims"], log2_hashmap_size=v["log2_hashmap_size"], num_images=NUM_IMAGES,
                         table_scale=1e-1)
    spec = fr.FieldSpec(max_res=v["max_res"], log2_hashmap_size=v["log2_hashmap_size"], geo_feat_dim=v["geo"])
    o, d, s, e, cam = syn.ray_batch(sample_rays, S_SAMPLES, num_images=NUM_IMAGES)
    img, mask = syn.targets(sample_rays)
    best = None
    for _ in range(repeats + 1):  # first pass = warm-up
        st = {k: t.clone().requires_grad_(t.is_floating_point() and k != "aabb") for k, t in sd.items()}
        t0 = time.perf_counter()
        f = fr.field_forward(st, spec, o[:, None, :], d[:, None, :], s[..., None], e[..., None], cam, True, "train")
        r = fr.render(f, s[..., None], e[..., None], training=True)
        ld = fr.loss_dict(r, img, mask)
        (ld["rgb_loss"] + ld["semantics_loss"]).backward()
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return {"value": sample_rays / best, "unit": "rays/s", "cores": torch.get_num_threads(), "host_cpus": os.cpu_count(), "kind": "port",
            "sample": f"{sample_rays} rays x {S_SAMPLES} samples of the same workload, fwd+bwd, best of {repeats} after 1 warm-up"}


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path.  The reference's
    arithmetic lives in nerfstudio/tinycudann (absent, not installable: BASELINE.md section 2), so
    the arm is the oracle port on all host threads; each step = a bounded sample of the workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    rays = args.cpu_rays
    vals = []
    for _ in range(max(1, min(args.warmup, 1))):
        cpu_baseline(args.variant, rays, repeats=0)
    t_all = time.perf_counter()
    for _ in range(max(1, min(args.steps, 3))):
        vals.append(cpu_baseline(args.variant, rays, repeats=0))
    dt = time.perf_counter() - t_all
    v = sum(x["value"] for x in vals) / len(vals)
    line = {
        "impl": "reference",
        "metric": "rays/sec (4096 rays x 192 samples) fused fwd+bwd",
        "value": v,
        "unit": "rays/s",
        "n_gpus": int(os.environ.get("WORLD_SIZE", "1")),
        "steps": len(vals),
        "warmup": 1,
        "ms_per_step": dt / len(vals) * 1e3,
        "higher_is_better": True,
        "scaling": "weak",
        "vs_baseline": None,
        "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": f"fruit_nerf field ({args.variant}): {rays}-ray x {S_SAMPLES}-sample sample of the 4096x192 batch, fwd+bwd, CPU",
                   "variant": args.variant},
        "cpu_baseline": {"value": v, "unit": "rays/s", "cores": torch.get_num_threads(), "kind": "port",
                         "sample": f"{rays} rays x {S_SAMPLES} samples per step"},
        "e2e": {"value": v, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--variant", default="small", choices=["small", "big"])
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--kernel", default="auto", choices=["auto", "simt", "tcgen05"])
    ap.add_argument("--cpu-rays", type=int, default=2048)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-train", action="store_true", help="skip the supplementary whole-training-iteration measurement")
    ap.add_argument("--no-graph", action="store_true", help="diagnostic: eager step instead of the CUDA-graph step")
    ap.add_argument("--no-flush", action="store_true", help="diagnostic: skip the L2 flush between timed steps")
    ap.add_argument("--no-clocks", action="store_true", help="diagnostic: do not sample nvidia-smi during the timed region")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
