"""Training-loop shell: what nerfstudio's ``Trainer`` does around the hot path for the
three FruitNeRF methods (fruit_nerf/fruit_nerf_config.py:27-164) -- callbacks before / after each iteration
(proposal-weight annealing, fruit_nerf/fruit_nerf.py:191-223), loss = sum(loss_dict), backward, gradient exchange,
one fused Adam / RAdam launch per param group with the exponential-decay schedule, periodic evaluation and
``step-XXXXXXXXX.ckpt`` checkpoints holding {step, pipeline, optimizers}.

``use_cuda_graph=True`` captures the device work of a whole iteration -- pixel sampling and ray generation, the two
proposal levels, the fused render forward, losses, all backward kernels and the optimiser launches -- into CUDA
graphs (one per branch of the proposal-update schedule) and replays them: the ~1.5 ms of GPU work per iteration is
otherwise buried under 4-40 ms of Python / launch overhead.  Schedules keep moving between replays because the
learning rates, bias corrections and the proposal-weight annealing exponent live in device memory.

Not carried over: torch.autocast + GradScaler (``mixed_precision=True`` upstream exists for tinycudann's fp16
parameters; the kernels here keep fp32 parameters and fp32-class arithmetic, so there is nothing to scale), the
viewer, tensorboard / wandb writers.
"""
from __future__ import annotations

import pathlib
import time
from typing import Dict, List, Optional

import torch
import torch.distributed as dist

from .fruit_nerf_config import TrainerSpec
from .optim import FusedAdam, build_optimizers


class Trainer:
    def __init__(self, spec: TrainerSpec, device="cuda:0", world_size: int = 1, local_rank: int = 0, output_dir: Optional[str] = None,
                 use_cuda_graph: bool = False, graph_warmup: int = 3):
        self.spec = spec
        self.device = torch.device(device)
        if self.device.type == "cuda":
            # the native launchers use the CURRENT device (one process per GPU): make it the trainer's
            torch.cuda.set_device(self.device)
        self.world_size, self.local_rank = world_size, local_rank
        self.output_dir = pathlib.Path(output_dir) if output_dir else None
        self.pipeline = spec.pipeline.setup(device=self.device, test_mode="val", world_size=world_size, local_rank=local_rank)
        self.pipeline.train()
        if world_size > 1 and dist.is_available() and dist.is_initialized():
            # DDP broadcasts rank 0's parameters and buffers at construction (fruit_pipeline.py:117): replicas must start
            # identical because only gradients are exchanged afterwards
            with torch.no_grad():
                for t in list(self.pipeline.parameters()) + list(self.pipeline.buffers()):
                    if t.numel():
                        dist.broadcast(t.data, src=0)
        if self.output_dir is not None and local_rank == 0:
            self.save_config()
        self.param_groups = self.pipeline.get_param_groups()
        self.optimizers: Dict[str, FusedAdam] = build_optimizers(self.param_groups, spec.optimizers)
        self.callbacks: List[Dict] = list(self.pipeline.get_training_callbacks(None))  # datamanager + model (fruit_pipeline.py:242-249)
        self.step = 0
        self.use_cuda_graph = bool(use_cuda_graph) and self.device.type == "cuda"
        self.graph_warmup = graph_warmup
        self._graphs: Dict[bool, tuple] = {}
        self._eager_iterations = 0
        self._side_stream = None
        self._sampler = getattr(self.pipeline.model, "proposal_sampler", None)
        if not hasattr(self._sampler, "wants_update"):
            self._sampler = None
        if self.device.type == "cuda":
            seed = getattr(spec.pipeline.datamanager, "seed", 0)
            torch.cuda.manual_seed(seed * 7919 + local_rank)  # each rank draws its own rays (fruit_pipeline.py:97-99)
        if self.use_cuda_graph and self._sampler is not None:
            self._sampler.enable_device_anneal(self.device)

    # ---- one iteration (nerfstudio Trainer.train_iteration) ---------------------------------------------------
    def _run_callbacks(self, where: str, step: int) -> None:
        for cb in self.callbacks:
            if cb["where_to_run"] == where and step % cb.get("update_every_num_iters", 1) == 0:
                cb["func"](step)

    def _exchange(self, grads) -> None:
        """Mean of the gradients over ranks (the reference wraps the model in DDP, fruit_pipeline.py:117).  Field
        gradients are views of one flat buffer: a single collective; other groups are coalesced first."""
        from .fruit_pipeline import sync_gradients

        base = grads[0]._base
        if base is not None and all(g._base is base for g in grads):
            sync_gradients(base, self.world_size)
            return
        flat = torch.cat([g.reshape(-1) for g in grads])
        sync_gradients(flat, self.world_size)
        off = 0
        for g in grads:
            g.copy_(flat[off:off + g.numel()].view_as(g))
            off += g.numel()

    def _device_work(self, step: int, launch_only: bool, with_optimizer: bool = True):
        """Everything an iteration enqueues on the GPU (this is what a CUDA graph captures)."""
        for params in self.param_groups.values():
            for p in params:
                p.grad = None  # optimizer.zero_grad(set_to_none=True)
        _, loss_dict, metrics_dict = self.pipeline.get_train_loss_dict(step)
        loss = sum(loss_dict.values())
        loss.backward()
        if with_optimizer:
            self._exchange_and_step({name: [p.grad for p in params] for name, params in self.param_groups.items()}, launch_only, step)
        # detached: nothing may keep the autograd graph (and its AccumulateGrad nodes, which remember their stream) alive
        # across iterations -- a later CUDA-graph capture runs on a different stream
        return loss.detach(), {k: v.detach() for k, v in loss_dict.items()}, {k: v.detach() for k, v in metrics_dict.items()}

    def _exchange_and_step(self, grads_by_group: Dict[str, list], launch_only: bool, step: Optional[int] = None) -> None:
        for name, opt in self.optimizers.items():
            grads = grads_by_group[name]
            if any(g is None for g in grads):
                # torch optimisers skip parameters without a gradient: the proposal networks on the iterations the
                # sampler runs them under no_grad (update_sched, fruit_nerf.py:131-136).  Their learning-rate schedule
                # still follows the trainer step (nerfstudio's scheduler_step_all): the next update reads lr(step).
                continue
            if self.world_size > 1:
                self._exchange(grads)
            if launch_only:
                opt.launch(grads)
            else:
                opt.step(grads, sched_step=step)

    def train_iteration(self, step: int):
        self._run_callbacks("BEFORE_TRAIN_ITERATION", step)
        if not self.use_cuda_graph:
            result = self._device_work(step, launch_only=False)
        else:
            sampler = self._sampler
            updated = sampler.wants_update() if sampler is not None else True
            for name, opt in self.optimizers.items():  # advance the schedules of the groups this iteration steps
                if updated or name != "proposal_networks":
                    opt.prepare(sched_step=step)
            if sampler is not None:
                sampler.force_updated = updated
            if self._eager_iterations < self.graph_warmup:
                # warm-up iterations run on a side stream, as torch asks for before a whole-network capture
                self._eager_iterations += 1
                if self._side_stream is None:
                    self._side_stream = torch.cuda.Stream(device=self.device)
                cur = torch.cuda.current_stream(self.device)
                self._side_stream.wait_stream(cur)
                with torch.cuda.stream(self._side_stream):
                    result = self._device_work(step, launch_only=True)
                cur.wait_stream(self._side_stream)
            else:
                if updated not in self._graphs:
                    counters = (self.pipeline.datamanager.train_count, sampler._steps_since_update if sampler is not None else 0)
                    graph = torch.cuda.CUDAGraph()
                    # multi-GPU: the collective and the optimiser launches stay outside the graph (NCCL inside a captured
                    # region deadlocked on this stack); single GPU: the whole iteration is one graph
                    in_graph_opt = self.world_size == 1
                    with torch.cuda.graph(graph):
                        captured = self._device_work(step, launch_only=True, with_optimizer=in_graph_opt)
                    grads = None if in_graph_opt else {name: [p.grad for p in params] for name, params in self.param_groups.items()}
                    self._graphs[updated] = (graph, captured, grads)
                    self.pipeline.datamanager.train_count = counters[0]  # capture ran the Python side once without executing
                    if sampler is not None:
                        sampler._steps_since_update = counters[1]
                graph, result, static_grads = self._graphs[updated]
                graph.replay()
                if static_grads is not None:
                    self._exchange_and_step(static_grads, launch_only=True, step=step)
                self.pipeline.datamanager.train_count += 1
                if sampler is not None and updated:
                    sampler._steps_since_update = 0
            if sampler is not None:
                sampler.force_updated = None
        self._run_callbacks("AFTER_TRAIN_ITERATION", step)
        return result

    def train(self, num_iterations: Optional[int] = None, eval_every: Optional[int] = None, log_every: int = 0) -> List[Dict]:
        n = num_iterations if num_iterations is not None else self.spec.max_num_iterations
        eval_every = eval_every or self.spec.steps_per_eval_batch
        history = []
        t0 = time.time()
        for _ in range(n):
            loss, loss_dict, metrics = self.train_iteration(self.step)
            self.step += 1
            if log_every and self.step % log_every == 0:
                row = {"step": self.step, "loss": float(loss), "psnr": float(metrics["psnr"]), "elapsed_s": time.time() - t0,
                       **{k: float(v.detach()) for k, v in loss_dict.items()}}
                if eval_every and self.step % eval_every == 0 and getattr(self.pipeline.datamanager, "eval_dataset", None) is not None:
                    row["eval"] = self.pipeline.get_eval_image_metrics_and_images(self.step)[0]
                history.append(row)
            if self.output_dir is not None and self.step % self.spec.steps_per_save == 0:
                self.save_checkpoint()
        return history

    # ---- config.yml (nerfstudio TrainerConfig.save_config): what eval_setup / ns-export-semantics load back ------------
    def save_config(self) -> pathlib.Path:
        import yaml

        self.output_dir.mkdir(parents=True, exist_ok=True)
        path = self.output_dir / "config.yml"
        path.write_text(yaml.dump(self.spec))
        outs = getattr(self.pipeline.datamanager, "train_dataparser_outputs", None)
        if outs is not None:
            outs.save_dataparser_transform(self.output_dir / "dataparser_transforms.json")
        else:  # in-memory scene: identity transform, the data set's own scale
            import json

            scale = float(getattr(self.pipeline.datamanager.train_dataset, "dataparser_scale", 1.0))
            (self.output_dir / "dataparser_transforms.json").write_text(
                json.dumps({"transform": [[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 1, 0]], "scale": scale}, indent=4))
        return path

    # ---- checkpoints (nerfstudio Trainer.save_checkpoint / _load_checkpoint) --------------------------------------
    def save_checkpoint(self, path: Optional[str] = None) -> pathlib.Path:
        if path is None:
            ckpt_dir = self.output_dir / "nerfstudio_models"
            ckpt_dir.mkdir(parents=True, exist_ok=True)
            path = ckpt_dir / f"step-{self.step:09d}.ckpt"
        path = pathlib.Path(path)
        if self.local_rank == 0:
            torch.save({"step": self.step, "pipeline": self.pipeline.state_dict(),
                        "optimizers": {k: o.state_dict() for k, o in self.optimizers.items()}}, path)
        return path

    def load_checkpoint(self, path) -> None:
        state = torch.load(path, map_location=self.device, weights_only=False)
        self.step = int(state["step"])
        self.pipeline.load_pipeline(state["pipeline"], self.step)
        for k, o in self.optimizers.items():
            o.load_state_dict(state["optimizers"][k])
            o.sched_step = self.step  # schedules follow the trainer step
        sampler = getattr(self.pipeline.model, "proposal_sampler", None)
        if hasattr(sampler, "step_cb"):
            sampler._step = self.step
