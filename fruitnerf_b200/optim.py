"""Optimisers and schedulers of the training shell (nerfstudio ``Optimizers`` equivalents).

The reference trains every param group with Adam / RAdam (lr 1e-2, eps 1e-15) and an exponential-decay
schedule (fruit_nerf/fruit_nerf_config.py:47-56, 90-103, 140-153).  Here a group's update is ONE launch of
``fnr_adam_step`` over all its tensors (the gradients already sit in the flat buffers the backward kernels
wrote); learning rate and bias corrections are read from an 8-float device array, so the launch is
CUDA-graph friendly.  No torch.optim on the product path.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch
from torch import Tensor

from . import _lib as L


class StagedDeviceBuffer:
    """A few floats in device memory that the host rewrites every iteration without ever synchronising: uploads go
    through a ring of pinned staging slots, each guarded by an event so a slot is not rewritten before its copy ran
    (the host may run many graph replays ahead of the GPU)."""

    def __init__(self, count: int, device, slots: int = 16):
        self.device_buffer = torch.zeros(count, dtype=torch.float32, device=device)
        self._pinned = device.type == "cuda" if isinstance(device, torch.device) else str(device).startswith("cuda")
        self._slots = [torch.zeros(count, dtype=torch.float32).pin_memory() if self._pinned else torch.zeros(count) for _ in range(slots)]
        self._events = [None] * slots
        self._next = 0

    def upload(self, values) -> None:
        i = self._next
        self._next = (i + 1) % len(self._slots)
        if self._events[i] is not None:
            self._events[i].synchronize()
        self._slots[i].copy_(torch.as_tensor(values, dtype=torch.float32).reshape(-1))
        self.device_buffer.copy_(self._slots[i], non_blocking=True)
        if self._pinned:
            ev = self._events[i] or torch.cuda.Event()
            ev.record()
            self._events[i] = ev


class ExponentialDecay:
    """nerfstudio ExponentialDecayScheduler (no warm-up, the configuration the reference uses):
    lr(step) = exp(log(lr_init) * (1 - t) + log(lr_final) * t), t = clip(step / max_steps, 0, 1)."""

    def __init__(self, lr_init: float, lr_final: Optional[float], max_steps: Optional[int]):
        self.lr_init, self.lr_final, self.max_steps = lr_init, lr_final, max_steps

    def lr(self, step: int) -> float:
        if self.lr_final is None or not self.max_steps:
            return self.lr_init
        t = min(max(step / self.max_steps, 0.0), 1.0)
        return math.exp(math.log(self.lr_init) * (1 - t) + math.log(self.lr_final) * t)


class FusedAdam:
    """torch.optim.Adam / RAdam semantics (betas (0.9, 0.999), weight decay 0, amsgrad off) on the native kernel."""

    def __init__(self, params: Sequence[Tensor], lr: float = 1e-2, eps: float = 1e-15, betas=(0.9, 0.999), kind: str = "Adam",
                 scheduler: Optional[ExponentialDecay] = None):
        self.params = [p for p in params]
        if not self.params:
            raise ValueError("empty parameter group")
        if len(self.params) > L.FNR_MAX_ADAM_TENSORS:
            raise ValueError(f"at most {L.FNR_MAX_ADAM_TENSORS} tensors per group")
        for p in self.params:
            if not p.is_cuda or p.dtype != torch.float32 or not p.is_contiguous():
                raise L.FruitNerfNativeError("FusedAdam needs contiguous fp32 CUDA parameters (no CPU fallback)")
        self.kind = {"Adam": L.FNR_OPT_ADAM, "RAdam": L.FNR_OPT_RADAM}[kind]
        self.lr, self.eps, self.betas = lr, eps, betas
        self.scheduler = scheduler
        self.step_count = 0   # updates applied to this group (Adam bias correction)
        self.sched_step = 0   # scheduler position = trainer iteration (nerfstudio steps every scheduler every iteration,
                              # also on iterations where the group has no gradient and its optimiser is skipped)
        dev = self.params[0].device
        self.exp_avg = [torch.zeros_like(p) for p in self.params]
        self.exp_avg_sq = [torch.zeros_like(p) for p in self.params]
        self._hyper = StagedDeviceBuffer(8, dev)
        self.hyper = self._hyper.device_buffer
        self._tensors = None
        self._grad_ptrs = None

    # ---- hyper-parameters of the coming step (host side, tiny) -----------------------------------------
    def _hyper_values(self, step: int, grad_scale: float, sched_step: Optional[int] = None) -> np.ndarray:
        b1, b2 = self.betas
        if sched_step is None:
            sched_step = self.sched_step
        lr = self.scheduler.lr(sched_step) * (self.lr / self.scheduler.lr_init) if self.scheduler else self.lr
        bc1, bc2 = 1.0 - b1 ** step, 1.0 - b2 ** step
        rect = -1.0
        if self.kind == L.FNR_OPT_RADAM:
            rho_inf = 2.0 / (1.0 - b2) - 1.0
            rho_t = rho_inf - 2.0 * step * (b2 ** step) / bc2
            if rho_t > 5.0:
                rect = math.sqrt((rho_t - 4) * (rho_t - 2) * rho_inf / ((rho_inf - 4) * (rho_inf - 2) * rho_t))
        return np.array([lr, b1, b2, self.eps, bc1, bc2, rect, grad_scale], dtype=np.float32)

    def prepare(self, grad_scale: float = 1.0, sched_step: Optional[int] = None) -> None:
        """Advance the update counter and upload the hyper-parameters (async, from pinned memory).  ``sched_step`` is the
        trainer iteration the learning rate is read at (the reference's ``scheduler_step_all`` advances every scheduler
        every iteration); without it the schedule advances once per call, which is only right for a group that is
        stepped on every iteration."""
        self.step_count += 1
        if sched_step is not None:
            self.sched_step = int(sched_step)
        self._hyper.upload(self._hyper_values(self.step_count, grad_scale, self.sched_step))
        self._last_sched_step = self.sched_step
        self.sched_step += 1

    def skip(self) -> None:
        """An iteration on which this group has no gradients: torch skips the update, the scheduler still advances."""
        self.sched_step += 1

    def _tensor_array(self, grads: Sequence[Tensor]):
        ptrs = tuple(g.data_ptr() for g in grads)
        if self._tensors is None or ptrs != self._grad_ptrs:
            arr = (L.AdamTensor * len(self.params))()
            for i, (p, g, m, v) in enumerate(zip(self.params, grads, self.exp_avg, self.exp_avg_sq)):
                if g.shape != p.shape or g.dtype != torch.float32 or not g.is_contiguous():
                    raise ValueError("gradient layout must match the parameter")
                arr[i] = L.AdamTensor(p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), p.numel())
            self._tensors, self._grad_ptrs = arr, ptrs
        return self._tensors

    def launch(self, grads: Optional[Sequence[Tensor]] = None) -> None:
        """Enqueue the update with the hyper-parameters currently on the device (graph-capturable)."""
        if grads is None:
            grads = [p.grad for p in self.params]
        if any(g is None for g in grads):
            raise ValueError("every parameter of the group needs a gradient (zero-fill unused ones: the reference's "
                             "find_unused_parameters=True semantics)")
        arr = self._tensor_array(grads)
        dev = self.params[0].device
        L.check(L.load().fnr_adam_step(arr, len(self.params), self.kind, self.hyper.data_ptr(), torch.cuda.current_stream(dev).cuda_stream))

    def step(self, grads: Optional[Sequence[Tensor]] = None, grad_scale: float = 1.0, sched_step: Optional[int] = None) -> None:
        self.prepare(grad_scale, sched_step)
        self.launch(grads)

    @property
    def current_lr(self) -> float:
        """Learning rate of the most recent update (of the coming one before any update)."""
        return float(self._hyper_values(max(self.step_count, 1), 1.0, getattr(self, "_last_sched_step", self.sched_step))[0])

    # ---- checkpointing (nerfstudio stores optimizer.state_dict() per group) --------------------------------
    def state_dict(self) -> Dict:
        return {"step": self.step_count, "sched_step": self.sched_step, "exp_avg": [t.clone() for t in self.exp_avg], "exp_avg_sq": [t.clone() for t in self.exp_avg_sq]}

    def load_state_dict(self, sd: Dict) -> None:
        if "state" in sd and "param_groups" in sd:
            # torch.optim.Adam / RAdam state_dict() as nerfstudio's Optimizers checkpoint it: state[i] = {step, exp_avg, exp_avg_sq}
            # in parameter order; the scheduler position is restored by the trainer (it follows the trainer step)
            steps = []
            for i, (m, v) in enumerate(zip(self.exp_avg, self.exp_avg_sq)):
                st = sd["state"].get(i)
                if st is None:
                    m.zero_()
                    v.zero_()
                    continue
                m.copy_(st["exp_avg"])
                v.copy_(st["exp_avg_sq"])
                steps.append(int(st["step"]))
            if len(set(steps)) > 1:
                raise ValueError(f"per-parameter step counts differ ({sorted(set(steps))}): the fused update keeps one count per group")
            self.step_count = steps[0] if steps else 0
            self.sched_step = self.step_count
            return
        self.step_count = int(sd["step"])
        self.sched_step = int(sd.get("sched_step", sd["step"]))
        for dst, src in zip(self.exp_avg, sd["exp_avg"]):
            dst.copy_(src)
        for dst, src in zip(self.exp_avg_sq, sd["exp_avg_sq"]):
            dst.copy_(src)


def build_optimizers(param_groups: Dict[str, List[Tensor]], optimizer_specs: Dict[str, Dict]) -> Dict[str, FusedAdam]:
    """``optimizer_specs`` is TrainerSpec.optimizers (fruit_nerf_config.py): {group: {"optimizer": {type, lr, eps},
    "scheduler": None | {"lr_final", "max_steps"}}}."""
    out = {}
    for name, params in param_groups.items():
        if not params:
            continue
        spec = optimizer_specs[name]
        o, s = spec["optimizer"], spec.get("scheduler")
        sched = ExponentialDecay(o["lr"], s["lr_final"], s["max_steps"]) if s else None
        out[name] = FusedAdam(params, lr=o["lr"], eps=o["eps"], kind=o["type"], scheduler=sched)
    return out
