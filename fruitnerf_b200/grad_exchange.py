"""Gradient exchange of the data-parallel path (the reference wraps the model in DDP,
fruit_nerf/fruit_pipeline.py:116-118): mean over ranks of ONE flat fp32 gradient buffer -- the backward kernels
accumulate every parameter gradient into it, so the whole exchange is a single collective.

``make_gradient_exchange(flat, world, kind)`` returns a callable that enqueues the exchange on the current stream:

* ``"nccl"``  -- ``torch.distributed.all_reduce(AVG)`` (NCCL over NVLink / NVSwitch; NVLS when NCCL selects it).
* ``"nvls"`` / ``"nvls_bf16"`` -- the library's own one-kernel all-reduce over NVSwitch multicast memory
  (``multimem.ld_reduce`` / ``multimem.st``, fnr_nvls.cu): every rank reduces 1/N of the buffer in the switch and
  broadcasts the mean back, fp32 or bf16 on the wire.  Needs symmetric (multicast-capable) memory; when the platform
  cannot provide it the factory falls back to NCCL and says so in ``describe()``.
* ``"auto"``  -- NVLS when available, else NCCL.
"""
from __future__ import annotations

from typing import Callable, Optional

import torch
import torch.distributed as dist


NVLS_AUTO_MIN_WORLD = 4  # "auto" uses the multimem kernel from this world size on


class GradientExchange:
    def __init__(self, flat: torch.Tensor, world: int, kind: str, note: str = ""):
        self.flat, self.world, self.kind, self.note = flat, world, kind, note

    def __call__(self) -> None:
        raise NotImplementedError

    def describe(self) -> dict:
        return {"kind": self.kind, "bytes": self.flat.numel() * self.flat.element_size(), "note": self.note}


class NcclExchange(GradientExchange):
    def __call__(self) -> None:
        dist.all_reduce(self.flat, op=dist.ReduceOp.AVG if self.flat.is_cuda else dist.ReduceOp.SUM)
        if not self.flat.is_cuda:  # gloo has no AVG (CPU tests of the host logic)
            self.flat.div_(self.world)


def make_gradient_exchange(numel: int, world: int, device, kind: str = "auto") -> Optional[GradientExchange]:
    """The exchange OWNS the flat gradient buffer (``.flat``, fp32, >= numel elements, persistent): pass it to the backward
    (``ops.render(..., flat_grad=exchange.flat)`` / ``GraphedTrainStep(..., flat_grad=...)``)."""
    if world <= 1:
        return None
    device = torch.device(device)
    note = ""
    if kind == "auto" and world < NVLS_AUTO_MIN_WORLD:
        kind = "nccl"  # measured at N = 2 (tools/r2/nvls_test.py): NCCL 154 us vs 201-213 us for 67 MB -- with two ranks in-switch reduction moves
                       # MORE bytes per GPU than a direct exchange (every operand, the local one included, travels to the switch)
    if kind in ("auto", "nvls", "nvls_bf16") and device.type == "cuda":
        try:
            from .nvls import NvlsExchange

            ok = torch.zeros(1, device=device)
            try:
                ex_obj = NvlsExchange(numel, world, device, bf16_wire=(kind == "nvls_bf16"))
            except Exception as ex:  # noqa: BLE001 -- no multicast support on this platform / torch build
                ex_obj, note = None, f"NVLS path unavailable ({type(ex).__name__}: {str(ex)[:120]}); NCCL all-reduce"
                ok.fill_(1.0)
            dist.all_reduce(ok)  # all ranks take the same path
            if ex_obj is not None and float(ok) == 0.0:
                return ex_obj
            if kind != "auto":
                raise RuntimeError(note or "NVLS path unavailable on another rank")
            note = note or "NVLS path unavailable on another rank; NCCL all-reduce"
        except ImportError as ex:
            if kind != "auto":
                raise
            note = f"NVLS path unavailable ({ex}); NCCL all-reduce"
    flat = torch.zeros(numel, dtype=torch.float32, device=device)
    return NcclExchange(flat, world, "nccl", note=note)
