"""Host side of the NVLS gradient exchange (fnr_nvls.cu): symmetric allocation + rendezvous through
``torch.distributed._symmetric_memory`` (plumbing: device memory, handle exchange over the process group), the reduction
itself is the library's own kernel (``fnr_nvls_allreduce_mean``: multimem.ld_reduce / multimem.st)."""
from __future__ import annotations

import ctypes as C

import torch
import torch.distributed as dist

from . import _lib as L
from .grad_exchange import GradientExchange

SLOT_BASE = 1024  # leave the first slots of the signal pad to torch's own symmetric-memory ops


class NvlsExchange(GradientExchange):
    """Owns the symmetric flat gradient buffer (``self.flat``, fp32, ``numel`` elements rounded up to a multiple of
    8 * world): hand it to the backward (``ops.render(..., flat_grad=exchange.flat)``) so that gradients are accumulated
    straight into multicast-mapped memory, then call the exchange."""

    def __init__(self, numel: int, world: int, device, bf16_wire: bool = False, group=None):
        import torch.distributed._symmetric_memory as symm_mem

        group = group or dist.group.WORLD
        quantum = 8 * world
        n = (numel + quantum - 1) // quantum * quantum
        total = n + (n // 2 if bf16_wire else 0)  # [fp32 gradients | bf16 staging]
        self._region = symm_mem.empty(total, dtype=torch.float32, device=device)
        self._region.zero_()
        self._hdl = symm_mem.rendezvous(self._region, group)
        h = self._hdl
        if not getattr(h, "has_multicast_support", False) or not h.multicast_ptr:
            raise RuntimeError("symmetric memory without multicast support on this platform")
        super().__init__(self._region[:numel], world, "nvls_bf16" if bf16_wire else "nvls")
        self.n = n
        self.bf16_wire = bf16_wire
        base_local = int(h.buffer_ptrs[h.rank])
        off = self._region.data_ptr() - base_local
        d = L.NvlsDesc()
        d.multicast_ptr = int(h.multicast_ptr) + off
        d.local_ptr = self._region.data_ptr()
        if bf16_wire:
            d.multicast_bf16 = int(h.multicast_ptr) + off + 4 * n
            d.local_bf16 = self._region.data_ptr() + 4 * n
            self._counter = torch.zeros(1, dtype=torch.int32, device=device)
            d.grid_counter = self._counter.data_ptr()
        d.signal_pads = int(h.signal_pad_ptrs_dev)
        d.rank, d.world_size = int(h.rank), int(h.world_size)
        d.signal_slots = int(h.signal_pad_size) // 4
        d.signal_slot_base = SLOT_BASE
        self._desc = d
        self._device = torch.device(device)
        torch.cuda.synchronize(self._device)
        dist.barrier(group)  # every rank has zeroed its region and its pad before anybody launches

    def __call__(self) -> None:
        stream = torch.cuda.current_stream(self._device).cuda_stream
        L.check(L.load().fnr_nvls_allreduce_mean(C.byref(self._desc), self.n, 1 if self.bf16_wire else 0, stream))

    def describe(self) -> dict:
        d = super().describe()
        d["note"] = ("one kernel: multimem.ld_reduce over this rank's 1/N slice, x 1/N, multimem.st back to all ranks"
                     + ("; bf16 on the wire, fp32 accumulation in the switch" if self.bf16_wire else "; fp32 on the wire"))
        return d
