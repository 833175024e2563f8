"""ProposalNetworkSampler -- nerfstudio's sampler as FruitModel configures it (fruit_nerf/fruit_nerf.py:151-158):
piecewise linear-in-disparity initial samples -> [proposal density -> weights -> PDF resample] x N -> final bins.
Host logic (level loop, update schedule, anneal) mirrors nerfstudio 0.3.2; the per-ray arithmetic is native
(ops.proposal_weights / ops.pdf_sample)."""
from __future__ import annotations

from typing import Callable, List, Optional, Tuple

import torch
from torch import nn

from .. import ops
from ..compat import RayBundle, RaySamples
from .ray_samplers import UniformLinDispPiecewiseSampler


class ProposalNetworkSampler(nn.Module):
    def __init__(self, num_proposal_samples_per_ray: Tuple[int, ...] = (64,), num_nerf_samples_per_ray: int = 32,
                 num_proposal_network_iterations: int = 2, single_jitter: bool = False, update_sched: Callable = lambda x: 1,
                 initial_sampler=None) -> None:
        super().__init__()
        self.num_proposal_samples_per_ray = num_proposal_samples_per_ray
        self.num_nerf_samples_per_ray = num_nerf_samples_per_ray
        self.num_proposal_network_iterations = num_proposal_network_iterations
        self.update_sched = update_sched
        if self.num_proposal_network_iterations < 1:
            raise ValueError("num_proposal_network_iterations must be >= 1")
        self.initial_sampler = initial_sampler if initial_sampler is not None else UniformLinDispPiecewiseSampler(single_jitter=single_jitter)
        self.single_jitter = single_jitter
        self.histogram_padding = 0.01  # PDFSampler default
        self._anneal = 1.0
        self._anneal_dev: Optional[torch.Tensor] = None  # device copy of the exponent (enable_device_anneal)
        self._anneal_staged = None
        self.force_updated: Optional[bool] = None  # set by a CUDA-graph trainer that picks the schedule branch itself
        self._steps_since_update = 0
        self._step = 0

    def set_anneal(self, anneal: float) -> None:
        self._anneal = anneal
        if self._anneal_staged is not None:
            self._anneal_staged.upload([float(anneal)])

    def enable_device_anneal(self, device) -> None:
        """Keep the annealing exponent in device memory: the kernels read it there, so the schedule can move between
        replays of a captured graph."""
        from ..optim import StagedDeviceBuffer

        self._anneal_staged = StagedDeviceBuffer(1, torch.device(device))
        self._anneal_dev = self._anneal_staged.device_buffer
        self._anneal_staged.upload([float(self._anneal)])

    def wants_update(self) -> bool:
        """nerfstudio ProposalNetworkSampler: run the proposal networks with gradients on this iteration?"""
        return bool(self._steps_since_update > self.update_sched(self._step) or self._step < 10)

    def step_cb(self, step):
        self._step = step
        self._steps_since_update += 1

    def forward(self, ray_bundle: Optional[RayBundle] = None, density_fns: Optional[List[Callable]] = None):
        return self.generate_ray_samples(ray_bundle, density_fns)

    def _pdf_level(self, ray_bundle: RayBundle, prev: RaySamples, weights: torch.Tensor, num_samples: int) -> RaySamples:
        R = weights.shape[0]
        dev = weights.device
        u_rand = None
        if self.training:  # PDFSampler(train_stratified=True)
            u_rand = torch.rand((R, 1) if self.single_jitter else (R, num_samples + 1), device=dev)
        existing = torch.cat([prev.spacing_starts[..., 0], prev.spacing_ends[..., -1:, 0]], dim=-1)
        anneal = self._anneal_dev if self._anneal_dev is not None else self._anneal
        bins, starts, ends = ops.pdf_sample(weights, existing, num_samples, u_rand, anneal, ray_bundle.nears, ray_bundle.fars,
                                            self.histogram_padding)
        return ray_bundle.get_ray_samples(bin_starts=starts[..., None], bin_ends=ends[..., None], spacing_starts=bins[..., :-1, None],
                                          spacing_ends=bins[..., 1:, None], spacing_to_euclidean_fn=prev.spacing_to_euclidean_fn)

    def generate_ray_samples(self, ray_bundle: Optional[RayBundle] = None, density_fns: Optional[List[Callable]] = None):
        assert ray_bundle is not None
        assert density_fns is not None
        weights_list, ray_samples_list = [], []
        n = self.num_proposal_network_iterations
        weights, ray_samples = None, None
        updated = self.wants_update() if self.force_updated is None else self.force_updated
        for i_level in range(n + 1):
            is_prop = i_level < n
            num_samples = self.num_proposal_samples_per_ray[i_level] if is_prop else self.num_nerf_samples_per_ray
            if i_level == 0:
                ray_samples = self.initial_sampler(ray_bundle, num_samples=num_samples)
            else:
                assert weights is not None
                ray_samples = self._pdf_level(ray_bundle, ray_samples, weights[..., 0], num_samples)
            if is_prop:
                net = getattr(density_fns[i_level], "__self__", density_fns[i_level])
                fr = ray_samples.frustums
                args = (ray_bundle.origins, ray_bundle.directions, fr.starts[..., 0], fr.ends[..., 0])
                if updated:
                    w = net.weights(*args)
                else:
                    with torch.no_grad():
                        w = net.weights(*args)
                weights = w[..., None]
                weights_list.append(weights)
                ray_samples_list.append(ray_samples)
        if updated:
            self._steps_since_update = 0
        assert ray_samples is not None
        return ray_samples, weights_list, ray_samples_list
