"""Ray samplers of the plugin surface (host side: they only produce bin edges).

UniformSamplerWithNoise  fruit_nerf/components/ray_samplers.py:31-104 (export sampler)
UniformLinDispPiecewiseSampler  nerfstudio's default ``initial_sampler`` of ProposalNetworkSampler
                         (fruit_nerf/fruit_nerf.py:151-158, initial_sampler=None)
"""
from __future__ import annotations

from typing import Callable, Optional

import torch
from torch import Tensor, nn

from ..compat import RayBundle, RaySamples


class SpacedSampler(nn.Module):
    """nerfstudio SpacedSampler: bins uniform in ``spacing_fn`` space, stratified jitter in training."""

    def __init__(self, spacing_fn: Callable, spacing_fn_inv: Callable, num_samples: Optional[int] = None,
                 train_stratified: bool = True, single_jitter: bool = False) -> None:
        super().__init__()
        self.num_samples = num_samples
        self.train_stratified = train_stratified
        self.single_jitter = single_jitter
        self.spacing_fn = spacing_fn
        self.spacing_fn_inv = spacing_fn_inv

    def forward(self, ray_bundle: Optional[RayBundle] = None, num_samples: Optional[int] = None) -> RaySamples:
        return self.generate_ray_samples(ray_bundle, num_samples)

    def spacing_bins(self, num_rays: int, num_samples: int, device) -> Tensor:
        bins = self._base_bins(num_samples, device)
        if self.train_stratified and self.training:
            if self.single_jitter:
                t_rand = torch.rand((num_rays, 1), dtype=bins.dtype, device=bins.device)
            else:
                t_rand = torch.rand((num_rays, num_samples + 1), dtype=bins.dtype, device=bins.device)
            centers = (bins[..., 1:] + bins[..., :-1]) / 2.0
            upper = torch.cat([centers, bins[..., -1:]], -1)
            lower = torch.cat([bins[..., :1], centers], -1)
            bins = lower + (upper - lower) * t_rand
        # nerfstudio's TensorDataclass broadcasts every RaySamples field to the ray batch shape
        return bins.expand(num_rays, num_samples + 1)

    native_mode: Optional[int] = None  # FNR_SPACING_* of the subclasses whose spacing functions the native kernel knows

    def _base_bins(self, num_samples: int, device) -> Tensor:
        # the reference builds the bins on the host and moves them (ray_samplers.py:75); cached per (size, device) so that
        # no host->device copy happens per call (and none inside a CUDA-graph capture)
        cache = self.__dict__.setdefault("_bins_cache", {})
        key = (num_samples, str(device))
        if key not in cache:
            cache[key] = torch.linspace(0.0, 1.0, num_samples + 1).to(device)[None, ...]
        return cache[key]

    def generate_ray_samples(self, ray_bundle: Optional[RayBundle] = None, num_samples: Optional[int] = None) -> RaySamples:
        assert ray_bundle is not None
        assert ray_bundle.nears is not None
        assert ray_bundle.fars is not None
        num_samples = num_samples or self.num_samples
        assert num_samples is not None
        num_rays = ray_bundle.origins.shape[0]
        nears, fars = ray_bundle.nears, ray_bundle.fars

        def spacing_to_euclidean_fn(x):
            s_near, s_far = (self.spacing_fn(v) for v in (nears, fars))
            return self.spacing_fn_inv(x * s_far + (1 - x) * s_near)

        dev = ray_bundle.origins.device
        if dev.type == "cuda" and self.native_mode is not None and ray_bundle.origins.dim() == 2:
            # one launch (fnr_spaced_bins) instead of ~15 elementwise kernels
            from .. import ops

            t_rand = None
            if self.train_stratified and self.training:
                t_rand = torch.rand((num_rays, 1) if self.single_jitter else (num_rays, num_samples + 1), dtype=torch.float32, device=dev)
            bins, starts, ends = ops.spaced_bins(self._base_bins(num_samples, dev), t_rand, nears, fars, num_samples, self.native_mode)
            return ray_bundle.get_ray_samples(bin_starts=starts[..., None], bin_ends=ends[..., None], spacing_starts=bins[..., :-1, None],
                                              spacing_ends=bins[..., 1:, None], spacing_to_euclidean_fn=spacing_to_euclidean_fn)
        bins = self.spacing_bins(num_rays, num_samples, dev)
        euclidean_bins = spacing_to_euclidean_fn(bins)
        return ray_bundle.get_ray_samples(
            bin_starts=euclidean_bins[..., :-1, None],
            bin_ends=euclidean_bins[..., 1:, None],
            spacing_starts=bins[..., :-1, None],
            spacing_ends=bins[..., 1:, None],
            spacing_to_euclidean_fn=spacing_to_euclidean_fn,
        )


class UniformSamplerWithNoise(SpacedSampler):
    """Linear spacing; jitter only when ``self.training`` (export runs in eval mode)."""

    native_mode = 0  # FNR_SPACING_UNIFORM

    def __init__(self, num_samples: Optional[int] = None, train_stratified=True, single_jitter=False) -> None:
        super().__init__(spacing_fn=lambda x: x, spacing_fn_inv=lambda x: x, num_samples=num_samples,
                         train_stratified=train_stratified, single_jitter=single_jitter)


class UniformLinDispPiecewiseSampler(SpacedSampler):
    """nerfstudio: linear up to distance 1, then linear in disparity."""

    native_mode = 1  # FNR_SPACING_LINDISP_PIECEWISE

    def __init__(self, num_samples: Optional[int] = None, train_stratified=True, single_jitter=False) -> None:
        super().__init__(
            spacing_fn=lambda x: torch.where(x < 1, x / 2, 1 - 1 / (2 * x)),
            spacing_fn_inv=lambda x: torch.where(x < 0.5, 2 * x, 1 / (2 - 2 * x)),
            num_samples=num_samples, train_stratified=train_stratified, single_jitter=single_jitter)
