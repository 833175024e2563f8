"""Export-side datamanager pieces of the plugin surface (fruit_nerf/data/fruit_datamanager.py).

In scope: ``get_corners_of_aabb`` (42-68), ``sample_surface_points`` (71-121),
``FruitDataManager.setup_inference`` (157-172) and ``next_sample_volume`` (199-204) -- they define
the export ray grid.  Image loading / pixel sampling (the rest of the reference datamanager) is
host-side I/O outside the hot path and is not rebuilt here.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, Optional, Tuple, Type, Union

import torch
from torch import nn

from ..compat import InstantiateConfig, RayBundle
from ..components.ray_generators import OrthographicRayGenerator


def get_corners_of_aabb(aabb, device):
    """8 corners of an AABB given as (min xyz, max xyz) -- fruit_datamanager.py:42-68."""
    mn, mx = aabb[0], aabb[1]
    return torch.asarray(
        [
            [mn[0], mn[1], mn[2]],
            [mx[0], mn[1], mn[2]],
            [mn[0], mx[1], mn[2]],
            [mx[0], mx[1], mn[2]],
            [mn[0], mn[1], mx[2]],
            [mx[0], mn[1], mx[2]],
            [mn[0], mx[1], mx[2]],
            [mx[0], mx[1], mx[2]],
        ],
        dtype=torch.float32,
        device=device,
    )


def sample_surface_points(aabb, n, device, noise=False):
    """Regular grid on the z = min face of the box spanned by the 8 corners ``aabb`` and the vector
    to the opposite face (fruit_datamanager.py:71-121): int(dx/dz*n) x int(dy/dz*n) points,
    meshgrid 'ij' flattened x-major."""
    c1, c2, c3 = aabb[0], aabb[1], aabb[2]
    dxyz = torch.abs(torch.max(aabb, dim=0).values - torch.min(aabb, dim=0).values)
    const_axis = int(torch.argmax(torch.logical_and((c1 == c2), (c2 == c3)).to(int)))
    ax = torch.argmax(torch.abs(c1 - c2))
    x = torch.linspace(float(c1[ax]), float(c2[ax]), int(dxyz[0] / dxyz[const_axis] * n), dtype=torch.float32, device=device)
    ay = torch.argmax(torch.abs(c1 - c3))
    y = torch.linspace(float(c1[ay]), float(c3[ay]), int(dxyz[1] / dxyz[const_axis] * n), dtype=torch.float32, device=device)
    xx, yy = torch.meshgrid(x, y, indexing="ij")
    pts = torch.column_stack((xx.flatten(), yy.flatten(), torch.full_like(xx.flatten(), float(c3[const_axis]))))
    c4 = aabb[-1]
    plane_vector = torch.asarray(
        [[0, 0, float(torch.sign(c4[const_axis]) * torch.abs(c1[const_axis]) + torch.abs(c4[const_axis]))]],
        dtype=torch.float32, device=device)
    return pts.clone(), plane_vector


@dataclass
class FruitDataManagerConfig(InstantiateConfig):
    _target: Type = field(default_factory=lambda: FruitDataManager)
    train_num_rays_per_batch: int = 4096
    eval_num_rays_per_batch: int = 4096


class FruitDataManager(nn.Module):
    """Export-side subset of fruit_nerf.data.fruit_datamanager.FruitDataManager."""

    config: FruitDataManagerConfig

    def __init__(self, config: FruitDataManagerConfig, device: Union[torch.device, str] = "cpu", test_mode: str = "val",
                 world_size: int = 1, local_rank: int = 0, **kwargs):
        super().__init__()
        self.config = config
        self.device = device
        self.test_mode = test_mode
        self.world_size, self.local_rank = world_size, local_rank
        self.train_count = 0
        self.eval_count = 0
        self.train_dataset = kwargs.get("train_dataset")
        self.orthographic_ray_generator: Optional[OrthographicRayGenerator] = None

    def setup_inference(self, aabb, num_points) -> int:
        """fruit_datamanager.py:157-172: ray grid for the uniform volume; returns the ray count."""
        corners = get_corners_of_aabb(aabb=aabb, device=self.device)
        surface_points, plane_vector = sample_surface_points(corners, n=num_points, device=self.device, noise=False)
        self.orthographic_ray_generator = OrthographicRayGenerator(
            surface_points=surface_points, plane_normal=plane_vector, ray_batch_size=self.config.eval_num_rays_per_batch,
            device=self.device, aabb=aabb)
        return surface_points.shape[0]

    def next_sample_volume(self, step: int) -> Tuple[RayBundle, Optional[Dict]]:
        """fruit_datamanager.py:199-204."""
        self.train_count += 1
        return self.orthographic_ray_generator(count=self.train_count), None

    def get_param_groups(self):
        return {}

    def get_training_callbacks(self, attrs):
        return []
