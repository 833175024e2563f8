"""OrthographicRayGenerator -- parallel export rays from one face of an AABB
(fruit_nerf/components/ray_generators.py:24-66).  Host-side indexing only."""
from __future__ import annotations

import torch
from torch import Tensor, nn

from ..compat import RayBundle


class OrthographicRayGenerator(nn.Module):
    def __init__(self, surface_points: Tensor, plane_normal: Tensor, ray_batch_size: int, device, aabb) -> None:
        super().__init__()
        self.surface_points = surface_points
        self.surface_normal = torch.nn.functional.normalize(plane_normal).to(device)
        self.surface_vector_norm = torch.linalg.norm(plane_normal).to(device)
        self.ray_batch_size = ray_batch_size
        self.device = device
        self.aabb = aabb

    def forward(self, count: int) -> RayBundle:
        """``count`` is 1-based (ray_generators.py:52-53); the last batch is short."""
        start = self.ray_batch_size * (count - 1)
        end = self.ray_batch_size * count
        if self.ray_batch_size * count >= self.surface_points.shape[0]:
            end = self.surface_points.shape[0]
        pts = self.surface_points[start:end]
        n = pts.shape[0]
        return RayBundle(
            origins=pts,
            directions=self.surface_normal.repeat(n, 1).to(self.device),
            pixel_area=torch.zeros(n, 1).to(self.device),
            nears=torch.zeros(n, 1).to(self.device),
            fars=torch.ones(n, 1).to(self.device) * self.surface_vector_norm,
        )
