"""Export-side ray generation: one bundle of parallel rays per batch of face-grid points.

Mirrors the interface of ``fruit_nerf.components.ray_generators.OrthographicRayGenerator`` (reference file lines 24-66):
``OrthographicRayGenerator(surface_points, plane_normal, ray_batch_size, device, aabb)`` and ``generator(count)`` with a
1-based batch counter.  Every ray of the export grid shares one direction (the unit plane normal), starts on the grid
point (near = 0) and ends on the opposite face (far = |plane_normal|); the last batch is simply shorter.  Only indexing
happens here -- the rays are consumed by ``fnr_export_forward``.
"""
from __future__ import annotations

import torch
from torch import Tensor, nn

from ..compat import RayBundle


class OrthographicRayGenerator(nn.Module):
    def __init__(self, surface_points: Tensor, plane_normal: Tensor, ray_batch_size: int, device, aabb) -> None:
        super().__init__()
        normal = plane_normal.reshape(1, 3).to(torch.float32)
        length = torch.linalg.norm(normal)
        self.device = device
        self.aabb = aabb
        self.ray_batch_size = int(ray_batch_size)
        self.surface_points = surface_points
        self.surface_vector_norm = length.to(device)        # distance to the opposite face = far plane of every ray
        self.surface_normal = (normal / length).to(device)  # shared unit direction [1,3]

    def batch_range(self, count: int):
        """Half-open row range of the ``count``-th batch (1-based), clipped to the grid."""
        total = int(self.surface_points.shape[0])
        lo = (count - 1) * self.ray_batch_size
        return min(lo, total), min(lo + self.ray_batch_size, total)

    def forward(self, count: int) -> RayBundle:
        lo, hi = self.batch_range(count)
        origins = self.surface_points[lo:hi]
        n = int(origins.shape[0])
        column = torch.ones((n, 1), dtype=torch.float32, device=self.device)
        return RayBundle(
            origins=origins,
            directions=self.surface_normal.expand(n, 3).contiguous(),
            pixel_area=column * 0.0,
            nears=column * 0.0,
            fars=column * self.surface_vector_norm,
        )
