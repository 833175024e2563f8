"""HashMLPDensityField -- parameter holder for nerfstudio's proposal network
(fruit_nerf/fruit_nerf.py:111-128 builds two of them from ``proposal_net_args_list``).

State-dict keys follow nerfstudio's torch path: ``encoding.hash_table``, ``mlp_base.0.hash_table`` (alias),
``mlp_base.1.layers.{0,1}.{weight,bias}``, buffers ``aabb / max_res / num_levels / log2_hashmap_size``.
The arithmetic (density + weights, backward) runs in fnr_proposal.cu through ``ops.proposal_weights``.
"""
from __future__ import annotations

from typing import Optional

import torch
from torch import Tensor, nn

from . import _lib as L
from . import ops
from .fruit_field import MLP, HashEncoding


class HashMLPDensityField(nn.Module):
    def __init__(self, aabb: Tensor, num_layers: int = 2, hidden_dim: int = 64, spatial_distortion: Optional[nn.Module] = None,
                 use_linear: bool = False, num_levels: int = 8, max_res: int = 1024, base_res: int = 16, log2_hashmap_size: int = 18,
                 features_per_level: int = 2, implementation: str = "b200") -> None:
        super().__init__()
        if use_linear or num_layers != 2 or features_per_level != 2:
            raise NotImplementedError("proposal networks are hash grid -> Linear -> ReLU -> Linear (FruitNeRF's proposal_net_args_list)")
        self.register_buffer("aabb", torch.as_tensor(aabb, dtype=torch.float32))
        self.spatial_distortion = spatial_distortion
        self.use_linear = use_linear
        self.register_buffer("max_res", torch.tensor(max_res))
        self.register_buffer("num_levels", torch.tensor(num_levels))
        self.register_buffer("log2_hashmap_size", torch.tensor(log2_hashmap_size))
        self.encoding = HashEncoding(num_levels=num_levels, min_res=base_res, max_res=max_res, log2_hashmap_size=log2_hashmap_size,
                                     features_per_level=features_per_level)
        network = MLP(in_dim=self.encoding.get_out_dim(), num_layers=num_layers, layer_width=hidden_dim, out_dim=1)
        self.mlp_base = torch.nn.Sequential(self.encoding, network)
        self.hidden_dim = hidden_dim

    def kernel_shape(self) -> ops.DensityShape:
        key = (self.aabb._version, self.aabb.data_ptr())
        cached = getattr(self, "_shape_cache", None)
        if cached is None or cached[0] != key:
            e = self.encoding
            shape = ops.DensityShape(num_levels=e.num_levels, log2_hashmap_size=e.log2_hashmap_size, hidden_dim=self.hidden_dim,
                                     scalings=[float(v) for v in e.scalings], aabb=[float(v) for v in self.aabb.reshape(-1)])
            object.__setattr__(self, "_shape_cache", (key, shape))
            cached = self._shape_cache
        return cached[1]

    def kernel_params(self):
        net = self.mlp_base[1]
        return [self.encoding.hash_table, net.layers[0].weight, net.layers[0].bias, net.layers[1].weight, net.layers[1].bias]

    def position_mode(self) -> int:
        return L.FNR_POS_CONTRACT if self.spatial_distortion is not None else L.FNR_POS_AABB

    def weights(self, origins: Tensor, directions: Tensor, starts: Tensor, ends: Tensor) -> Tensor:
        """density_fn(frustum midpoints) -> RaySamples.get_weights, fused.  starts/ends [R,S] -> weights [R,S]."""
        return ops.proposal_weights(self.kernel_shape(), self.kernel_params(), origins, directions, starts, ends, self.position_mode())

    def density_fn(self, *a, **k):  # the sampler recognises this bound method and calls ``weights`` on its owner
        raise RuntimeError("HashMLPDensityField.density_fn is evaluated inside the fused proposal kernel (use .weights)")
