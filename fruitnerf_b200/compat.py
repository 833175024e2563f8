"""Stand-ins for the handful of nerfstudio types the FruitNeRF plugin surface names.

When ``nerfstudio`` is importable the real classes are re-exported and nothing here is used.
It is not installable in this image (no network), so these minimal dataclasses carry the same
field names and the few methods the reference calls on them (file:line cited per method).  They
are containers: no field / sampler / renderer arithmetic lives here.
"""
from __future__ import annotations

from dataclasses import dataclass, field, fields
from enum import Enum
from typing import Any, Callable, Dict, List, Optional, Type

import torch
from torch import Tensor

try:  # pragma: no cover - not available in this image
    from nerfstudio.cameras.rays import Frustums, RayBundle, RaySamples  # type: ignore
    from nerfstudio.data.scene_box import SceneBox  # type: ignore
    from nerfstudio.data.dataparsers.base_dataparser import Semantics  # type: ignore
    from nerfstudio.field_components.field_heads import FieldHeadNames  # type: ignore
    from nerfstudio.configs.base_config import InstantiateConfig  # type: ignore

    HAVE_NERFSTUDIO = True
except Exception:  # noqa: BLE001
    HAVE_NERFSTUDIO = False

    class FieldHeadNames(Enum):
        """nerfstudio.field_components.field_heads.FieldHeadNames (subset + same string values)."""

        RGB = "rgb"
        SH = "sh"
        DENSITY = "density"
        NORMALS = "normals"
        PRED_NORMALS = "pred_normals"
        UNCERTAINTY = "uncertainty"
        BACKGROUND_RGB = "background_rgb"
        TRANSIENT_RGB = "transient_rgb"
        TRANSIENT_DENSITY = "transient_density"
        SEMANTICS = "semantics"

    @dataclass
    class Frustums:
        """nerfstudio.cameras.rays.Frustums: per-sample origins/directions/starts/ends."""

        origins: Tensor  # [..., 3]
        directions: Tensor  # [..., 3]
        starts: Tensor  # [..., 1]
        ends: Tensor  # [..., 1]
        pixel_area: Optional[Tensor] = None
        offsets: Optional[Tensor] = None

        @property
        def shape(self):
            return self.starts.shape[:-1]

        def get_positions(self) -> Tensor:
            # o + d * (start + end) / 2 ; used by the reference for 'point_location' (fruit_nerf.py:259)
            pos = self.origins + self.directions * (self.starts + self.ends) / 2
            if self.offsets is not None:
                pos = pos + self.offsets
            return pos

    @dataclass
    class RaySamples:
        """nerfstudio.cameras.rays.RaySamples (container part only)."""

        frustums: Frustums
        camera_indices: Optional[Tensor] = None  # [..., 1]
        deltas: Optional[Tensor] = None  # [..., 1]
        spacing_starts: Optional[Tensor] = None
        spacing_ends: Optional[Tensor] = None
        spacing_to_euclidean_fn: Optional[Callable] = None
        metadata: Optional[Dict[str, Tensor]] = None
        times: Optional[Tensor] = None

        @property
        def shape(self):
            return self.frustums.shape

    @dataclass
    class RayBundle:
        """nerfstudio.cameras.rays.RayBundle."""

        origins: Tensor  # [R, 3]
        directions: Tensor  # [R, 3]
        pixel_area: Optional[Tensor] = None  # [R, 1]
        camera_indices: Optional[Tensor] = None  # [R, 1]
        nears: Optional[Tensor] = None  # [R, 1]
        fars: Optional[Tensor] = None  # [R, 1]
        metadata: Dict[str, Tensor] = field(default_factory=dict)
        times: Optional[Tensor] = None

        def __len__(self) -> int:
            return int(self.origins.shape[0]) if self.origins.dim() == 2 else int(self.origins.numel() // 3)

        @property
        def shape(self):
            return self.origins.shape[:-1]

        def _apply(self, fn) -> "RayBundle":
            kw = {}
            for f in fields(self):
                v = getattr(self, f.name)
                if torch.is_tensor(v):
                    v = fn(v)
                elif isinstance(v, dict):
                    v = {k: fn(t) if torch.is_tensor(t) else t for k, t in v.items()}
                kw[f.name] = v
            return RayBundle(**kw)

        def to(self, device) -> "RayBundle":
            return self._apply(lambda t: t.to(device))

        def flatten(self) -> "RayBundle":
            return self._apply(lambda t: t.reshape(-1, t.shape[-1]))

        def get_row_major_sliced_ray_bundle(self, start_idx: int, end_idx: int) -> "RayBundle":
            # used by get_outputs_for_camera_ray_bundle (fruit_nerf.py:239)
            return self._apply(lambda t: t.reshape(-1, t.shape[-1])[start_idx:end_idx])

        def get_ray_samples(
            self,
            bin_starts: Tensor,
            bin_ends: Tensor,
            spacing_starts: Optional[Tensor] = None,
            spacing_ends: Optional[Tensor] = None,
            spacing_to_euclidean_fn: Optional[Callable] = None,
        ) -> RaySamples:
            # nerfstudio RayBundle.get_ray_samples (components/ray_samplers.py:96-102)
            deltas = bin_ends - bin_starts
            cam = self.camera_indices[..., None, :] if self.camera_indices is not None else None
            S = bin_starts.shape[-2]
            fr = Frustums(
                origins=self.origins[..., None, :].expand(*self.origins.shape[:-1], S, 3),
                directions=self.directions[..., None, :].expand(*self.directions.shape[:-1], S, 3),
                starts=bin_starts,
                ends=bin_ends,
                pixel_area=self.pixel_area[..., None, :] if self.pixel_area is not None else None,
            )
            return RaySamples(
                frustums=fr,
                camera_indices=cam.expand(*bin_starts.shape[:-1], 1) if cam is not None else None,
                deltas=deltas,
                spacing_starts=spacing_starts,
                spacing_ends=spacing_ends,
                spacing_to_euclidean_fn=spacing_to_euclidean_fn,
                metadata=None,
                times=None,
            )

    @dataclass
    class SceneBox:
        """nerfstudio.data.scene_box.SceneBox."""

        aabb: Tensor  # [2, 3]

    @dataclass
    class Semantics:
        """nerfstudio.data.dataparsers.base_dataparser.Semantics (fruitnerf_dataparser.py:286-290)."""

        filenames: List[Any]
        classes: List[str]
        colors: Tensor
        mask_classes: List[str] = field(default_factory=list)

    @dataclass
    class InstantiateConfig:
        """nerfstudio.configs.base_config.InstantiateConfig: ``setup`` instantiates ``_target``."""

        _target: Type = field(default=object)

        def setup(self, **kwargs) -> Any:
            return self._target(self, **kwargs)


__all__ = [
    "HAVE_NERFSTUDIO",
    "FieldHeadNames",
    "Frustums",
    "RaySamples",
    "RayBundle",
    "SceneBox",
    "Semantics",
    "InstantiateConfig",
]
