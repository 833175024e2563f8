"""``ns-export-semantics semantic-pointcloud`` (fruit_nerf/scripts/exporter.py:54-144).

Same dataclass fields and defaults; ``main`` takes an already-built pipeline (nerfstudio's
``eval_setup`` -- checkpoint discovery from a YAML config -- is control plane and out of scope).
"""
from __future__ import annotations

import json
import os
from dataclasses import dataclass
from pathlib import Path
from typing import Optional, Tuple

from ..export.exporter_utils import sample_volume, write_ply


@dataclass
class Exporter:
    load_config: Optional[Path]
    output_dir: Path


@dataclass
class ExportSemanticPointCloud(Exporter):
    """exporter.py:64-77."""

    use_bounding_box: bool = True
    bounding_box_min: Tuple[float, float, float] = (-1, -1, -1)
    bounding_box_max: Tuple[float, float, float] = (1, 1, 1)
    num_rays_per_batch: int = 32768
    num_points_per_side: int = 1000

    def main(self, pipeline=None, config=None, transform_json: Optional[dict] = None) -> dict:
        """exporter.py:80-121 with ``pipeline`` supplied by the caller (test_mode='export')."""
        if pipeline is None:
            raise NotImplementedError("pass a FruitPipeline built with test_mode='export' (nerfstudio eval_setup is out of scope)")
        self.output_dir = Path(self.output_dir)
        self.output_dir.mkdir(parents=True, exist_ok=True)
        pipeline.datamanager.config.eval_num_rays_per_batch = self.num_rays_per_batch
        pipeline.model.setup_inference(render_rgb=True, num_inference_samples=self.num_points_per_side)
        num_points = pipeline.datamanager.setup_inference(num_points=self.num_points_per_side,
                                                          aabb=(self.bounding_box_min, self.bounding_box_max))
        if transform_json is None and self.load_config is not None:
            with open(Path(self.load_config).parent / "dataparser_transforms.json", "r") as fp:
                transform_json = json.load(fp)
        pcds = sample_volume(pipeline=pipeline, num_points=num_points, output_dir=self.output_dir, config=config,
                             transform_json=transform_json, world_size=getattr(pipeline, "world_size", 1),
                             rank=getattr(pipeline, "local_rank", 0))
        for name, pcd in pcds.items():
            path = pcd["path"] or str(self.output_dir / f"{name}.ply")
            os.makedirs(os.path.dirname(path), exist_ok=True)
            write_ply(path, pcd["points"], pcd["colors"])
            pcd["path"] = path
        return pcds


def entrypoint():
    """``ns-export-semantics`` (exporter.py:124-144 upstream parses the sub-command with tyro and calls ``main``).
    Building the pipeline from a nerfstudio YAML config (``eval_setup``) is control plane and not rebuilt: use
    ``ExportSemanticPointCloud(load_config=None, output_dir=...).main(pipeline=...)`` from Python."""
    raise SystemExit("ns-export-semantics: construct a FruitPipeline(test_mode='export') and call "
                     "ExportSemanticPointCloud(...).main(pipeline=pipeline); nerfstudio's eval_setup is out of scope")
