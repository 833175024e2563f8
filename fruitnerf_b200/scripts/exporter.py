"""``ns-export-semantics semantic-pointcloud`` (fruit_nerf/scripts/exporter.py:54-144).

Same dataclass fields and defaults as the reference.  ``main`` either takes an already-built pipeline or builds one
from ``load_config`` with ``eval_setup`` below -- the counterpart of nerfstudio's ``eval_setup`` for the run folders
``fruitnerf_b200.trainer.Trainer`` writes (``config.yml`` + ``nerfstudio_models/step-*.ckpt`` +
``dataparser_transforms.json``).
"""
from __future__ import annotations

import argparse
import json
import os
from dataclasses import dataclass
from pathlib import Path
from typing import Optional, Tuple

import torch

from ..export.exporter_utils import sample_volume, write_ply


def eval_setup(config_path, eval_num_rays_per_chunk: Optional[int] = None, test_mode: str = "test", device: Optional[str] = None):
    """nerfstudio.utils.eval_utils.eval_setup for this package's run folders: load ``config.yml`` (the pickled-by-yaml
    TrainerSpec, as nerfstudio does with its TrainerConfig), build the pipeline in ``test_mode``, load the newest
    ``step-*.ckpt`` under ``<run>/nerfstudio_models`` and put the pipeline in eval mode.
    Returns (config, pipeline, checkpoint_path, step); ``config.load_dir`` is set like upstream (exporter.py:111)."""
    import yaml

    config_path = Path(config_path)
    config = yaml.load(config_path.read_text(), Loader=yaml.Loader)
    if eval_num_rays_per_chunk:
        config.pipeline.model.eval_num_rays_per_chunk = eval_num_rays_per_chunk
    config.load_dir = config_path.parent / "nerfstudio_models"
    dev = torch.device(device or ("cuda:0" if torch.cuda.is_available() else "cpu"))
    if dev.type == "cuda":
        torch.cuda.set_device(dev)
    pipeline = config.pipeline.setup(device=dev, test_mode=test_mode)
    pipeline.eval()
    ckpts = sorted(config.load_dir.glob("step-*.ckpt"))
    if not ckpts:
        raise FileNotFoundError(f"no step-*.ckpt under {config.load_dir}")
    path = ckpts[-1]
    state = torch.load(path, map_location=dev, weights_only=False)
    step = int(state["step"])
    pipeline.load_pipeline(state["pipeline"], step)
    return config, pipeline, path, step


@dataclass
class Exporter:
    load_config: Optional[Path]
    output_dir: Path


@dataclass
class ExportSemanticPointCloud(Exporter):
    """exporter.py:64-77 (+ ``stratified_jitter``, see ``FruitModel.get_export_outputs``)."""

    use_bounding_box: bool = True
    bounding_box_min: Tuple[float, float, float] = (-1, -1, -1)
    bounding_box_max: Tuple[float, float, float] = (1, 1, 1)
    num_rays_per_batch: int = 32768
    num_points_per_side: int = 1000
    stratified_jitter: bool = True
    """True = the reference's behaviour: its export sampler is created after ``eval_setup`` and therefore still in
    training mode, so every sample is jittered inside its bin.  False = the deterministic regular grid."""

    def main(self, pipeline=None, config=None, transform_json: Optional[dict] = None) -> dict:
        """exporter.py:80-121; ``pipeline`` may be supplied by the caller (test_mode='export')."""
        if pipeline is None:
            if self.load_config is None:
                raise ValueError("pass load_config (a run folder's config.yml) or an already built pipeline")
            config, pipeline, _, _ = eval_setup(self.load_config, test_mode="export")
        self.output_dir = Path(self.output_dir)
        self.output_dir.mkdir(parents=True, exist_ok=True)
        pipeline.datamanager.config.eval_num_rays_per_batch = self.num_rays_per_batch
        pipeline.model.setup_inference(render_rgb=True, num_inference_samples=self.num_points_per_side)
        pipeline.model.proposal_sampler.train(self.stratified_jitter)
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


def entrypoint(argv=None):
    """``ns-export-semantics semantic-pointcloud --load-config RUN/config.yml --output-dir OUT [...]`` (exporter.py:124-144;
    upstream parses the same dataclass with tyro, which is not available offline -- argparse with the same option names)."""
    ap = argparse.ArgumentParser(prog="ns-export-semantics")
    sub = ap.add_subparsers(dest="command", required=True)
    p = sub.add_parser("semantic-pointcloud", help="uniform-volume export of the fruit point clouds")
    p.add_argument("--load-config", type=Path, required=True)
    p.add_argument("--output-dir", type=Path, required=True)
    p.add_argument("--use-bounding-box", type=lambda s: s.lower() in ("1", "true", "yes"), default=True)
    p.add_argument("--bounding-box-min", type=float, nargs=3, default=(-1, -1, -1))
    p.add_argument("--bounding-box-max", type=float, nargs=3, default=(1, 1, 1))
    p.add_argument("--num-rays-per-batch", type=int, default=32768)
    p.add_argument("--num-points-per-side", type=int, default=1000)
    p.add_argument("--stratified-jitter", type=lambda s: s.lower() in ("1", "true", "yes"), default=True)
    a = ap.parse_args(argv)
    exp = ExportSemanticPointCloud(load_config=a.load_config, output_dir=a.output_dir, use_bounding_box=a.use_bounding_box,
                                   bounding_box_min=tuple(a.bounding_box_min), bounding_box_max=tuple(a.bounding_box_max),
                                   num_rays_per_batch=a.num_rays_per_batch, num_points_per_side=a.num_points_per_side,
                                   stratified_jitter=a.stratified_jitter)
    pcds = exp.main()
    for name, pcd in pcds.items():
        print(f"{name}: {pcd['points'].shape[0]} points -> {pcd['path']}")
    return pcds


if __name__ == "__main__":
    entrypoint()
