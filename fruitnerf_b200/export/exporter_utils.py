"""Uniform-volume point-cloud export (fruit_nerf/export/exporter_utils.py:47-258).

The reference loops ``model(ray_bundle)`` -> dense [B,S,*] outputs -> three boolean-mask gathers ->
``.cpu()`` per batch.  Here each batch is ONE kernel that evaluates the field and compacts the three
point sets on the device (ops.export_batch); a single device->host copy happens at the end.  PLY
files are written directly (open3d is not needed for a binary point-cloud PLY).
"""
from __future__ import annotations

import pathlib
from typing import Dict, Optional

import numpy as np
import torch

from .. import ops

SET_NAMES = ("semantic_colormap", "semantic", "density")  # exporter_utils.py:193-256


def export_slab(num_rays: int, world_size: int, rank: int):
    """Contiguous slab [lo, hi) of the export rays owned by ``rank``: the rays of the face grid are independent
    (components/ray_generators.py:52-64), so the volume shards with no data-path collective."""
    per = (num_rays + world_size - 1) // world_size
    lo = min(rank * per, num_rays)
    return lo, min(lo + per, num_rays)


def merge_export_shards(local: Dict[str, tuple], world_size: int) -> Dict[str, tuple]:
    """All ranks contribute (rows [n,7], keys [n]) per set; everyone receives the concatenation (the surviving point
    lists are small next to the volume).  Keys are global point indices, so sorting by key restores the reference order."""
    if world_size <= 1:
        return local
    import torch.distributed as dist

    gathered = [None] * world_size
    dist.all_gather_object(gathered, {k: (r.cpu().numpy(), q.cpu().numpy()) for k, (r, q) in local.items()})
    out = {}
    for name in local:
        rows = np.concatenate([g[name][0] for g in gathered], axis=0)
        keys = np.concatenate([g[name][1] for g in gathered], axis=0)
        out[name] = (torch.from_numpy(rows), torch.from_numpy(keys))
    return out


def sample_volume(pipeline, num_points: int, output_dir: Optional[pathlib.Path] = None, config=None, transform_json: dict = None,
                  capacity: Optional[int] = None, world_size: int = 1, rank: int = 0) -> Dict[str, Dict]:
    """Returns {name: {'points' [N,3] float64, 'colors' [N,3] float64, 'alpha' [N], 'path'}} for the
    three clouds.  ``num_points`` is the number of export rays (datamanager.setup_inference).  With ``world_size`` > 1
    each rank evaluates its slab of rays and the (small) selected point lists are gathered at the end."""
    model = pipeline.model
    dm = pipeline.datamanager
    dev = next(model.parameters()).device
    S = model.num_inference_samples
    lo, hi = export_slab(num_points, world_size, rank)
    total = (hi - lo) * S
    capacity = capacity or max(1, min(total, 1 << 24))
    buffers = ops.ExportBuffers(capacity=capacity, device=dev)
    gen = dm.orthographic_ray_generator
    B = gen.ray_batch_size
    with torch.no_grad():
        if world_size <= 1:
            done = 0
            while done < num_points:
                ray_bundle, _ = dm.next_sample_volume(0)
                n = ray_bundle.origins.shape[0]
                if n == 0:
                    break
                model.get_export_outputs(ray_bundle.to(dev) if hasattr(ray_bundle, "to") else ray_bundle, buffers=buffers,
                                         point_base=done * S, dense=False)
                done += n
        else:  # this rank's slab, same batch size; point keys stay global
            for start in range(lo, hi, B):
                full = gen(count=start // B + 1) if start % B == 0 and start + B <= hi else None
                if full is None:
                    from ..compat import RayBundle

                    pts = gen.surface_points[start:min(start + B, hi)]
                    n = pts.shape[0]
                    full = RayBundle(origins=pts, directions=gen.surface_normal.repeat(n, 1).to(dev), pixel_area=torch.zeros(n, 1, device=dev),
                                     nears=torch.zeros(n, 1, device=dev), fars=torch.ones(n, 1, device=dev) * gen.surface_vector_norm)
                model.get_export_outputs(full.to(dev), buffers=buffers, point_base=start * S, dense=False)
    counts = buffers.counts.cpu().tolist()  # the single D2H sync
    if max(counts) > capacity:
        raise RuntimeError(f"export capacity {capacity} too small for {max(counts)} selected points; pass capacity=")
    scale = 1.0
    if transform_json is not None:
        scale = 2.0 / float(transform_json["scale"])  # pcd.scale(1/scale) then pcd.scale(2) (exporter_utils.py:190-191)
    local = {name: (buffers.rows[k][: counts[k]], buffers.keys[k][: counts[k]]) for k, name in enumerate(SET_NAMES)}
    merged = merge_export_shards(local, world_size)
    out = {}
    for k, name in enumerate(SET_NAMES):
        rows, keys = merged[name]
        order = torch.argsort(keys)  # reference order = batch-major point order
        rows = rows[order].double().cpu().numpy()
        colors = rows[:, 3:6].copy()
        if name != "semantic_colormap" and colors.shape[0] != 0:
            full = rows[:, 3:7]
            colors = (full / full.max())[:, :3]  # exporter_utils.py:203, 228: normalise by the max over rgb+alpha
        path = None
        if output_dir is not None and config is not None and getattr(config, "load_dir", None) is not None:
            parts = pathlib.Path(config.load_dir).parts  # upstream: outputs/<experiment>/<method>/<timestamp>/nerfstudio_models
            sub = parts[-3] if len(parts) >= 3 else ""
            path = str(pathlib.Path(output_dir) / sub / f"{name}.ply")
        out[name] = {"points": rows[:, :3] * scale, "colors": colors, "alpha": rows[:, 6], "path": path}
    return out


def write_ply(path, points: np.ndarray, colors: np.ndarray) -> None:
    """Binary little-endian PLY with double xyz and uchar rgb -- what open3d's write_point_cloud
    produces for a coloured cloud (fruit_nerf/scripts/exporter.py:116-119)."""
    pts = np.asarray(points, dtype="<f8")
    col = np.clip(np.asarray(colors, dtype=np.float64) * 255.0, 0, 255).astype(np.uint8)
    n = pts.shape[0]
    header = (
        "ply\nformat binary_little_endian 1.0\ncomment fruitnerf_b200 export\n"
        f"element vertex {n}\nproperty double x\nproperty double y\nproperty double z\n"
        "property uchar red\nproperty uchar green\nproperty uchar blue\nend_header\n"
    )
    rec = np.empty(n, dtype=[("x", "<f8"), ("y", "<f8"), ("z", "<f8"), ("r", "u1"), ("g", "u1"), ("b", "u1")])
    rec["x"], rec["y"], rec["z"] = pts[:, 0], pts[:, 1], pts[:, 2]
    rec["r"], rec["g"], rec["b"] = col[:, 0], col[:, 1], col[:, 2]
    pathlib.Path(path).parent.mkdir(parents=True, exist_ok=True)
    with open(path, "wb") as f:
        f.write(header.encode("ascii"))
        f.write(rec.tobytes())
