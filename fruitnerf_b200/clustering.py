"""Fruit counting on the exported semantic point cloud: stages 1-2 of the reference's clustering
(clustering/clustering_base.py:138-143 radius-outlier removal + voxel down-sampling, :183-207 DBSCAN,
:209-259 merging of cluster centres closer than ``cluster_merge_distance``).  The template-matching split of
oversized clusters (stage 3, :261-) needs open3d / alphashape and the LFS fruit templates, none of which exist
offline; it is not restated.  CPU post-processing (numpy / scikit-learn), outside the GPU hot path.
"""
from __future__ import annotations

from typing import Dict

import numpy as np
from sklearn.cluster import DBSCAN
from sklearn.neighbors import NearestNeighbors


def remove_radius_outliers(points: np.ndarray, nb_points: int, radius: float) -> np.ndarray:
    """open3d remove_radius_outlier: keep points with at least ``nb_points`` neighbours within ``radius``."""
    if points.shape[0] == 0:
        return points
    nn = NearestNeighbors(radius=radius).fit(points)
    counts = np.array([len(ix) for ix in nn.radius_neighbors(points, return_distance=False)])
    return points[counts - 1 >= nb_points]  # the query point itself is excluded, as in open3d


def voxel_down_sample(points: np.ndarray, voxel: float) -> np.ndarray:
    """open3d voxel_down_sample: one point (the mean) per occupied voxel."""
    if points.shape[0] == 0 or voxel <= 0:
        return points
    keys = np.floor((points - points.min(axis=0)) / voxel).astype(np.int64)
    _, inv, cnt = np.unique(keys, axis=0, return_inverse=True, return_counts=True)
    out = np.zeros((cnt.shape[0], 3))
    np.add.at(out, inv.reshape(-1), points)
    return out / cnt[:, None]


def count_fruits(points: np.ndarray, eps: float, min_samples: int, cluster_merge_distance: float, down_sample: float = 0.0,
                 remove_outliers_nb_points: int = 0, remove_outliers_radius: float = 0.0) -> Dict:
    """Returns {'count', 'count_before_merge', 'centers' [count,3], 'num_points'}."""
    pts = np.asarray(points, dtype=np.float64).reshape(-1, 3)
    if remove_outliers_nb_points > 0 and remove_outliers_radius > 0:
        pts = remove_radius_outliers(pts, remove_outliers_nb_points, remove_outliers_radius)
    pts = voxel_down_sample(pts, down_sample)
    if pts.shape[0] == 0:
        return {"count": 0, "count_before_merge": 0, "centers": np.zeros((0, 3)), "num_points": 0}
    labels = DBSCAN(eps=eps, min_samples=min_samples, n_jobs=-1).fit(pts).labels_
    centers, members = [], []
    first_stage = 0
    for lab in np.unique(labels):
        if lab == -1:
            continue
        first_stage += 1
        cluster = pts[labels == lab]
        c = cluster.mean(axis=0)
        if centers:
            d = np.linalg.norm(np.vstack(centers) - c, axis=1)
            j = int(np.argmin(d))
            if d[j] < cluster_merge_distance:  # fuse with the nearest earlier cluster: centre = midpoint of the two
                centers[j] = (members[j].mean(axis=0) + c) / 2
                members[j] = np.vstack([members[j], cluster])
                continue
        centers.append(c)
        members.append(cluster)
    return {"count": len(centers), "count_before_merge": first_stage, "centers": np.vstack(centers) if centers else np.zeros((0, 3)),
            "num_points": int(pts.shape[0])}
