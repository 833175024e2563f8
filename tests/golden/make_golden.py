"""Generates tests/golden/*.npz from the CPU oracle (run in the build container:
``python -m tests.golden.make_golden``).

PARITY UNPINNED: the reference ships no tests / golden vectors and its arithmetic lives in
nerfstudio 0.3.2 + tinycudann, neither of which is installable here, so these vectors freeze the
ORACLE's output (nerfstudio torch-fallback semantics as restated in oracle/ns_torch.py).  They pin
the oracle against accidental change and travel to the GPU box, where /root/reference and the
oracle's provenance are not available; the `-m gpu` tests compare the CUDA path with them too.
"""
from pathlib import Path

import numpy as np
import torch

from fruitnerf_b200 import synthetic as syn
from oracle import fruit_ref as fr
from oracle import ns_torch as ns

OUT = Path(__file__).resolve().parent


def hash_cases():
    pts = torch.tensor(
        [[0.0, 0.0, 0.0], [0.5, 0.5, 0.5], [0.25, 0.75, 0.125], [0.999999, 0.000001, 0.5], [1.0 / 3, 2.0 / 3, 0.1],
         [0.0625, 0.0625, 0.0625], [0.9, 0.1, 0.7]], dtype=torch.float32)
    res = {}
    for T, max_res in ((17, 2048), (19, 2048), (21, 4096)):
        idx, off = ns.hash_corner_indices(pts, ns.hash_scalings(16, 16, max_res), T)
        res[f"idx_T{T}"] = idx.numpy().astype(np.int64)
        res[f"off_T{T}"] = off.numpy()
        res[f"scalings_{max_res}"] = ns.hash_scalings(16, 16, max_res).numpy()
    res["points"] = pts.numpy()
    np.savez_compressed(OUT / "hash_indices.npz", **res)


def field_cases():
    res = {}
    for name, v in (("small", syn.SMALL), ("big", syn.BIG)):
        sd = syn.field_state(geo=v["geo"], sem_dims=v["sem_dims"], log2_hashmap_size=v["log2_hashmap_size"], num_images=7,
                             table_scale=0.5, weight_gain=1.5)
        spec = fr.FieldSpec(max_res=v["max_res"], log2_hashmap_size=v["log2_hashmap_size"], geo_feat_dim=v["geo"])
        o, d, s, e, cam = syn.ray_batch(32, 24, salt=77, far=3.0, num_images=7)
        for mode, contraction in (("train", True), ("mean", False)):
            f = fr.field_forward(sd, spec, o[:, None, :], d[:, None, :], s[..., None], e[..., None], cam, contraction=contraction,
                                 appearance=mode)
            r = fr.render(f, s[..., None], e[..., None], training=True)
            tag = f"{name}_{mode}"
            for k in ("density", "rgb", "semantics"):
                res[f"{tag}_{k}"] = f[k].numpy()
            for k in ("rgb", "accumulation", "depth", "depth_index", "semantics", "weights"):
                res[f"{tag}_render_{k}"] = r[k].numpy()
    np.savez_compressed(OUT / "field_forward.npz", **res)


def export_cases():
    res = {}
    for n, aabb in ((4, ((-1, -1, -1), (1, 1, 1))), (8, ((-1, -1, -1), (1, 1, 1))), (64, ((-1, -1, -1), (1, 1, 1))),
                    (10, ((-1.0, -0.5, -0.25), (1.0, 0.5, 0.75)))):
        pts, plane = ns.surface_points(aabb, n)
        tag = f"n{n}_{'cube' if aabb[0][1] == -1 else 'box'}"
        res[f"{tag}_points"] = pts.numpy()
        res[f"{tag}_plane"] = plane.numpy()
    np.savez_compressed(OUT / "export_grid.npz", **res)


if __name__ == "__main__":
    torch.set_num_threads(1)
    hash_cases()
    field_cases()
    export_cases()
    print("golden vectors written to", OUT)
