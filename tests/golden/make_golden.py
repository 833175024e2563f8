"""Generates tests/golden/*.npz from the CPU oracle (run in the build container:
``python -m tests.golden.make_golden``).

PARITY UNPINNED: the reference ships no tests / golden vectors and its arithmetic lives in
nerfstudio 0.3.2 + tinycudann, neither of which is installable here, so these vectors freeze the
ORACLE's output (nerfstudio torch-fallback semantics as restated in oracle/ns_torch.py).  They pin
the oracle against accidental change and travel to the GPU box, where /root/reference and the
oracle's provenance are not available; the `-m gpu` tests compare the CUDA path with them too.
"""
from pathlib import Path

import math

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


def proposal_cases():
    """Proposal stage: piecewise lin-disp bins, proposal weights, PDF resampling (eval / single jitter / per-bin, with
    annealing and an all-zero histogram), interlevel loss."""
    res = {}
    R = 12
    aabb = torch.tensor([[-1.0, -1.0, -1.0], [1.0, 1.0, 1.0]])
    o, d, _, _, _ = syn.ray_batch(R, 4, salt=901, far=3.0, num_images=3)
    nears, fars = torch.full((R, 1), 0.05), torch.full((R, 1), 1000.0)
    sd = syn.density_state(num_levels=5, log2_hashmap_size=12, salt=6100, table_scale=1.0, weight_gain=2.0)
    spec = fr.DensitySpec(num_levels=5, max_res=128, log2_hashmap_size=12)
    bins = ns.spaced_bins(R, 64, syn.hash_uniform(R, 911).abs().view(R, 1)).expand(R, 65).contiguous()
    e = ns.spacing_to_euclidean(bins, nears, fars)
    w = fr.proposal_weights(sd, spec, o, d, e[:, :-1], e[:, 1:], aabb)
    res["bins0"], res["euclid0"], res["weights0"] = bins.numpy(), e.numpy(), w.numpy()
    w_zero = w.clone()
    w_zero[0] = 0.0  # histogram padding + the eps guard take over
    for tag, u, anneal in (("eval", None, 1.0), ("single", syn.hash_uniform(R, 912).abs().view(R, 1), 0.37),
                           ("perbin", syn.hash_uniform(R * 25, 913).abs().view(R, 25), 1.0)):
        nb = ns.pdf_sample(torch.pow(w_zero, anneal), bins, 24, u)
        res[f"pdf_{tag}_bins"] = nb.numpy()
        if u is not None:
            res[f"pdf_{tag}_u"] = u.numpy()
    nb = torch.from_numpy(res["pdf_eval_bins"])
    e2 = ns.spacing_to_euclidean(nb, nears, fars)
    w2 = fr.proposal_weights(sd, spec, o, d, e2[:, :-1], e2[:, 1:], aabb)
    res["weights1"] = w2.numpy()
    res["interlevel"] = np.array(float(ns.interlevel_loss([w, w2], [bins, nb])))
    res["outer"] = ns.lossfun_outer(nb, w2, bins, w).numpy()
    res["origins"], res["directions"] = o.numpy(), d.numpy()
    for k, v in sd.items():
        res[f"sd_{k}"] = v.numpy()
    np.savez_compressed(OUT / "proposal.npz", **res)


def composite_and_gradient_cases():
    """Hand-made compositing inputs (sigma = 0, opaque first sample, NaN density, exact 0.5 median tie) and the oracle's
    autograd gradients of the small field on a tiny batch (incl. the semantic detach)."""
    res = {}
    dens = torch.tensor([[0.0, 0.0, 0.0, 0.0], [1e9, 1.0, 1.0, 1.0], [0.5, float("nan"), 0.5, 0.5], [math.log(2.0), 1e9, 0.0, 0.0],
                         [0.3, 0.7, 1.1, 2.0]])[..., None]
    starts = torch.arange(4, dtype=torch.float32).expand(5, 4)[..., None]
    ends = starts + 1.0
    rgb = syn.hash_uniform(5 * 4 * 3, 921).abs().view(5, 4, 3)
    sem = syn.hash_uniform(5 * 4, 922).view(5, 4, 1) * 4
    r = fr.render({"density": dens, "rgb": rgb, "semantics": sem}, starts, ends, training=False)
    res["c_density"], res["c_rgb"], res["c_semantics"] = dens.numpy(), rgb.numpy(), sem.numpy()
    for k in ("rgb", "accumulation", "depth", "depth_index", "semantics", "weights"):
        res[f"c_out_{k}"] = r[k].numpy()
    v = syn.SMALL
    sd = syn.field_state(geo=v["geo"], sem_dims=v["sem_dims"], log2_hashmap_size=12, num_images=3, table_scale=0.5, weight_gain=1.5)
    spec = fr.FieldSpec(max_res=v["max_res"], log2_hashmap_size=12, geo_feat_dim=v["geo"])
    o, d, s, e, cam = syn.ray_batch(16, 12, salt=931, far=3.0, num_images=3)
    img, mask = syn.targets(16, salt=932)
    sdg = {k: t.clone().requires_grad_(k != "aabb") for k, t in sd.items()}
    f = fr.field_forward(sdg, spec, o[:, None, :], d[:, None, :], s[..., None], e[..., None], cam, True, "train")
    out = fr.render(f, s[..., None], e[..., None], training=True)
    loss = sum(fr.loss_dict(out, img, mask).values())
    loss.backward()
    res["g_loss"] = np.array(float(loss))
    for k in ("mlp_base_mlp.layers.0.weight", "mlp_base_mlp.layers.1.bias", "mlp_semantics.layers.0.weight", "field_head_semantics.net.weight",
              "mlp_head.layers.2.weight", "embedding_appearance.embedding.weight"):
        res[f"g_{k}"] = sdg[k].grad.numpy()
    tg = sdg["mlp_base_grid.hash_table"].grad
    nz = torch.nonzero(tg.abs().sum(-1)).reshape(-1)[:512]
    res["g_table_rows"], res["g_table_vals"] = nz.numpy(), tg[nz].numpy()
    np.savez_compressed(OUT / "composite_gradients.npz", **res)


if __name__ == "__main__":
    import sys

    torch.set_num_threads(1)
    if "--new-only" not in sys.argv:  # the first three files are frozen; regenerate them only on purpose
        hash_cases()
        field_cases()
        export_cases()
    proposal_cases()
    composite_and_gradient_cases()
    print("golden vectors written to", OUT)
