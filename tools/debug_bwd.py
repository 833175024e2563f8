"""Localise differences between the tensor-core and simt backward on the GPU."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch

from fruitnerf_b200 import _lib as L
from fruitnerf_b200 import ops
from fruitnerf_b200 import synthetic as syn
from tests.util import make_field, make_state

dev = torch.device("cuda:0")


def grads(field, o, d, s, e, cam, impl, mode):
    for p in field.parameters():
        p.grad = None
    if mode == "render":
        out = ops.render(field.kernel_shape(), field.kernel_params(), o, d, s, e, cam, field.position_mode(), field.appearance_mode(), impl=impl)
        img, mask = syn.targets(o.shape[0])
        loss = torch.nn.functional.mse_loss(img.to(dev), out["rgb"]) + torch.nn.functional.binary_cross_entropy_with_logits(
            out["semantics"][:, None], mask.to(dev))
    else:
        sd_, srgb, ssem = ops.field(field.kernel_shape(), field.kernel_params(), o, d, s, e, cam, field.position_mode(), field.appearance_mode(), impl=impl)
        loss = srgb.sum() + 0.1 * sd_.sum() + ssem.pow(2).sum()
    loss.backward()
    return {k: p.grad.clone() for k, p in field.named_parameters() if p.grad is not None}


for (R, S, salt, mode) in ((32, 16, 9, "field"), (37, 50, 5, "render"), (64, 48, 5, "render"), (256, 192, 1, "render")):
    sd, spec = make_state("small", log2T=15)
    field = make_field("small", sd, spec, dev).train()
    o, d, s, e, cam = [t.to(dev) for t in syn.ray_batch(R, S, salt=salt, far=3.0 if mode == "render" else 2.0, num_images=7)]
    ga = grads(field, o, d, s, e, cam, L.FNR_IMPL_SIMT, mode)
    gb = grads(field, o, d, s, e, cam, L.FNR_IMPL_AUTO, mode)
    print(f"=== R={R} S={S} mode={mode}")
    for k in ga:
        a, b = ga[k], gb[k]
        sc = float(a.abs().max())
        err = (a - b).abs()
        print(f"  {k:45s} scale {sc:.3e} max|diff| {float(err.max()):.3e} rel-to-scale {float(err.max())/max(sc,1e-30):.2e}  n(>1e-3 scale) {int((err > 1e-3*sc).sum())}")
    k = "mlp_base_grid.hash_table"
    a, b = ga[k], gb[k]
    sc = float(a.abs().max())
    bad_rows = torch.nonzero(((a - b).abs() > 1e-3 * sc).any(dim=1)).reshape(-1)
    print("  bad table rows:", bad_rows.numel())
    if bad_rows.numel():
        rows, pos = ops.hash_indices(field.kernel_shape(), o, d, s, e, field.position_mode())
        rows = rows.reshape(R * S, -1).long()
        badset = torch.zeros(a.shape[0], dtype=torch.bool, device=dev)
        badset[bad_rows] = True
        frac = badset[rows].float().mean(dim=1)  # fraction of each point's 128 rows that are bad
        top = torch.topk(frac, k=min(8, frac.numel()))
        for f_, idx in zip(top.values.tolist(), top.indices.tolist()):
            print(f"    point {idx} (ray {idx // S}, sample {idx % S}; tile {idx // 128} row {idx % 128}): {f_*100:.0f}% of its rows bad; pos {pos.reshape(-1,3)[idx].tolist()}")
