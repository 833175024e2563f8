"""Shared helpers of the test-suite (oracle-side glue + tolerances)."""
from __future__ import annotations

import torch

from fruitnerf_b200 import synthetic as syn
from fruitnerf_b200.fruit_field import FruitField, SceneContraction
from oracle import fruit_ref as fr

REL = 1e-3  # north_star tolerance: 1e-3 relative on RGB / density / semantics
FLOOR = 0.01  # elements below 1% of the tensor's max magnitude are compared absolutely (against rel * 1% * max)


def assert_rel(actual, expected, rel=REL, floor=FLOOR, what=""):
    """|a-b| <= rel * max(|b|, floor * max|b|) element-wise."""
    a = actual.detach().double().cpu().reshape(-1)
    b = expected.detach().double().cpu().reshape(-1)
    assert a.shape == b.shape, (what, actual.shape, expected.shape)
    assert torch.isfinite(a).all(), f"{what}: non-finite values"
    scale = float(b.abs().max()) if b.numel() else 0.0
    tol = rel * torch.maximum(b.abs(), torch.full_like(b, floor * scale)) + 1e-30
    err = (a - b).abs()
    bad = err > tol
    if bad.any():
        i = int(torch.argmax(err / tol))
        raise AssertionError(
            f"{what}: {int(bad.sum())}/{a.numel()} elements off; worst idx {i}: got {a[i]:.8g} want {b[i]:.8g} "
            f"(err {err[i]:.3g}, tol {tol[i]:.3g}, scale {scale:.3g})"
        )


def variant(name: str):
    return dict(syn.SMALL if name == "small" else syn.BIG)


def make_state(name: str, table_scale=0.5, weight_gain=1.5, num_images=7, log2T=None):
    v = variant(name)
    T = log2T or v["log2_hashmap_size"]
    sd = syn.field_state(geo=v["geo"], sem_dims=v["sem_dims"], log2_hashmap_size=T, num_images=num_images,
                         table_scale=table_scale, weight_gain=weight_gain)
    spec = fr.FieldSpec(max_res=v["max_res"], log2_hashmap_size=T, geo_feat_dim=v["geo"])
    return sd, spec


def make_field(name: str, sd, spec, device, contraction=True, test_mode=None, **kw) -> FruitField:
    v = variant(name)
    f = FruitField(
        aabb=sd["aabb"], num_images=sd["embedding_appearance.embedding.weight"].shape[0], geo_feat_dim=v["geo"],
        max_res=v["max_res"], log2_hashmap_size=spec.log2_hashmap_size, num_layers_semantic=len(v["sem_dims"]) - 1,
        hidden_dim_semantics=v["sem_dims"][1], use_semantics=True, num_semantic_classes=1, test_mode=test_mode,
        spatial_distortion=SceneContraction(order=float("inf")) if contraction else None, **kw,
    )
    missing, unexpected = f.load_state_dict({k: v_ for k, v_ in sd.items()}, strict=False)
    # only the Sequential aliases and registered scalar buffers may be absent from the synthetic dict
    assert not unexpected, unexpected
    assert all(k.startswith("mlp_base.") or k in ("max_res", "num_levels", "log2_hashmap_size") for k in missing), missing
    return f.to(device)
