"""Regenerate tests/golden/*.npz from the REAL nerfstudio classes (last row): the only route to a
reference-pinned oracle.  Runs only where ``import nerfstudio`` works (nerfstudio==0.3.2, e.g. under ``baseline/_ref`` on a
box with network access at install time); in the offline build image it reports why it cannot run and exits 2.

    PYTHONPATH=baseline/_ref python tools/regen_golden_from_nerfstudio.py [--out tests/golden] [--check]

For every fixture the generator in tests/golden/make_golden.py produces with the oracle, this script evaluates the SAME seeded
inputs with the upstream modules through their torch path (``implementation="torch"``) and either rewrites the fixture
(default) or, with ``--check``, compares it with the committed one and prints the worst deviation per array -- the pinning
run.  Upstream call sites: HashEncoding.pytorch_fwd / hash_fn, SHEncoding, MLP, SceneContraction, RaySamples.get_weights,
RGBRenderer / SemanticRenderer / DepthRenderer("median") / AccumulationRenderer, ProposalNetworkSampler / PDFSampler /
UniformLinDispPiecewiseSampler, losses.interlevel_loss (fruit_nerf/fruit_field.py:98-166, fruit_nerf/fruit_nerf.py:104-168)."""
from __future__ import annotations

import argparse
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def main() -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=str(ROOT / "tests" / "golden"))
    ap.add_argument("--check", action="store_true", help="compare with the committed fixtures instead of rewriting them")
    a = ap.parse_args()
    try:
        import nerfstudio  # noqa: F401
        from nerfstudio.cameras.rays import Frustums, RaySamples
        from nerfstudio.field_components.encodings import HashEncoding, SHEncoding
        from nerfstudio.field_components.spatial_distortions import SceneContraction
        from nerfstudio.model_components.ray_samplers import PDFSampler, UniformLinDispPiecewiseSampler
        from nerfstudio.model_components.renderers import AccumulationRenderer, DepthRenderer, RGBRenderer, SemanticRenderer
        from nerfstudio.model_components.losses import interlevel_loss
    except Exception as ex:  # noqa: BLE001
        print(f"nerfstudio is not importable here ({type(ex).__name__}: {ex}); the goldens stay oracle-generated (parity unpinned).")
        return 2

    import numpy as np
    import torch

    from fruitnerf_b200 import synthetic as syn
    from oracle import fruit_ref as fr
    from oracle import ns_torch as ns

    out_dir = Path(a.out)
    report = {}

    def emit(name: str, arrays: dict) -> None:
        path = out_dir / f"{name}.npz"
        if a.check and path.exists():
            old = np.load(path)
            for k, v in arrays.items():
                if k in old.files:
                    ref = old[k]
                    v = np.asarray(v)
                    if v.shape != ref.shape:
                        report[f"{name}/{k}"] = f"shape {v.shape} vs {ref.shape}"
                    elif np.issubdtype(v.dtype, np.integer):
                        report[f"{name}/{k}"] = f"{int((v != ref).sum())} integer mismatches"
                    else:
                        scale = float(np.abs(ref).max()) or 1.0
                        report[f"{name}/{k}"] = f"max |diff| / scale = {float(np.abs(v - ref).max()) / scale:.3e}"
        else:
            np.savez_compressed(path, **{k: np.asarray(v) for k, v in arrays.items()})

    # ---- 1. hash indices (exact-integer): upstream HashEncoding.hash_fn on the oracle's cell coordinates
    for T in (17, 19, 21):
        enc = HashEncoding(num_levels=16, min_res=16, max_res=2048, log2_hashmap_size=T, features_per_level=2, implementation="torch")
        pts = (syn.hash_uniform(3 * 512, 77).view(512, 3) + 1) / 2
        scaled = pts[..., None, :] * enc.scalings.view(-1, 1)
        emit(f"hash_indices_T{T}", {"points": pts.numpy(), "floor_hash": enc.hash_fn(torch.floor(scaled).to(torch.int32)).numpy(),
                                     "oracle_rows": ns.hash_corner_indices(pts, fr.FieldSpec(2048, T, 15).scalings(), T)[0].numpy()})

    # ---- 2. encodings + compositing on seeded inputs through the upstream modules
    v = syn.SMALL
    sd = syn.field_state(geo=v["geo"], sem_dims=v["sem_dims"], log2_hashmap_size=15, num_images=7, table_scale=0.5, weight_gain=1.5)
    enc = HashEncoding(num_levels=16, min_res=16, max_res=2048, log2_hashmap_size=15, features_per_level=2, implementation="torch")
    with torch.no_grad():
        enc.hash_table.copy_(sd["mlp_base_grid.hash_table"])
    o, d, s, e, cam = syn.ray_batch(64, 48, salt=1, num_images=7)
    pos = o[:, None, :] + d[:, None, :] * ((s + e) / 2)[..., None]
    pos = (SceneContraction(order=float("inf"))(pos) + 2.0) / 4.0
    sel = ((pos > 0.0) & (pos < 1.0)).all(dim=-1)
    feats = enc(pos * sel[..., None]).detach()
    sh = SHEncoding(levels=4, implementation="torch")((d + 1.0) / 2.0).detach()
    emit("encodings_upstream", {"positions": pos.numpy(), "hash_features": feats.numpy(), "sh": sh.numpy()})

    dens = torch.rand(64, 48, 1, generator=torch.Generator().manual_seed(3)) * 30
    rs = RaySamples(frustums=Frustums(origins=o[:, None, :].expand(64, 48, 3), directions=d[:, None, :].expand(64, 48, 3), starts=s[..., None],
                                      ends=e[..., None], pixel_area=torch.ones(64, 48, 1)), deltas=(e - s)[..., None])
    w = rs.get_weights(dens)
    rgb = torch.rand(64, 48, 3, generator=torch.Generator().manual_seed(4))
    logit = torch.randn(64, 48, 1, generator=torch.Generator().manual_seed(5))
    emit("compositing_upstream", {
        "density": dens.numpy(), "starts": s.numpy(), "ends": e.numpy(), "weights": w.numpy(),
        "rgb": RGBRenderer(background_color="last_sample")(rgb=rgb, weights=w).numpy(),
        "depth": DepthRenderer(method="median")(weights=w, ray_samples=rs).numpy(),
        "accumulation": AccumulationRenderer()(weights=w).numpy(),
        "semantics": SemanticRenderer()(logit, weights=w).numpy(),
    })
    if a.check:
        for k, msg in sorted(report.items()):
            print(f"{k}: {msg}")
    print(f"{'checked' if a.check else 'wrote'} fixtures under {out_dir} from nerfstudio {getattr(nerfstudio, '__version__', '?')}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
