"""CUDA path against the committed golden vectors (tests/golden/*.npz) -- runs on the GPU box without the oracle
having to recompute anything: proposal weights, PDF resampling, interlevel loss, training gradients."""
from pathlib import Path

import numpy as np
import pytest
import torch

from fruitnerf_b200 import _lib as L
from fruitnerf_b200 import ops
from fruitnerf_b200 import synthetic as syn
from fruitnerf_b200.density_field import HashMLPDensityField
from fruitnerf_b200.fruit_field import SceneContraction

from .util import assert_rel, make_field

pytestmark = pytest.mark.gpu
GOLD = Path(__file__).resolve().parent / "golden"


def _gold(name):
    return {k: torch.from_numpy(v) for k, v in np.load(GOLD / name).items()}


def test_proposal_stage_against_golden(native_lib, cuda_device):
    g = _gold("proposal.npz")
    R = g["origins"].shape[0]
    aabb = torch.tensor([[-1.0, -1.0, -1.0], [1.0, 1.0, 1.0]])
    net = HashMLPDensityField(aabb, hidden_dim=16, log2_hashmap_size=12, num_levels=5, max_res=128, spatial_distortion=SceneContraction())
    net.load_state_dict({k[3:]: v for k, v in g.items() if k.startswith("sd_")}, strict=False)
    net = net.to(cuda_device)
    e = g["euclid0"].cuda()
    o, d = g["origins"].cuda(), g["directions"].cuda()
    with torch.no_grad():
        w = net.weights(o, d, e[:, :-1].contiguous(), e[:, 1:].contiguous())
    assert_rel(w, g["weights0"], what="proposal weights")
    nears, fars = torch.full((R, 1), 0.05, device=cuda_device), torch.full((R, 1), 1000.0, device=cuda_device)
    w_zero = g["weights0"].clone()
    w_zero[0] = 0.0
    for tag, anneal in (("eval", 1.0), ("single", 0.37), ("perbin", 1.0)):
        u = g.get(f"pdf_{tag}_u")
        bins, starts, ends = ops.pdf_sample(w_zero.cuda(), g["bins0"].cuda(), 24, None if u is None else u.cuda(), anneal, nears, fars)
        assert torch.allclose(bins.cpu(), g[f"pdf_{tag}_bins"], rtol=0, atol=2e-6), tag
        # the same exponent through the device-scalar route used by CUDA-graph replays
        bins2, _, _ = ops.pdf_sample(w_zero.cuda(), g["bins0"].cuda(), 24, None if u is None else u.cuda(),
                                     torch.tensor([anneal], device=cuda_device), nears, fars)
        assert torch.equal(bins, bins2)
        assert bool((ends >= starts).all())
    loss = ops.interlevel_loss([g["weights0"].cuda(), g["weights1"].cuda()], [g["bins0"].cuda(), g["pdf_eval_bins"].cuda()], 1.0)
    assert float(loss) == pytest.approx(float(g["interlevel"]), rel=1e-4)


@pytest.mark.parametrize("impl", [L.FNR_IMPL_SIMT, L.FNR_IMPL_AUTO], ids=["simt", "auto"])
def test_training_gradients_against_golden(native_lib, cuda_device, impl):
    g = _gold("composite_gradients.npz")
    v = syn.SMALL
    sd = syn.field_state(geo=v["geo"], sem_dims=v["sem_dims"], log2_hashmap_size=12, num_images=3, table_scale=0.5, weight_gain=1.5)
    from oracle.fruit_ref import FieldSpec  # shape record only (no arithmetic)

    spec = FieldSpec(max_res=v["max_res"], log2_hashmap_size=12, geo_feat_dim=v["geo"])
    field = make_field("small", sd, spec, cuda_device).train()
    o, d, s, e, cam = syn.ray_batch(16, 12, salt=931, far=3.0, num_images=3)
    img, mask = syn.targets(16, salt=932)
    out = ops.render(field.kernel_shape(), field.kernel_params(), o.cuda(), d.cuda(), s.cuda(), e.cuda(), cam.cuda(), field.position_mode(),
                     L.FNR_APP_PER_CAMERA, impl=impl)
    loss = torch.nn.functional.mse_loss(img.cuda(), out["rgb"]) + torch.nn.functional.binary_cross_entropy_with_logits(
        out["semantics"][:, None], mask.cuda())
    loss.backward()
    assert float(loss) == pytest.approx(float(g["g_loss"]), rel=1e-4)
    named = dict(field.named_parameters())
    # 192 samples only: a flipped ReLU mask (tensor-core path, |pre-activation| < 1e-5) would show up as a 1-sample outlier;
    # none occurs for this seed on either implementation
    for k, ref in g.items():
        if k.startswith("g_") and k not in ("g_loss", "g_table_rows", "g_table_vals"):
            assert_rel(named[k[2:]].grad, ref, rel=2e-3, floor=0.25, what=k)
    assert_rel(named["mlp_base_grid.hash_table"].grad[g["g_table_rows"].cuda()], g["g_table_vals"], rel=2e-3, floor=0.25, what="table rows")
