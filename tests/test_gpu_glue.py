"""GPU parity of the per-ray glue kernels (fnr_glue.cu) against the torch formulas the reference evaluates
(nerfstudio PixelSampler / RayGenerator / SpacedSampler / MSELoss / BCEWithLogitsLoss / distortion_loss / median depth)."""
import pytest
import torch

from fruitnerf_b200 import _lib as L
from fruitnerf_b200 import ops
from fruitnerf_b200 import synthetic as syn
from fruitnerf_b200.components.ray_samplers import UniformLinDispPiecewiseSampler, UniformSamplerWithNoise
from fruitnerf_b200.compat import RayBundle
from fruitnerf_b200.data.synthetic_scene import camera_rays, make_apple_scene
from oracle import ns_torch as ns

pytestmark = pytest.mark.gpu


def test_pixel_batch_matches_torch_indexing(native_lib, cuda_device):
    train, _ = make_apple_scene(num_images=10, height=24, width=40, num_fruits=3, noise_std=0.0)
    cams = train.cameras
    R = 3000
    r = (syn.hash_uniform(R * 3, 41).view(R, 3) + 1) * 0.5
    r[0] = torch.tensor([0.0, 0.0, 0.0])
    r[1] = torch.tensor([0.9999999, 0.9999999, 0.9999999])  # rounds to N / H / W after the multiply: must clamp, not overflow
    o, d, cam, idx, img, mask = ops.pixel_batch(r.cuda(), cams.camera_to_worlds.cuda(), train.images.cuda(), train.fruit_masks.cuda(), cams.fx,
                                                cams.fy, cams.cx, cams.cy)
    want = torch.floor(r * torch.tensor([len(train), 24.0, 40.0])).long()
    want = torch.minimum(want, torch.tensor([len(train) - 1, 23, 39]))
    assert torch.equal(idx.cpu(), want) and torch.equal(cam.cpu().long(), want[:, 0])
    ro, rd = camera_rays(cams.camera_to_worlds[want[:, 0]], cams.fx, cams.fy, cams.cx, cams.cy, want[:, 1], want[:, 2])
    assert torch.equal(o.cpu(), ro) and torch.allclose(d.cpu(), rd, atol=2e-7)
    assert torch.equal(img.cpu(), train.images[want[:, 0], want[:, 1], want[:, 2]])
    assert torch.equal(mask.cpu(), train.fruit_masks[want[:, 0], want[:, 1], want[:, 2]])


@pytest.mark.parametrize("cls,far", [(UniformLinDispPiecewiseSampler, 1000.0), (UniformLinDispPiecewiseSampler, 0.8), (UniformSamplerWithNoise, 2.0)])
@pytest.mark.parametrize("jitter", ["eval", "single", "perbin"])
def test_spaced_bins_match_torch_sampler_and_oracle(native_lib, cuda_device, cls, far, jitter):
    R, S = 77, 48
    nears, fars = torch.full((R, 1), 0.05), torch.full((R, 1), far)
    nears[3], fars[3] = 0.5, 3.0  # a ray that starts in the linear part and ends in the disparity part
    t = None if jitter == "eval" else (syn.hash_uniform(R if jitter == "single" else R * (S + 1), 7).view(R, -1) + 1) * 0.5
    mode = cls.native_mode
    base = torch.linspace(0.0, 1.0, S + 1)
    bins, starts, ends = ops.spaced_bins(base.cuda(), None if t is None else t.cuda(), nears.cuda(), fars.cuda(), S, mode)
    ref_bins = ns.spaced_bins(R, S, t).expand(R, S + 1)
    assert torch.equal(bins.cpu(), ref_bins), "spacing bins differ bitwise"
    sampler = cls(num_samples=S)
    s_near, s_far = sampler.spacing_fn(nears), sampler.spacing_fn(fars)
    e = sampler.spacing_fn_inv(ref_bins * s_far + (1 - ref_bins) * s_near)
    assert torch.equal(starts.cpu(), e[:, :-1]) and torch.equal(ends.cpu(), e[:, 1:]), "euclidean bins differ bitwise"
    if mode == 1:
        assert torch.equal(e, ns.spacing_to_euclidean(ref_bins, nears, fars))
    # the module itself takes the native route on CUDA tensors
    rb = RayBundle(origins=torch.zeros(R, 3, device=cuda_device), directions=torch.ones(R, 3, device=cuda_device), nears=nears.cuda(), fars=fars.cuda())
    rs = sampler.eval()(rb)
    assert rs.frustums.starts.shape == (R, S, 1)
    e_eval = sampler.spacing_fn_inv(ns.spaced_bins(R, S, None) * s_far + (1 - ns.spaced_bins(R, S, None)) * s_near)
    assert torch.equal(rs.frustums.ends[..., 0].cpu(), e_eval[:, 1:]) and torch.equal(rs.spacing_starts[..., 0].cpu(), ns.spaced_bins(R, S, None).expand(R, S + 1)[:, :-1])


def test_render_losses_match_torch(native_lib, cuda_device):
    R = 4096
    rgb = ((syn.hash_uniform(R * 3, 11).view(R, 3) + 1) * 0.5).cuda().requires_grad_()
    sem = (syn.hash_uniform(R, 12) * 12).cuda().requires_grad_()  # logits incl. large |x|
    img, mask = syn.targets(R, salt=13)
    mse, bce, psnr = ops.render_losses(rgb, sem, img.cuda(), mask.cuda(), 0.7)
    (1.3 * mse + 0.5 * bce).backward()
    rgb_r, sem_r = rgb.detach().cpu().double().requires_grad_(), sem.detach().cpu().double().requires_grad_()
    mse_r = torch.nn.functional.mse_loss(img.double(), rgb_r)
    bce_r = 0.7 * torch.nn.functional.binary_cross_entropy_with_logits(sem_r[:, None], mask.double())
    (1.3 * mse_r + 0.5 * bce_r).backward()
    assert float(mse) == pytest.approx(float(mse_r), rel=1e-5) and float(bce) == pytest.approx(float(bce_r), rel=1e-5)
    assert float(psnr) == pytest.approx(float(-10 * torch.log10(mse_r)), rel=1e-5)
    assert torch.allclose(rgb.grad.cpu().double(), rgb_r.grad, rtol=1e-5, atol=1e-10)
    assert torch.allclose(sem.grad.cpu().double(), sem_r.grad, rtol=1e-4, atol=1e-10)
    # only one of the two losses used downstream
    rgb2 = rgb.detach().clone().requires_grad_()
    ops.render_losses(rgb2, sem.detach(), img.cuda(), mask.cuda())[0].backward()
    assert torch.allclose(rgb2.grad.cpu().double(), rgb_r.grad / 1.3, rtol=1e-5, atol=1e-10)


@pytest.mark.parametrize("S", [48, 96, 37])
def test_distortion_and_median_depth_match_torch(native_lib, cuda_device, S):
    R = 200
    w = ((syn.hash_uniform(R * S, 21).view(R, S) + 1) * 0.5) ** 4
    w = w / w.sum(-1, keepdim=True) * ((syn.hash_uniform(R, 22).view(R, 1) + 1) * 0.6)  # accumulation in [0, 1.2): some rays never reach 0.5
    w[0] = 0.0
    w[1] = 0.0
    w[1, 5], w[1, 9] = 0.5, 0.5  # exact tie at 0.5 -> index 5 (searchsorted left)
    t = torch.sort((syn.hash_uniform(R * (S + 1), 23).view(R, S + 1) + 1) * 0.5, dim=-1).values
    ut = (t[..., 1:] + t[..., :-1]) / 2
    inter = torch.sum(w * torch.sum(w[..., None, :] * torch.abs(ut[..., :, None] - ut[..., None, :]), dim=-1), dim=-1)
    intra = torch.sum(w**2 * (t[..., 1:] - t[..., :-1]), dim=-1) / 3
    want = torch.mean(inter + intra)
    got = ops.distortion_metric(w.cuda(), t.cuda())
    assert float(got) == pytest.approx(float(want), rel=1e-4)
    starts, ends = t[:, :-1].contiguous(), t[:, 1:].contiguous()
    depth = ops.median_depth(w.cuda(), starts.cuda(), ends.cuda()).cpu()
    cum = torch.cumsum(w, dim=-1)
    idx = torch.clamp(torch.searchsorted(cum, torch.full((R, 1), 0.5), side="left"), 0, S - 1)
    ref = torch.gather((starts + ends) / 2, -1, idx)
    # the kernel's running sum is a warp scan, torch's cumsum a serial one: indices may differ where cum crosses 0.5 within 1e-6
    near_tie = (torch.gather(cum, -1, idx) - 0.5).abs()[:, 0] < 1e-6
    assert torch.equal(depth[~near_tie], ref[~near_tie])
    assert float(depth[1]) == float((starts[1, 5] + ends[1, 5]) / 2) and float(depth[0]) == float((starts[0, S - 1] + ends[0, S - 1]) / 2)
