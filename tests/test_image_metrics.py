"""Full-image evaluation metrics (FruitModel.get_image_metrics_and_images, fruit_nerf/fruit_nerf.py:403-458) and their
routing through FruitPipeline (fruit_pipeline.py:155-227) -- host code, CPU."""
import math

import numpy as np
import pytest
import torch

from fruitnerf_b200 import image_metrics as im
from fruitnerf_b200.scripts.train import synthetic_spec


def _ssim_naive(a: np.ndarray, b: np.ndarray, k=11, sigma=1.5, k1=0.01, k2=0.03, data_range=None) -> float:
    """Direct per-pixel SSIM of two [H,W] float64 images: gaussian window on the reflect-padded image, the padded border
    of the SSIM map dropped (the recipe of torchmetrics' structural_similarity_index_measure), written with plain loops."""
    H, W = a.shape
    pad = (k - 1) // 2
    L = data_range if data_range is not None else max(a.max() - a.min(), b.max() - b.min())
    c1, c2 = (k1 * L) ** 2, (k2 * L) ** 2
    g = np.exp(-(((np.arange(k) - (k - 1) / 2) / sigma) ** 2) / 2)
    g /= g.sum()
    win = np.outer(g, g)
    ap, bp = np.pad(a, pad, mode="reflect"), np.pad(b, pad, mode="reflect")
    vals = []
    # the convolution output has the size of the unpadded image; its outer `pad` rows / columns are cropped
    for y in range(pad, H - pad):
        for x in range(pad, W - pad):
            pa, pb = ap[y:y + k, x:x + k], bp[y:y + k, x:x + k]
            mu_a, mu_b = (win * pa).sum(), (win * pb).sum()
            va, vb, cab = (win * pa * pa).sum() - mu_a ** 2, (win * pb * pb).sum() - mu_b ** 2, (win * pa * pb).sum() - mu_a * mu_b
            vals.append(((2 * mu_a * mu_b + c1) * (2 * cab + c2)) / ((mu_a ** 2 + mu_b ** 2 + c1) * (va + vb + c2)))
    return float(np.mean(vals))


def test_ssim_matches_a_direct_evaluation_and_basic_properties():
    g = torch.Generator().manual_seed(0)
    a = torch.rand(1, 1, 26, 30, generator=g)
    b = (a + 0.15 * torch.randn(1, 1, 26, 30, generator=g)).clamp(0, 1)
    got = float(im.ssim(a, b))
    want = _ssim_naive(a[0, 0].double().numpy(), b[0, 0].double().numpy())
    assert got == pytest.approx(want, abs=2e-5)
    assert float(im.ssim(a, a)) == pytest.approx(1.0, abs=1e-6)
    assert float(im.ssim(a, b)) == pytest.approx(float(im.ssim(b, a)), abs=1e-6)
    worse = (a + 0.4 * torch.randn(1, 1, 26, 30, generator=g)).clamp(0, 1)
    assert float(im.ssim(a, worse)) < got < 1.0
    # channels are filtered independently and averaged
    a3 = torch.rand(2, 3, 24, 24, generator=g)
    b3 = (a3 + 0.1 * torch.randn(2, 3, 24, 24, generator=g)).clamp(0, 1)
    # (ssim() takes data_range over the whole batch: pin it for the per-image comparison)
    per = [_ssim_naive(a3[i, c].double().numpy(), b3[i, c].double().numpy(), data_range=1.0) for i in range(2) for c in range(3)]
    assert float(im.ssim(a3, b3, data_range=1.0)) == pytest.approx(np.mean(per), abs=2e-5)
    with pytest.raises(ValueError):
        im.ssim(a, a[0])


def test_psnr_and_jaccard_known_answers():
    a = torch.zeros(1, 3, 4, 4)
    b = torch.full((1, 3, 4, 4), 0.1)
    assert float(im.psnr(a, b)) == pytest.approx(20.0, abs=1e-4)  # mse 0.01
    p = torch.tensor([0.9, 0.6, 0.4, 0.1, 0.7])
    t = torch.tensor([1.0, 0.0, 1.0, 0.0, 1.0])
    assert float(im.binary_jaccard(p, t)) == pytest.approx(2 / 4)  # pred {0,1,4}, target {0,2,4}: 2 common, 4 in the union
    assert float(im.binary_jaccard(torch.zeros(5), torch.zeros(5))) == 0.0


def test_colormaps_shapes_ranges_and_depth_fade():
    x = torch.linspace(0, 1, 64).view(8, 8, 1)
    c = im.apply_colormap(x)
    assert c.shape == (8, 8, 3) and float(c.min()) >= 0 and float(c.max()) <= 1
    lo, hi = im.apply_colormap(torch.full((1, 1, 1), 0.1))[0, 0], im.apply_colormap(torch.full((1, 1, 1), 0.9))[0, 0]
    assert float(lo[2]) > 0.7 and float(lo[2]) > 2 * float(lo[0])  # turbo: blue at the low end ...
    assert float(hi[0]) > 0.7 and float(hi[0]) > 4 * float(hi[2])  # ... red at the high end
    mid = im.apply_colormap(torch.full((1, 1, 1), 0.5))[0, 0]
    assert float(mid[1]) > 0.9 and float(mid[1]) > float(mid[0]) > float(mid[2])  # green-yellow in the middle
    rgb = torch.rand(4, 4, 3)
    assert im.apply_colormap(rgb) is rgb
    depth = torch.rand(8, 8, 1) * 3 + 1
    acc = torch.zeros(8, 8, 1)
    assert torch.equal(im.apply_depth_colormap(depth, acc), torch.ones(8, 8, 3))  # nothing accumulated: white
    full = im.apply_depth_colormap(depth, torch.ones(8, 8, 1))
    near = (depth == depth.min()).nonzero()[0]
    assert torch.allclose(full[near[0], near[1]], im.apply_colormap(torch.zeros(1, 1, 1))[0, 0])


def _fake_outputs(H, W, mask, g):
    image = torch.rand(H, W, 3, generator=g)
    rgb = image + 0.05 * torch.randn(H, W, 3, generator=g)  # not clamped: the metric clamps (fruit_nerf.py:408)
    sem = torch.where(mask > 0.5, torch.tensor(8.0), torch.tensor(-8.0))
    out = {"rgb": rgb, "accumulation": torch.rand(H, W, 1, generator=g), "depth": torch.rand(H, W, 1, generator=g) + 0.5, "semantics": sem,
           "prop_depth_0": torch.rand(H, W, 1, generator=g), "prop_depth_1": torch.rand(H, W, 1, generator=g),
           "semantics_colormap": torch.zeros(H, W, 3, dtype=torch.long)}
    return image, out


def test_image_metrics_and_images_follow_the_reference_keys():
    g = torch.Generator().manual_seed(1)
    H, W = 24, 20
    mask = (torch.rand(H, W, 1, generator=g) > 0.7).float()
    image, out = _fake_outputs(H, W, mask, g)
    metrics, images = im.image_metrics_and_images(out, {"image": image, "fruit_mask": mask}, 2, torch.device("cpu"))
    assert set(metrics) == {"psnr", "ssim", "iou", "fruit_iou"}
    assert all(isinstance(v, float) for v in metrics.values())
    assert set(images) == {"img", "accumulation", "depth", "prop_depth_0", "prop_depth_1", "semantics_colormap", "fruit_mask"}
    assert images["img"].shape == (H, 2 * W, 3)  # ground truth | render, side by side (fruit_nerf.py:415)
    assert images["accumulation"].shape == images["depth"].shape == images["prop_depth_1"].shape == (H, W, 3)
    assert images["fruit_mask"].shape == (H, W, 3) and images["semantics_colormap"].shape == (H, W, 1)
    clamped = out["rgb"].clamp(0, 1)
    assert metrics["psnr"] == pytest.approx(-10 * math.log10(float(((clamped - image) ** 2).mean())), rel=1e-5)
    assert 0.0 < metrics["ssim"] < 1.0
    assert metrics["fruit_iou"] == 1.0  # logits +-8 follow the mask exactly
    assert metrics["iou"] == 0.0  # the reference's softmax over image rows never exceeds the 0.5 threshold (see image_metrics.py)
    # empty mask, nothing predicted: agreement
    m2, _ = im.image_metrics_and_images({**out, "semantics": torch.full((H, W, 1), -8.0)}, {"image": image, "fruit_mask": torch.zeros(H, W, 1)}, 2,
                                        torch.device("cpu"))
    assert m2["fruit_iou"] == 1.0


def test_pipeline_eval_metrics_route_through_the_model(tmp_path):
    """get_eval_image_metrics_and_images / get_average_eval_image_metrics with the renderer replaced by a stub (the render itself
    needs the GPU): image_idx / num_rays bookkeeping, averaging of every metric, images written under output_path."""
    spec = synthetic_spec("fruit_nerf", num_images=10, image_size=24, num_fruits=3, seed=0)
    pipeline = spec.pipeline.setup(device="cpu", test_mode="val")
    pipeline.train()
    dm, model = pipeline.datamanager, pipeline.model
    assert hasattr(model, "get_image_metrics_and_images")
    g = torch.Generator().manual_seed(2)
    calls = []

    def fake_render(bundle):
        i = int(bundle.camera_indices[0, 0, 0])
        calls.append(i)
        assert not model.training  # evaluation renders run in eval mode (fruit_pipeline.py:163)
        ds = dm.eval_dataset
        _, out = _fake_outputs(24, 24, ds.fruit_masks[i], g)
        out["rgb"] = ds.images[i] + 0.02 * torch.randn(24, 24, 3, generator=g)
        return out

    model.get_outputs_for_camera_ray_bundle = fake_render
    n = len(dm.eval_dataset)
    metrics, images = pipeline.get_eval_image_metrics_and_images(0)
    assert metrics["image_idx"] == 0 and metrics["num_rays"] == 24 * 24 and metrics["psnr"] > 25
    assert pipeline.training  # restored
    avg = pipeline.get_average_eval_image_metrics(0, output_path=tmp_path / "renders")
    assert {"psnr", "ssim", "iou", "fruit_iou", "num_rays_per_sec", "fps"} <= set(avg) and "image_idx" not in avg
    assert avg["psnr"] > 25 and avg["fruit_iou"] == 1.0
    assert len(calls) == 1 + n
    files = sorted(p.name for p in (tmp_path / "renders").iterdir())
    assert len(files) == n * len(images) and files[0].endswith(".jpg")
