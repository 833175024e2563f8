"""CPU tests of the training shell's host logic: synthetic scene, pixel batches, schedules, optimiser
hyper-parameters (against torch.optim's own formulas), clustering."""
import math

import numpy as np
import pytest
import torch

from fruitnerf_b200 import _lib as L
from fruitnerf_b200.clustering import count_fruits, voxel_down_sample
from fruitnerf_b200.data.fruit_datamanager import FruitDataManagerConfig
from fruitnerf_b200.data.synthetic_scene import camera_rays, look_at_c2w, make_apple_scene, make_geometry, trace, write_dataset
from fruitnerf_b200.optim import ExponentialDecay, FusedAdam


def test_scene_geometry_and_masks():
    train, ev = make_apple_scene(num_images=20, height=48, width=48, num_fruits=6, seed=3)
    assert len(train) == 18 and len(ev) == 2  # every 10th image is held out (train_split_fraction 0.9)
    assert train.images.shape == (18, 48, 48, 3) and train.fruit_masks.shape == (18, 48, 48, 1)
    assert set(train.fruit_masks.unique().tolist()) <= {0.0, 1.0}
    assert 0.0 < float(train.fruit_masks.mean()) < 0.3
    assert float(train.images.min()) >= 0.0 and float(train.images.max()) <= 1.0
    # cameras inside the +/-1 scene box (auto_scale_poses), rotation part orthonormal, looking at the tree
    c2w = train.cameras.camera_to_worlds
    assert float(c2w[:, :, 3].abs().max()) <= 1.0
    eye = torch.eye(3).expand(len(train), 3, 3)
    assert torch.allclose(c2w[:, :, :3].transpose(1, 2) @ c2w[:, :, :3], eye, atol=1e-5)
    # fruit pixels are red-dominant, the mask marks exactly the fruit hits of the tracer
    fruit = train.fruit_masks[..., 0] > 0.5
    assert bool((train.images[fruit][:, 0] > train.images[fruit][:, 1]).all())
    g = train.geometry
    d = torch.cdist(g.fruit_centers, g.fruit_centers) + 10 * torch.eye(6)
    assert float(d.min()) > 0.14


def test_sensor_noise_is_seeded_and_bounded():
    a, _ = make_apple_scene(num_images=5, height=16, width=16, num_fruits=2, seed=1, noise_std=0.02)
    b, _ = make_apple_scene(num_images=5, height=16, width=16, num_fruits=2, seed=1, noise_std=0.02)
    clean, _ = make_apple_scene(num_images=5, height=16, width=16, num_fruits=2, seed=1, noise_std=0.0)
    assert torch.equal(a.images, b.images) and torch.equal(a.fruit_masks, clean.fruit_masks)
    assert float(a.images.min()) >= 0.0 and float(a.images.max()) <= 1.0
    assert 0.01 < float((a.images - clean.images).std()) < 0.025


def test_center_pixel_ray_hits_target():
    eye, target = torch.tensor([0.9, 0.1, 0.4]), torch.tensor([0.0, 0.0, 0.02])
    c2w = look_at_c2w(eye, target)
    # pixel centre (cx - 0.5) looks exactly along the optical axis
    o, d = camera_rays(c2w, 100.0, 100.0, 32.5, 24.5, torch.tensor([24]), torch.tensor([32]))
    want = (target - eye) / (target - eye).norm()
    assert torch.allclose(d[0], want, atol=1e-6) and torch.allclose(o[0], eye)
    # a ray straight at a fruit centre reports a fruit hit at distance |c - o| - r
    geom = make_geometry(4, seed=1)
    c = geom.fruit_centers[0]
    o = c + torch.tensor([0.0, 0.0, 1.0])
    rgb, mask, t = trace(geom, o[None], torch.tensor([[0.0, 0.0, -1.0]]))
    if float(mask[0, 0]) == 1.0:  # unless a leaf blob sits above it
        assert abs(float(t[0]) - (1.0 - geom.fruit_radius)) < 1e-5


def test_datamanager_pixel_batches_are_consistent():
    cfg = FruitDataManagerConfig(train_num_rays_per_batch=300, eval_num_rays_per_batch=64,
                                 synthetic_scene=dict(num_images=10, height=24, width=32, num_fruits=3, noise_std=0.0))
    dm = cfg.setup(device="cpu")
    rb, batch = dm.next_train(0)
    assert rb.origins.shape == (300, 3) and rb.camera_indices.shape == (300, 1) and batch["image"].shape == (300, 3)
    assert batch["fruit_mask"].shape == (300, 1)
    assert torch.allclose(rb.directions.norm(dim=-1), torch.ones(300), atol=1e-6)
    ds = dm.train_dataset
    idx = batch["indices"]
    assert int(idx[:, 0].max()) < len(ds) and int(idx[:, 1].max()) < 24 and int(idx[:, 2].max()) < 32
    assert torch.equal(batch["image"], ds.images[idx[:, 0], idx[:, 1], idx[:, 2]])
    assert torch.equal(rb.origins, ds.cameras.camera_to_worlds[idx[:, 0], :, 3])
    # the pixel colour is what the tracer returns for that ray
    rgb, mask, _ = trace(ds.geometry, rb.origins, rb.directions)
    assert torch.allclose(rgb, batch["image"], atol=1e-5) and torch.equal(mask, batch["fruit_mask"])
    rb2, _ = dm.next_train(1)
    assert not torch.equal(rb.directions, rb2.directions)
    i, cam_bundle, b = dm.next_eval_image(0)
    assert cam_bundle.origins.shape == (24, 32, 3) and b["image"].shape == (24, 32, 3)


def test_write_dataset_layout(tmp_path):
    train, _ = make_apple_scene(num_images=5, height=16, width=16, num_fruits=2)
    path = write_dataset(train, tmp_path / "apple")
    import json

    meta = json.loads(path.read_text())
    assert {"fl_x", "fl_y", "cx", "cy", "w", "h", "frames"} <= set(meta)
    fr0 = meta["frames"][0]
    assert {"file_path", "semantic_path", "transform_matrix"} <= set(fr0)
    from PIL import Image

    m = np.array(Image.open(tmp_path / "apple" / fr0["semantic_path"]))
    assert set(np.unique(m).tolist()) <= {0, 255}  # fruit_dataset.py:47-52 normalises by 255


def test_exponential_decay_schedule():
    s = ExponentialDecay(1e-2, 1e-4, 200000)
    assert s.lr(0) == pytest.approx(1e-2) and s.lr(200000) == pytest.approx(1e-4) and s.lr(10**7) == pytest.approx(1e-4)
    assert s.lr(100000) == pytest.approx(1e-3)
    assert ExponentialDecay(1e-2, None, None).lr(12345) == 1e-2


@pytest.mark.parametrize("kind", ["Adam", "RAdam"])
def test_hyper_values_follow_torch_formulas(kind):
    opt = FusedAdam.__new__(FusedAdam)
    opt.kind = {"Adam": L.FNR_OPT_ADAM, "RAdam": L.FNR_OPT_RADAM}[kind]
    opt.lr, opt.eps, opt.betas, opt.scheduler = 1e-2, 1e-15, (0.9, 0.999), ExponentialDecay(1e-2, 1e-4, 1000)
    for step in (1, 2, 5, 6, 7, 100):
        h = opt._hyper_values(step, 0.5, sched_step=step - 1)
        assert h[0] == pytest.approx(opt.scheduler.lr(step - 1), rel=1e-6)  # LambdaLR: the k-th iteration uses lambda(k-1)
        # a group whose optimiser was skipped on some iterations: the lr is read at the TRAINER step, the bias corrections
        # at the update count (nerfstudio steps every scheduler every iteration, torch skips parameters without a gradient)
        h2 = opt._hyper_values(step, 0.5, sched_step=6 * step)
        assert h2[0] == pytest.approx(opt.scheduler.lr(6 * step), rel=1e-6) and h2[4] == h[4] and h2[5] == h[5]
        assert h[4] == pytest.approx(1 - 0.9**step, rel=1e-6) and h[5] == pytest.approx(1 - 0.999**step, rel=1e-5)
        assert h[7] == 0.5
        if kind == "RAdam":
            rho_inf = 2 / (1 - 0.999) - 1
            rho_t = rho_inf - 2 * step * 0.999**step / (1 - 0.999**step)
            if rho_t > 5:
                want = math.sqrt((rho_t - 4) * (rho_t - 2) * rho_inf / ((rho_inf - 4) * (rho_inf - 2) * rho_t))
                assert h[6] == pytest.approx(want, rel=1e-5)
            else:
                assert h[6] < 0
        else:
            assert h[6] < 0


def test_fused_adam_refuses_cpu_parameters():
    with pytest.raises(L.FruitNerfNativeError):
        FusedAdam([torch.zeros(4)])


def test_clustering_counts_blobs_and_merges_fragments():
    rng = np.random.default_rng(0)
    centers = np.array([[0, 0, 0], [0.5, 0, 0], [0, 0.5, 0.2], [0.4, 0.4, 0.4]], dtype=float)
    pts = np.concatenate([c + 0.02 * rng.standard_normal((400, 3)) for c in centers])
    # a small fragment 0.05 away from blob 0 (what DBSCAN splits off a fruit): merged by the centre-distance rule
    frag = centers[0] + np.array([0.05, 0, 0]) + 0.004 * rng.standard_normal((60, 3))
    noise = rng.uniform(-1, 1, (30, 3))
    res = count_fruits(np.concatenate([pts, frag, noise]), eps=0.012, min_samples=8, cluster_merge_distance=0.08)
    assert res["count"] == 4 and res["count_before_merge"] >= 4
    d = np.linalg.norm(res["centers"][:, None] - centers[None], axis=-1).min(axis=0)
    assert d.max() < 0.03
    assert count_fruits(np.zeros((0, 3)), 0.1, 5, 0.1)["count"] == 0
    ds = voxel_down_sample(np.array([[0.0, 0, 0], [0.01, 0, 0], [1.0, 1, 1]]), 0.1)
    assert ds.shape == (2, 3)


def test_synthetic_spec_keeps_reference_hyperparameters_and_optional_short_schedule():
    from fruitnerf_b200.fruit_nerf_config import METHODS
    from fruitnerf_b200.scripts.train import synthetic_spec

    stock = synthetic_spec("fruit_nerf")
    assert stock.optimizers["fields"]["scheduler"] == METHODS["fruit_nerf"].optimizers["fields"]["scheduler"]  # 1e-4 after 200 k steps
    assert stock.pipeline.datamanager.synthetic_scene["num_fruits"] == 12
    short = synthetic_spec("fruit_nerf_big", schedule_steps=1234)
    for group in ("fields", "proposal_networks"):
        assert short.optimizers[group]["scheduler"] == {"type": "ExponentialDecay", "lr_final": 1e-4, "max_steps": 1234}
        assert short.optimizers[group]["optimizer"]["type"] == "RAdam"  # the optimiser itself is the method's
    assert METHODS["fruit_nerf_big"].optimizers["proposal_networks"]["scheduler"] is None  # the shared config is not mutated


def test_staged_device_buffer_on_cpu_and_export_slabs():
    from fruitnerf_b200.export.exporter_utils import export_slab
    from fruitnerf_b200.optim import StagedDeviceBuffer

    buf = StagedDeviceBuffer(3, torch.device("cpu"), slots=2)
    for k in range(5):  # more uploads than slots: the ring wraps
        buf.upload([k, 2 * k, 3 * k])
        assert buf.device_buffer.tolist() == [k, 2 * k, 3 * k]
    # slabs: disjoint cover, empty slabs when there are more ranks than rays
    for n, world in ((1000, 8), (5, 8), (0, 2), (262144, 3)):
        slabs = [export_slab(n, world, r) for r in range(world)]
        assert slabs[0][0] == 0 and slabs[-1][1] == n
        assert all(lo <= hi for lo, hi in slabs) and all(slabs[i][1] == slabs[i + 1][0] for i in range(world - 1))


def test_dataparser_reads_back_a_written_dataset(tmp_path):
    """write_dataset -> FruitNerf dataparser (fruit_nerf/data/fruitnerf_dataparser.py:73-292): split, pose pipeline,
    intrinsics, pixels and {0,1} fruit masks."""
    from fruitnerf_b200.data.fruitnerf_dataparser import FruitNerf, FruitNerfDataParserConfig, auto_orient_and_center_poses, load_fruit_datasets

    train, ev = make_apple_scene(num_images=20, height=24, width=32, num_fruits=4, seed=2, noise_std=0.0)
    import dataclasses

    # one data set holding all 20 frames in their original order
    full_imgs = torch.zeros(20, 24, 32, 3)
    full_masks = torch.zeros(20, 24, 32, 1)
    full_c2w = torch.zeros(20, 3, 4)
    idx = torch.arange(20)
    full_imgs[idx % 10 != 9], full_imgs[idx % 10 == 9] = train.images, ev.images
    full_masks[idx % 10 != 9], full_masks[idx % 10 == 9] = train.fruit_masks, ev.fruit_masks
    full_c2w[idx % 10 != 9], full_c2w[idx % 10 == 9] = train.cameras.camera_to_worlds, ev.cameras.camera_to_worlds
    full = dataclasses.replace(train, images=full_imgs, fruit_masks=full_masks,
                               cameras=dataclasses.replace(train.cameras, camera_to_worlds=full_c2w))
    write_dataset(full, tmp_path / "scene")
    cfg = FruitNerfDataParserConfig(data=tmp_path / "scene")
    tr, va, outs = load_fruit_datasets(cfg)
    # split: ceil(0.9 * 20) = 18 equally spaced training frames (first and last included), the rest held out
    i_train = np.linspace(0, 19, 18, dtype=int)
    i_eval = np.setdiff1d(np.arange(20), i_train)
    assert len(tr) == 18 and len(va) == 2
    assert [p.name for p in outs.image_filenames] == [f"frame_{i:05d}.png" for i in i_train]
    # pixels: 8-bit quantisation of the written PNGs; masks exactly {0,1}
    assert torch.allclose(tr.images, full_imgs[i_train], atol=0.5 / 255 + 1e-6)
    assert torch.equal(tr.fruit_masks, full_masks[i_train]) and torch.equal(va.fruit_masks, full_masks[i_eval])
    # poses: "up" orientation + centring on the mean camera + auto-scale into the +/-1 box, applied BEFORE the split
    m = torch.eye(4).repeat(20, 1, 1)
    m[:, :3, :4] = full_c2w
    oriented, transform = auto_orient_and_center_poses(m, "up", "poses")
    scale = 1.0 / float(oriented[:, :3, 3].abs().max())
    assert outs.dataparser_scale == pytest.approx(scale, rel=1e-6)
    want = oriented[i_train].clone()
    want[:, :3, 3] *= scale
    assert torch.allclose(tr.cameras.camera_to_worlds, want[:, :3, :4], atol=1e-6)
    assert float(tr.cameras.camera_to_worlds[:, :, 3].abs().max()) <= 1.0 + 1e-6
    assert torch.allclose(outs.dataparser_transform, transform)
    # rotations stay orthonormal, the mean "up" of the cameras points along +z after orientation
    R = oriented[:, :3, :3]
    assert torch.allclose(R.transpose(1, 2) @ R, torch.eye(3).expand(20, 3, 3), atol=1e-5)
    up = oriented[:, :3, 1].mean(0)
    assert float(up[2] / up.norm()) > 0.999
    # shared intrinsics, scene box, semantics metadata (classes of fruitnerf_dataparser.py:254)
    assert (tr.cameras.fx, tr.cameras.height, tr.cameras.width) == (train.cameras.fx, 24, 32)
    assert torch.equal(tr.scene_box.aabb, torch.tensor([[-1.0, -1, -1], [1, 1, 1]]))
    assert tr.metadata["semantics"].classes == ["apple", "stuff"] and len(tr.metadata["semantics"].filenames) == 18
    # the datamanager builds on it and the exporter's transform file round-trips
    dm = FruitDataManagerConfig(dataparser=cfg).setup(device="cpu")
    assert len(dm.train_dataset) == 18 and len(dm.eval_dataset) == 2
    outs.save_dataparser_transform(tmp_path / "run" / "dataparser_transforms.json")
    import json

    j = json.loads((tmp_path / "run" / "dataparser_transforms.json").read_text())
    assert j["scale"] == pytest.approx(scale) and np.allclose(j["transform"], transform.numpy())


def test_dataparser_rejects_what_the_device_ray_generator_cannot_model(tmp_path):
    from fruitnerf_b200.data.fruitnerf_dataparser import FruitNerf, FruitNerfDataParserConfig
    import json

    train, _ = make_apple_scene(num_images=10, height=8, width=8, num_fruits=2, seed=0, noise_std=0.0)
    tj = write_dataset(train, tmp_path / "s")
    meta = json.loads(tj.read_text())
    meta["k1"] = 0.1
    tj.write_text(json.dumps(meta))
    with pytest.raises(NotImplementedError):
        FruitNerf(FruitNerfDataParserConfig(data=tmp_path / "s")).get_dataparser_outputs("train")
    meta["k1"] = 0.0
    meta["train_filenames"] = ["images/frame_00000.png", "images/frame_00003.png"]
    tj.write_text(json.dumps(meta))
    out = FruitNerf(FruitNerfDataParserConfig(data=tj)).get_dataparser_outputs("train")  # explicit json path + split file list
    assert len(out.image_filenames) == 2 and out.images.shape[0] == 2
    with pytest.raises(RuntimeError):
        FruitNerf(FruitNerfDataParserConfig(data=tj)).get_dataparser_outputs("val")


def test_eval_setup_rebuilds_a_run_folder_on_cpu(tmp_path):
    """config.yml + newest step-*.ckpt -> pipeline (scripts/exporter.py eval_setup; nerfstudio eval_utils.eval_setup)."""
    import yaml

    from fruitnerf_b200.scripts.exporter import eval_setup
    from fruitnerf_b200.scripts.train import synthetic_spec

    spec = synthetic_spec("fruit_nerf", num_images=10, image_size=16, num_fruits=2)
    spec.pipeline.model.log2_hashmap_size = 10
    pipe = spec.pipeline.setup(device="cpu", test_mode="val")
    run = tmp_path / "outputs" / "exp" / "fruit_nerf" / "run0"
    (run / "nerfstudio_models").mkdir(parents=True)
    (run / "config.yml").write_text(yaml.dump(spec))
    for step in (3, 12):
        with torch.no_grad():
            pipe.model.field.mlp_head.layers[0].bias.fill_(float(step))
        torch.save({"step": step, "pipeline": pipe.state_dict()}, run / "nerfstudio_models" / f"step-{step:09d}.ckpt")
    config, loaded, path, step = eval_setup(run / "config.yml", test_mode="export", device="cpu")
    assert step == 12 and path.name == "step-000000012.ckpt" and not loaded.training
    assert config.load_dir.parts[-3] == "fruit_nerf" and loaded.model.test_mode == "export"
    a, b = pipe.state_dict(), loaded.state_dict()
    assert set(a) == set(b) and all(torch.equal(a[k], b[k]) for k in a)
    # reference-written checkpoints: upstream-only modules are dropped by the tolerant loader, strict loading refuses them
    extra = dict(a)
    extra["_model.lpips.net.weight"] = torch.zeros(3)
    extra["datamanager.train_camera_optimizer.pose_adjustment"] = torch.zeros(5, 6)
    with pytest.raises(RuntimeError):
        loaded.load_pipeline(extra, 12)
    loaded.load_pipeline({("module." + k): v for k, v in extra.items()}, 12, strict=False)
    with pytest.raises(ValueError):
        loaded.load_pipeline({**extra, "_model.field.mlp_base.params": torch.zeros(4)}, 12, strict=False)


def test_oracle_trainer_tool_runs_and_learns(tmp_path):
    """tools/oracle_train.py (the CPU oracle trained with torch.optim.Adam: the reference-side run of DESIGN.md section 7): a few
    iterations on a tiny scene lower the loss, the evaluation and the export + counting stages run."""
    import importlib.util
    import pathlib

    path = pathlib.Path(__file__).resolve().parent.parent / "tools" / "oracle_train.py"
    spec = importlib.util.spec_from_file_location("oracle_train_tool", path)
    tool = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(tool)
    lines = []
    res = tool.train(steps=6, seed=0, short_schedule=True, log_every=1, image_size=16, num_images=10, num_fruits=3, rays_per_batch=256,
                     points_per_side=12, state_path=str(tmp_path / "state.pt"), do_export=True, log=lines.append)
    h = res["history"]
    assert [r["step"] for r in h] == [1, 2, 3, 4, 5, 6]
    assert h[-1]["loss"] < h[0]["loss"] and all(math.isfinite(r["loss"]) for r in h)
    assert {"rgb_loss", "semantics_loss", "interlevel_loss", "fields_grad_norm"} <= set(h[0])
    assert set(res["eval"]) == {"psnr", "fruit_iou"} and math.isfinite(res["eval"]["psnr"])
    assert res["export"]["export_points"] == 12 ** 3 and set(res["export"]["cloud_sizes"]) == {"semantic_colormap", "semantic", "density"}
    assert (tmp_path / "state.pt").exists() and not res["events"]
