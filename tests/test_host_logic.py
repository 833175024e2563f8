"""CPU tests of the host-side mirror of the reference's plugin surface."""
import os
import socket
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest
import torch

from fruitnerf_b200 import synthetic as syn
from fruitnerf_b200.compat import RayBundle, SceneBox, Semantics
from fruitnerf_b200.components.ray_samplers import UniformLinDispPiecewiseSampler, UniformSamplerWithNoise
from fruitnerf_b200.data.fruit_datamanager import FruitDataManager, FruitDataManagerConfig, get_corners_of_aabb, sample_surface_points
from fruitnerf_b200.export.exporter_utils import write_ply
from fruitnerf_b200.fruit_field import FruitField, SceneContraction
from fruitnerf_b200.fruit_nerf import FruitModel, FruitNerfModel, FruitNerfModelConfig
from fruitnerf_b200.fruit_nerf_config import METHODS
from oracle import ns_torch as ns

ROOT = Path(__file__).resolve().parent.parent


def _field(**kw):
    return FruitField(torch.tensor([[-1.0, -1, -1], [1, 1, 1]]), num_images=5, use_semantics=True, num_semantic_classes=1,
                      log2_hashmap_size=10, spatial_distortion=SceneContraction(), **kw)


def test_state_dict_keys_match_reference_naming():
    keys = set(_field().state_dict().keys())
    want = {
        "aabb", "max_res", "num_levels", "log2_hashmap_size", "embedding_appearance.embedding.weight",
        "mlp_base_grid.hash_table", "mlp_base.0.hash_table",
        "mlp_base_mlp.layers.0.weight", "mlp_base_mlp.layers.0.bias", "mlp_base_mlp.layers.1.weight", "mlp_base_mlp.layers.1.bias",
        "mlp_base.1.layers.0.weight", "mlp_base.1.layers.0.bias", "mlp_base.1.layers.1.weight", "mlp_base.1.layers.1.bias",
        "mlp_semantics.layers.0.weight", "mlp_semantics.layers.0.bias", "mlp_semantics.layers.1.weight", "mlp_semantics.layers.1.bias",
        "field_head_semantics.net.weight", "field_head_semantics.net.bias",
        "mlp_head.layers.0.weight", "mlp_head.layers.0.bias", "mlp_head.layers.1.weight", "mlp_head.layers.1.bias",
        "mlp_head.layers.2.weight", "mlp_head.layers.2.bias",
    }
    assert keys == want
    f = _field()
    f2 = _field()
    f2.load_state_dict(f.state_dict(), strict=True)  # fruit_pipeline.py:240
    assert torch.equal(f2.mlp_base_grid.hash_table, f.mlp_base_grid.hash_table)


def test_field_shapes_and_scalings_follow_reference_defaults():
    f = _field()
    shp = f.kernel_shape()
    assert shp.base_dims == [32, 64, 16] and shp.semantic_dims == [15, 64, 64] and shp.color_dims == [63, 64, 64, 3]
    assert list(shp.scalings) == ns.hash_scalings(16, 16, 2048).tolist()
    big = _field(geo_feat_dim=30, num_layers_semantic=3, hidden_dim_semantics=128, max_res=4096)
    assert big.kernel_shape().semantic_dims == [30, 128, 128, 64] and big.kernel_shape().color_dims == [78, 64, 64, 3]
    assert f.position_mode() == 0 and f.train().appearance_mode() == 0 and f.eval().appearance_mode() == 2
    f.test_mode = "export"
    assert f.appearance_mode() == 1


def test_holders_have_no_python_forward():
    f = _field()
    with pytest.raises(RuntimeError, match="no PyTorch fallback"):
        f.mlp_base_mlp(torch.zeros(1, 32))


def test_method_configs_carry_reference_hyperparameters():
    small, big, huge = METHODS["fruit_nerf"], METHODS["fruit_nerf_big"], METHODS["fruit_nerf_huge"]
    assert small.max_num_iterations == 30000 and small.pipeline.datamanager.train_num_rays_per_batch == 4096
    assert big.pipeline.model.log2_hashmap_size == 21 and big.pipeline.model.max_res == 4096 and big.pipeline.model.geo_feat_dim == 30
    assert big.pipeline.model.num_nerf_samples_per_ray == 128 and huge.pipeline.model.num_nerf_samples_per_ray == 64
    assert huge.pipeline.model.max_res == 8192 and huge.pipeline.datamanager.train_num_rays_per_batch == 16384
    assert small.optimizers["fields"]["optimizer"]["type"] == "Adam" and big.optimizers["fields"]["optimizer"]["type"] == "RAdam"
    assert FruitNerfModel is FruitModel


def test_method_specification_against_a_stand_in_nerfstudio(monkeypatch):
    """The ``MethodSpecification`` branch of fruit_nerf_config (taken when nerfstudio is importable, which it is not in this
    image): run it against stand-in modules with the constructor signatures of nerfstudio 0.3.2's config classes and check
    every value the reference sets (fruit_nerf/fruit_nerf_config.py:27-61, 63-111, 113-164)."""
    import sys
    import types
    from dataclasses import dataclass, field as dfield
    from typing import Any, Dict, Optional

    from fruitnerf_b200 import fruit_nerf_config as fc

    @dataclass
    class ViewerConfig:
        num_rays_per_chunk: int = 32768

    @dataclass
    class AdamOptimizerConfig:
        lr: float = 0.0005
        eps: float = 1e-08
        weight_decay: float = 0

    @dataclass
    class RAdamOptimizerConfig(AdamOptimizerConfig):
        pass

    @dataclass
    class ExponentialDecaySchedulerConfig:
        lr_final: float = 0.000005
        max_steps: int = 100000

    @dataclass
    class TrainerConfig:
        method_name: Optional[str] = None
        steps_per_eval_batch: int = 500
        steps_per_save: int = 1000
        max_num_iterations: int = 1000000
        mixed_precision: bool = False
        pipeline: Any = None
        optimizers: Dict[str, Any] = dfield(default_factory=dict)
        viewer: Any = None
        vis: str = "wandb"

    @dataclass
    class MethodSpecification:
        config: Any
        description: str

    tree = {
        "nerfstudio": {}, "nerfstudio.configs": {}, "nerfstudio.engine": {}, "nerfstudio.plugins": {},
        "nerfstudio.configs.base_config": {"ViewerConfig": ViewerConfig},
        "nerfstudio.engine.optimizers": {"AdamOptimizerConfig": AdamOptimizerConfig, "RAdamOptimizerConfig": RAdamOptimizerConfig},
        "nerfstudio.engine.schedulers": {"ExponentialDecaySchedulerConfig": ExponentialDecaySchedulerConfig},
        "nerfstudio.engine.trainer": {"TrainerConfig": TrainerConfig},
        "nerfstudio.plugins.types": {"MethodSpecification": MethodSpecification},
    }
    for name, attrs in tree.items():
        mod = types.ModuleType(name)
        mod.__dict__.update(attrs)
        monkeypatch.setitem(sys.modules, name, mod)

    want = {"fruit_nerf": (30000, 1 << 13, AdamOptimizerConfig, 200000, 200000), "fruit_nerf_big": (100000, 1 << 15, RAdamOptimizerConfig, None, 50000),
            "fruit_nerf_huge": (100000, 1 << 15, RAdamOptimizerConfig, None, 50000)}
    for name, (iters, chunk, opt_cls, prop_steps, field_steps) in want.items():
        ms = fc.method_specification(fc.METHODS[name])
        assert isinstance(ms, MethodSpecification) and ms.description.startswith("Base config for FruitNeRF")
        c = ms.config
        assert (c.method_name, c.max_num_iterations, c.steps_per_eval_batch, c.steps_per_save, c.mixed_precision, c.vis) == (
            name, iters, 500, 2000, True, "viewer")
        assert c.viewer.num_rays_per_chunk == chunk
        assert c.pipeline is fc.METHODS[name].pipeline and c.pipeline._target.__name__ == "FruitPipeline"
        assert set(c.optimizers) == {"proposal_networks", "fields"}
        for group, steps in (("proposal_networks", prop_steps), ("fields", field_steps)):
            o = c.optimizers[group]
            assert type(o["optimizer"]) is opt_cls and o["optimizer"].lr == 1e-2 and o["optimizer"].eps == 1e-15
            if steps is None:
                assert o["scheduler"] is None  # fruit_nerf_config.py:90-94, 140-144: no scheduler on the proposal group
            else:
                assert o["scheduler"].lr_final == 1e-4 and o["scheduler"].max_steps == steps


def test_model_requires_semantics_metadata_and_builds():
    cfg = FruitNerfModelConfig()
    sem = Semantics(filenames=[], classes=["fruit"], colors=torch.tensor([[0.0, 0, 0], [1.0, 0, 0]]))
    with pytest.raises(AssertionError):
        FruitModel(cfg, metadata={}, scene_box=SceneBox(torch.tensor([[-1.0, -1, -1], [1, 1, 1]])), num_train_data=3, test_mode="val")
    cfg.log2_hashmap_size = 10
    m = FruitModel(cfg, metadata={"semantics": sem}, scene_box=SceneBox(torch.tensor([[-1.0, -1, -1], [1, 1, 1]])), num_train_data=3,
                   test_mode="val")
    assert set(m.get_param_groups().keys()) == {"proposal_networks", "fields"}
    assert len(m.get_param_groups()["fields"]) == len(list(m.field.parameters()))
    m.setup_inference(render_rgb=True, num_inference_samples=20)
    assert isinstance(m.proposal_sampler, UniformSamplerWithNoise) and m.field.spatial_distortion is None
    rb = RayBundle(origins=torch.zeros(4, 3), directions=torch.ones(4, 3))
    rb = m.collider(rb)
    assert torch.all(rb.nears == 0.05) and torch.all(rb.fars == 1000.0)
    # nerfstudio NearFarCollider: the near plane only applies while training; given nears / fars are kept
    m.eval()
    rb = m.collider(RayBundle(origins=torch.zeros(4, 3), directions=torch.ones(4, 3)))
    assert torch.all(rb.nears == 0.0) and torch.all(rb.fars == 1000.0)
    rb = m.collider(RayBundle(origins=torch.zeros(4, 3), directions=torch.ones(4, 3), nears=torch.full((4, 1), 0.3), fars=torch.full((4, 1), 2.0)))
    assert torch.all(rb.nears == 0.3) and torch.all(rb.fars == 2.0)
    m.train()
    cfg2 = FruitNerfModelConfig(use_gradient_scaling=True)
    cfg2.log2_hashmap_size = 10
    with pytest.raises(NotImplementedError):
        FruitModel(cfg2, metadata={"semantics": sem}, scene_box=SceneBox(torch.tensor([[-1.0, -1, -1], [1, 1, 1]])), num_train_data=3, test_mode="val")


def test_export_grid_matches_oracle_restatement():
    for n, aabb in ((6, ((-1, -1, -1), (1, 1, 1))), (10, ((-1.0, -0.5, -0.25), (1.0, 0.5, 0.75)))):
        corners = get_corners_of_aabb(aabb, "cpu")
        pts, plane = sample_surface_points(corners, n, "cpu")
        pts_ref, plane_ref = ns.surface_points(aabb, n)
        assert torch.equal(pts, pts_ref) and torch.equal(plane, plane_ref)
    dm = FruitDataManager(FruitDataManagerConfig(eval_num_rays_per_batch=10), device="cpu")
    assert dm.setup_inference(aabb=((-1, -1, -1), (1, 1, 1)), num_points=5) == 25
    sizes = []
    for _ in range(3):
        rb, _ = dm.next_sample_volume(0)
        sizes.append(len(rb))
        assert torch.all(rb.nears == 0) and torch.allclose(rb.fars, torch.full_like(rb.fars, 2.0))
        assert torch.allclose(rb.directions, torch.tensor([[0.0, 0.0, 1.0]]).expand(len(rb), 3))
    assert sizes == [10, 10, 5]


def test_samplers_match_oracle_bins():
    rb = RayBundle(origins=torch.zeros(3, 3), directions=torch.ones(3, 3), nears=torch.zeros(3, 1), fars=torch.full((3, 1), 2.0))
    s = UniformSamplerWithNoise(num_samples=7).eval()(rb)
    st, en = ns.uniform_bins(rb.nears, rb.fars, 7)
    assert torch.equal(s.frustums.starts, st) and torch.equal(s.frustums.ends, en)
    assert torch.equal(s.deltas, en - st) and s.frustums.origins.shape == (3, 7, 3)
    p = UniformLinDispPiecewiseSampler(num_samples=8).eval()(RayBundle(origins=torch.zeros(2, 3), directions=torch.ones(2, 3),
                                                                     nears=torch.full((2, 1), 0.05), fars=torch.full((2, 1), 1000.0)))
    e = torch.cat([p.frustums.starts[0, :, 0], p.frustums.ends[0, -1:, 0]])
    assert abs(float(e[0]) - 0.05) < 1e-6 and abs(float(e[-1]) - 1000.0) < 0.1 and bool((e[1:] > e[:-1]).all())


def test_ply_writer_roundtrip(tmp_path):
    pts = np.array([[0.0, 1.0, 2.0], [3.0, 4.0, 5.0]])
    col = np.array([[1.0, 0.5, 0.0], [0.0, 0.0, 1.0]])
    path = tmp_path / "a" / "cloud.ply"
    write_ply(path, pts, col)
    raw = path.read_bytes()
    head, body = raw.split(b"end_header\n")
    assert b"element vertex 2" in head and b"property double x" in head and b"property uchar red" in head
    rec = np.frombuffer(body, dtype=[("x", "<f8"), ("y", "<f8"), ("z", "<f8"), ("r", "u1"), ("g", "u1"), ("b", "u1")])
    assert np.allclose(np.stack([rec["x"], rec["y"], rec["z"]], -1), pts) and rec["r"].tolist() == [255, 0] and rec["g"][0] == 127


WORKER = r"""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from fruitnerf_b200 import ops
from fruitnerf_b200.fruit_pipeline import sync_gradients
dist.init_process_group("gloo", rank=int(os.environ["RANK"]), world_size=int(os.environ["WORLD_SIZE"]))
rank, world = dist.get_rank(), dist.get_world_size()
params = [torch.zeros(5, 2), torch.zeros(3), torch.zeros(7, 7)]
flat, views = ops.flat_zero_grads(params)
assert flat.numel() == 12 + 4 + 52 and all(v.data_ptr() % 16 == 0 for v in views)
for i, v in enumerate(views):
    v += float(rank + 1) * (i + 1)       # what the backward kernels would have accumulated
sync_gradients(flat, world)
for i, v in enumerate(views):
    want = (i + 1) * sum(r + 1 for r in range(world)) / world   # DDP mean
    assert torch.allclose(v, torch.full_like(v, want)), (rank, i, v)
# ray sharding: rank r owns rays [r*R, (r+1)*R) of the global batch, no overlap, full cover
R = 8
own = torch.arange(rank * R, (rank + 1) * R)
allr = [torch.zeros(R, dtype=torch.long) for _ in range(world)]
dist.all_gather(allr, own)
assert torch.equal(torch.cat(allr), torch.arange(world * R))
# trainer gradient exchange: views of one flat buffer -> one collective; unrelated tensors -> coalesced
from fruitnerf_b200.trainer import Trainer
class _T:
    world_size = world
flat = torch.arange(10.0) * (rank + 1)
views = [flat[:4].view(2, 2), flat[4:10]]
Trainer._exchange(_T(), views)
assert torch.allclose(flat, torch.arange(10.0) * (sum(range(1, world + 1)) / world))
loose = [torch.full((3,), float(rank)), torch.full((2, 2), 10.0 * rank)]
Trainer._exchange(_T(), loose)
assert torch.allclose(loose[0], torch.full((3,), (world - 1) / 2)) and torch.allclose(loose[1], torch.full((2, 2), 10.0 * (world - 1) / 2))
# export sharding: contiguous slabs cover the ray grid once; merged point lists are identical on every rank
from fruitnerf_b200.export.exporter_utils import export_slab, merge_export_shards
N = 1001
slabs = [export_slab(N, world, r) for r in range(world)]
assert slabs[0][0] == 0 and slabs[-1][1] == N and all(slabs[i][1] == slabs[i + 1][0] for i in range(world - 1))
lo, hi = slabs[rank]
keys = torch.arange(lo, hi, dtype=torch.int64)[::7]
rows = torch.stack([keys.float() * k for k in range(7)], dim=1)
merged = merge_export_shards({"density": (rows, keys)}, world)
mk = merged["density"][1]
want = torch.cat([torch.arange(*export_slab(N, world, r), dtype=torch.int64)[::7] for r in range(world)])
assert torch.equal(torch.sort(mk).values, torch.sort(want).values)
assert torch.equal(merged["density"][0][:, 1].long(), mk)
# gradient-exchange factory: the exchange owns the flat buffer; on CPU tensors "auto" is the process group's all-reduce (mean);
# the NVLS kernel needs CUDA + multicast memory and is refused when asked for explicitly
from fruitnerf_b200.grad_exchange import make_gradient_exchange
from fruitnerf_b200 import ops
ex = make_gradient_exchange(ops.flat_grad_numel([torch.zeros(5), torch.zeros(3, 3)]), world, "cpu", kind="auto")
assert ex.flat.numel() == 8 + 12 and ex.kind == "nccl" and ex.describe()["bytes"] == 80
ex.flat.fill_(float(rank + 1))
ex()
assert torch.allclose(ex.flat, torch.full_like(ex.flat, sum(range(1, world + 1)) / world))
flat, views = ops.flat_zero_grads([torch.zeros(5), torch.zeros(3, 3)], out=ex.flat)  # persistent buffer is re-zeroed, views alias it
assert flat.data_ptr() == ex.flat.data_ptr() and float(flat.abs().sum()) == 0 and views[1].shape == (3, 3)
assert make_gradient_exchange(8, 1, "cpu") is None
dist.barrier()
print("rank", rank, "ok")
"""


def test_gradient_exchange_factory_world2_gloo(tmp_path):
    """Covered by the same two-process worker as below (kept as its own test id for the coverage table): see WORKER."""
    assert "make_gradient_exchange" in WORKER and "flat_zero_grads" in WORKER


def test_flat_gradient_allreduce_world2_gloo(tmp_path):
    """N>1 path on CPU: flat gradient buffer + mean all-reduce (the DDP exchange of
    fruit_pipeline.py:117) and the per-rank ray partition, world_size 2 over gloo."""
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, str(script), str(ROOT)], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    for p in procs:
        out, _ = p.communicate(timeout=120)
        assert p.returncode == 0, out


def test_bf16_hi_lo_split_error_model():
    """The numerical contract bench.py states in ``dtype``: the tensor-core MLPs multiply fp32 values as bf16 hi/lo pairs
    (x = hi + lo, hi = bf16(x), lo = bf16(x - hi)) with three MMAs per product (hi*hi + lo*hi + hi*lo, fp32 accumulate; the lo*lo
    term is dropped).  Emulated here in torch: the result stays within ~2^-16 of the exact product relative to sum |a||b| -- two
    orders of magnitude inside the 1e-3 parity bar, and ~250x tighter than a single bf16 MMA."""
    g = torch.Generator().manual_seed(0)
    a = torch.randn(256, 64, generator=g)
    b = torch.randn(64, 48, generator=g) * 0.3
    split = lambda x: (x.bfloat16().float(), (x - x.bfloat16().float()).bfloat16().float())  # noqa: E731
    (ah, al), (bh, bl) = split(a), split(b)
    got = ah @ bh + al @ bh + ah @ bl
    exact = a.double() @ b.double()
    scale = a.abs().double() @ b.abs().double()
    err3 = float(((got.double() - exact).abs() / scale).max())
    err1 = float((((ah @ bh).double() - exact).abs() / scale).max())
    assert err3 < 2.0 ** -15, err3
    assert err1 > 50 * err3  # what a plain bf16 product would cost
