"""Method specifications of the plugin (fruit_nerf/fruit_nerf_config.py:27-164), restated as plain
dataclasses.  With nerfstudio installed these are wrapped into ``MethodSpecification`` objects and
discovered through the ``nerfstudio.method_configs`` entry points (pyproject.toml); without it they
still carry every hyper-parameter of the three shipped configs."""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Any, Dict

from .data.fruit_datamanager import FruitDataManagerConfig
from .fruit_nerf import FruitNerfModelConfig
from .fruit_pipeline import FruitPipelineConfig


@dataclass
class TrainerSpec:
    """The TrainerConfig fields the reference sets (fruit_nerf_config.py:28-59)."""

    method_name: str
    max_num_iterations: int
    pipeline: FruitPipelineConfig
    optimizers: Dict[str, Any]
    steps_per_eval_batch: int = 500
    steps_per_save: int = 2000
    mixed_precision: bool = True
    description: str = ""


def _opt(kind: str, lr_final, max_steps):
    sched = None if lr_final is None else {"type": "ExponentialDecay", "lr_final": lr_final, "max_steps": max_steps}
    return {"optimizer": {"type": kind, "lr": 1e-2, "eps": 1e-15}, "scheduler": sched}


fruit_nerf_method = TrainerSpec(
    method_name="fruit_nerf",
    max_num_iterations=30000,
    pipeline=FruitPipelineConfig(
        datamanager=FruitDataManagerConfig(train_num_rays_per_batch=4096, eval_num_rays_per_batch=4096),
        model=FruitNerfModelConfig(eval_num_rays_per_chunk=1 << 15),
    ),
    optimizers={"proposal_networks": _opt("Adam", 1e-4, 200000), "fields": _opt("Adam", 1e-4, 200000)},
    description="Base config for FruitNeRF",
)

_big_model = dict(
    eval_num_rays_per_chunk=1 << 15,
    num_nerf_samples_per_ray=128,
    num_proposal_samples_per_ray=(512, 256),
    hidden_dim=128,            # dead upstream: never forwarded to FruitField (fruit_nerf.py:88-103)
    geo_feat_dim=30,
    hidden_dim_color=128,      # dead upstream
    hidden_dim_semantics=128,
    num_layers_semantic=3,
    appearance_embed_dim=128,  # dead upstream
    max_res=4096,
    proposal_weights_anneal_max_num_iters=5000,
    log2_hashmap_size=21,
)

fruit_nerf_method_big = TrainerSpec(
    method_name="fruit_nerf_big",
    max_num_iterations=100000,
    pipeline=FruitPipelineConfig(
        datamanager=FruitDataManagerConfig(train_num_rays_per_batch=4096 * 2, eval_num_rays_per_batch=4096),
        model=FruitNerfModelConfig(**_big_model),
    ),
    optimizers={"proposal_networks": _opt("RAdam", None, None), "fields": _opt("RAdam", 1e-4, 50000)},
    description="Base config for FruitNeRF-Big",
)

_huge_model = dict(_big_model, num_nerf_samples_per_ray=64, num_proposal_samples_per_ray=(512, 512), max_res=8192,
                   proposal_net_args_list=[
                       {"hidden_dim": 16, "log2_hashmap_size": 17, "num_levels": 5, "max_res": 512, "use_linear": False},
                       {"hidden_dim": 16, "log2_hashmap_size": 17, "num_levels": 7, "max_res": 2048, "use_linear": False},
                   ])

fruit_nerf_method_huge = TrainerSpec(
    method_name="fruit_nerf_huge",
    max_num_iterations=100000,
    pipeline=FruitPipelineConfig(
        datamanager=FruitDataManagerConfig(train_num_rays_per_batch=4096 * 4, eval_num_rays_per_batch=4096),
        model=FruitNerfModelConfig(**_huge_model),
    ),
    optimizers={"proposal_networks": _opt("RAdam", None, None), "fields": _opt("RAdam", 1e-4, 50000)},
    description="Base config for FruitNeRF-Huge",
)

METHODS = {m.method_name: m for m in (fruit_nerf_method, fruit_nerf_method_big, fruit_nerf_method_huge)}
