"""Method specifications of the plugin (fruit_nerf/fruit_nerf_config.py:27-164).

``TrainerSpec`` restates the TrainerConfig fields the reference sets as plain dataclasses: it is what
``fruitnerf_b200.trainer.Trainer`` consumes and what ``METHODS`` holds, with or without nerfstudio.

The ``nerfstudio.method_configs`` entry points (pyproject.toml) resolve to the module attributes
``fruit_nerf_method`` / ``fruit_nerf_method_big`` / ``fruit_nerf_method_huge``.  When nerfstudio is importable
these are ``MethodSpecification(config=TrainerConfig(...))`` objects built by ``method_specification`` (same
``method_name``, iteration counts, pipeline config and per-group optimizer / scheduler configs as the reference,
fruit_nerf_config.py:27-61 / 63-111 / 113-164), so ``ns-train fruit_nerf`` discovers them and drives
``FruitPipeline`` with its own trainer and torch optimizers (the kernels expose ordinary ``.grad``s through
``ops.render``'s autograd function).  Without nerfstudio the same names are the plain ``TrainerSpec``s.  The
nerfstudio branch cannot be exercised in this image (the package is not installable offline): INTEGRATION.md."""
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
    viewer_num_rays_per_chunk: int = 1 << 15  # ViewerConfig(num_rays_per_chunk=...): 1 << 13 for fruit_nerf (:57), 1 << 15 for big / huge (:107, :160)


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
    viewer_num_rays_per_chunk=1 << 13,
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


def method_specification(spec: TrainerSpec):
    """``MethodSpecification`` of a TrainerSpec for nerfstudio's plugin discovery (fruit_nerf_config.py:27, 63, 113)."""
    from nerfstudio.configs.base_config import ViewerConfig  # type: ignore
    from nerfstudio.engine.optimizers import AdamOptimizerConfig, RAdamOptimizerConfig  # type: ignore
    from nerfstudio.engine.schedulers import ExponentialDecaySchedulerConfig  # type: ignore
    from nerfstudio.engine.trainer import TrainerConfig  # type: ignore
    from nerfstudio.plugins.types import MethodSpecification  # type: ignore

    kinds = {"Adam": AdamOptimizerConfig, "RAdam": RAdamOptimizerConfig}
    optimizers = {}
    for group, o in spec.optimizers.items():
        sched = o["scheduler"]
        optimizers[group] = {
            "optimizer": kinds[o["optimizer"]["type"]](lr=o["optimizer"]["lr"], eps=o["optimizer"]["eps"]),
            "scheduler": None if sched is None else ExponentialDecaySchedulerConfig(lr_final=sched["lr_final"], max_steps=sched["max_steps"]),
        }
    config = TrainerConfig(
        method_name=spec.method_name,
        steps_per_eval_batch=spec.steps_per_eval_batch,
        steps_per_save=spec.steps_per_save,
        max_num_iterations=spec.max_num_iterations,
        mixed_precision=spec.mixed_precision,
        pipeline=spec.pipeline,
        optimizers=optimizers,
        viewer=ViewerConfig(num_rays_per_chunk=spec.viewer_num_rays_per_chunk),
        vis="viewer",
    )
    return MethodSpecification(config=config, description=spec.description)


try:  # pragma: no cover - nerfstudio is not installable in this image
    import nerfstudio.plugins.types  # type: ignore  # noqa: F401

    fruit_nerf_method, fruit_nerf_method_big, fruit_nerf_method_huge = (
        method_specification(METHODS[n]) for n in ("fruit_nerf", "fruit_nerf_big", "fruit_nerf_huge"))
except ImportError:
    pass
