"""CPU oracle for the FruitNeRF hot path -- TEST INFRASTRUCTURE, NOT PRODUCT.

Pure-PyTorch fp32 restatement of the reference's per-ray
sample -> encode -> MLP -> composite path (fruit_nerf/fruit_field.py:168-301,
fruit_nerf/fruit_nerf.py:251-269,316-372) and of the uniform-volume export
(fruit_nerf/components/*.py, fruit_nerf/data/fruit_datamanager.py:42-121,157-172,
fruit_nerf/export/exporter_utils.py:100-153).

PARITY UNPINNED.  The arithmetic of this path lives in third-party packages that are
not vendored under /root/reference and are not installable here (no network):
``nerfstudio==0.3.2`` (pinned at reference pyproject.toml:10) and an un-pinned
``tinycudann``.  The reference itself ships no tests, golden vectors or fixtures
(SURVEY.md section 4), so nothing upstream can pin this restatement.  What is restated
is nerfstudio 0.3.2's *torch-fallback* semantics (HashEncoding.pytorch_fwd, MLP,
SHEncoding, trunc_exp, SceneContraction, RaySamples.get_weights, the renderers and
samplers); every function cites the reference call site it follows.  The golden
vectors under tests/golden/ are produced by THIS oracle (tests/golden/make_golden.py).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` legs may import this package.  The product package
``fruitnerf_b200`` never imports it and has no CPU fallback.
"""
