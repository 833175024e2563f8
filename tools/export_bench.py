"""Uniform-volume export throughput (BASELINE.json configs[4]): N^3 samples of the fruit_nerf field through
fnr_export_forward in batches of 32768 rays, tensor-core kernel vs the fp32 simt kernel, plus the
size-independent checks the full grid allows (set inclusions, unique keys, cross-implementation counts).

    python tools/export_bench.py [--n 512] [--batch 32768] [--json gpurun_out/export_bench.json]
"""
from __future__ import annotations

import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))

from fruitnerf_b200 import _lib as L  # noqa: E402
from fruitnerf_b200 import ops  # noqa: E402
from fruitnerf_b200 import synthetic as syn  # noqa: E402
from fruitnerf_b200.fruit_field import FruitField  # noqa: E402


def surface_grid(n: int, dev):
    """fruit_datamanager.py:71-121 for the cube [-1,1]^3: n x n points on the z = -1 face, x-major."""
    lin = torch.linspace(-1.0, 1.0, n)
    gx, gy = torch.meshgrid(lin, lin, indexing="ij")
    pts = torch.stack([gx.reshape(-1), gy.reshape(-1), torch.full((n * n,), -1.0)], dim=-1)
    return pts.to(dev), (0.0, 0.0, 1.0), 2.0


def run(field, pts, normal, far, n, batch, impl, thresholds, capacity):
    dev = pts.device
    bins = torch.linspace(0.0, 1.0, n + 1).to(dev)
    buf = ops.ExportBuffers(capacity=capacity, device=dev)
    shape, params = field.kernel_shape(), field.kernel_params()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    torch.cuda.synchronize()
    ev[0].record()
    done = 0
    while done < pts.shape[0]:
        o = pts[done:done + batch]
        ops.export_batch(shape, params, o, normal, bins, 0.0, far, buf, point_base=done * n, dense_out=False, thresholds=thresholds,
                         impl=impl)
        done += o.shape[0]
    ev[1].record()
    torch.cuda.synchronize()
    return ev[0].elapsed_time(ev[1]), buf


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=512)
    ap.add_argument("--batch", type=int, default=32768)
    ap.add_argument("--json", default=None)
    ap.add_argument("--skip-simt", action="store_true")
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    v = dict(syn.SMALL)
    sd = syn.field_state(geo=v["geo"], sem_dims=v["sem_dims"], log2_hashmap_size=v["log2_hashmap_size"], num_images=7, table_scale=2.0,
                         weight_gain=2.5)
    field = FruitField(aabb=sd["aabb"], num_images=7, geo_feat_dim=v["geo"], max_res=v["max_res"], log2_hashmap_size=v["log2_hashmap_size"],
                       num_layers_semantic=len(v["sem_dims"]) - 1, hidden_dim_semantics=v["sem_dims"][1], use_semantics=True,
                       num_semantic_classes=1, test_mode="export", spatial_distortion=None)
    field.load_state_dict(sd, strict=False)
    field = field.to(dev).eval()
    pts, normal, far = surface_grid(a.n, dev)
    total = a.n ** 3

    # pick thresholds from a probe batch so that the three sets are populated (random weights never reach 70 / 3)
    probe_buf = ops.ExportBuffers(capacity=1, device=dev)
    bins = torch.linspace(0.0, 1.0, a.n + 1).to(dev)
    mid = (a.n * a.n // 2 // 2048) * 2048
    dense = ops.export_batch(field.kernel_shape(), field.kernel_params(), pts[mid:mid + 2048], normal, bins, 0.0, far, probe_buf,
                             dense_out=True)
    thr = (float(dense["semantics"].quantile(0.97)), float(dense["density"].quantile(0.97)), 0.5)
    capacity = min(total, 1 << 25)

    res = {"n": a.n, "points": total, "batch_rays": a.batch, "thresholds": thr}
    bufs = {}
    for name, impl in (("tcgen05", L.FNR_IMPL_TCGEN05), ("simt", L.FNR_IMPL_SIMT)):
        if name == "simt" and a.skip_simt:
            continue
        run(field, pts[: a.batch], normal, far, a.n, a.batch, impl, thr, capacity)  # warm-up
        ms, buf = run(field, pts, normal, far, a.n, a.batch, impl, thr, capacity)
        counts = buf.counts.cpu().tolist()
        res[name] = {"ms": ms, "points_per_s": total / (ms * 1e-3), "counts": counts}
        bufs[name] = (buf, counts)
        print(f"{name}: {ms:.1f} ms  {total / ms * 1e-6:.2f} Gpoints/s  counts {counts}", flush=True)

    # size-independent properties of the tensor-core result
    buf, counts = bufs["tcgen05"]
    keys = [buf.keys[k][: min(counts[k], capacity)] for k in range(3)]
    sets_ok = True
    for k in range(3):
        u = torch.unique(keys[k])
        sets_ok &= bool(u.numel() == keys[k].numel()) and (keys[k].numel() == 0 or int(u.max()) < total)
    s2 = torch.sort(keys[2]).values
    for k in (0, 1):  # every semantic selection also passed the density threshold
        if keys[k].numel():
            pos = torch.searchsorted(s2, keys[k]).clamp_(max=max(s2.numel() - 1, 0))
            sets_ok &= bool((s2[pos] == keys[k]).all())
    res["keys_unique_and_nested"] = bool(sets_ok)
    if "simt" in bufs:
        c2 = bufs["simt"][1]
        res["count_diff_vs_simt"] = [abs(x - y) for x, y in zip(counts, c2)]
        res["count_rel_diff_vs_simt"] = [abs(x - y) / max(y, 1) for x, y in zip(counts, c2)]
    # algorithmic hash bytes: 1024 B / point (16 levels x 8 corners x 8 B)
    res["tcgen05"]["hash_GBps"] = total * 1024 / (res["tcgen05"]["ms"] * 1e-3) / 1e9
    print(json.dumps(res))
    if a.json:
        os.makedirs(os.path.dirname(a.json), exist_ok=True)
        with open(a.json, "w") as f:
            json.dump(res, f, indent=1)


if __name__ == "__main__":
    main()
