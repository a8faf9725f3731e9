"""Seeded end-to-end cases evaluated with the oracle (TEST INFRASTRUCTURE, see ``oracle/__init__.py``).

``run_case`` builds the oracle FruitModel from a host-generated state dict (the same tensors the product loads), runs
the reference-shaped render / training step and returns everything the parity tests compare.  ``oracle/make_golden.py``
freezes small cases into ``tests/golden/*.npz``.
"""
from __future__ import annotations

import os
import sys
from typing import Dict, Optional

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from cropnerf_b200 import synthetic  # noqa: E402  (pure host-side data generation, shared with the product)

from . import fruit_torch as ft  # noqa: E402
from . import nerfstudio_torch as ns  # noqa: E402

CASES = {
    # name: (num_rays, config overrides, table_scale)
    "tiny_eval": dict(num_rays=96, training=False, cfg=dict(log2_hashmap_size=14), table_scale=0.5),
    "tiny_train": dict(num_rays=96, training=True, cfg=dict(log2_hashmap_size=14), table_scale=0.5),
    "tiny_aabb_eval": dict(num_rays=64, training=False, cfg=dict(log2_hashmap_size=14, disable_scene_contraction=True), table_scale=0.5),
}


def small_prop_args():
    return [
        {"hidden_dim": 16, "log2_hashmap_size": 13, "num_levels": 5, "max_res": 128, "use_linear": False},
        {"hidden_dim": 16, "log2_hashmap_size": 13, "num_levels": 5, "max_res": 256, "use_linear": False},
    ]


def make_config(overrides: Optional[dict] = None, small: bool = True) -> ft.FruitNerfModelConfig:
    kw = dict(overrides or {})
    if small and "proposal_net_args_list" not in kw:
        kw["proposal_net_args_list"] = small_prop_args()
    return ft.FruitNerfModelConfig(**kw)


def build_oracle(cfg: ft.FruitNerfModelConfig, num_images: int, seed: int, table_scale: float, dtype=torch.float32):
    torch.manual_seed(seed)
    model = ft.FruitModel(cfg, num_train_data=num_images)
    state = synthetic.randomize_state(model.state_dict(), seed=seed, table_scale=table_scale)
    model.load_state_dict(state)
    if dtype != torch.float32:
        model = model.to(dtype)
    return model, state


def oracle_bundle(rays: Dict[str, torch.Tensor], dtype=torch.float32, with_near_far: Optional[tuple] = None) -> ns.RayBundle:
    rb = ns.RayBundle(
        origins=rays["origins"].to(dtype),
        directions=rays["directions"].to(dtype),
        pixel_area=rays["pixel_area"].to(dtype),
        camera_indices=rays["camera_indices"],
    )
    if with_near_far is not None:
        rb.nears = torch.full_like(rb.pixel_area, with_near_far[0])
        rb.fars = torch.full_like(rb.pixel_area, with_near_far[1])
    return rb


def run_case(name: str, num_images: int = 20, seed: int = 0, spec: Optional[dict] = None) -> Dict[str, np.ndarray]:
    spec = spec or CASES[name]
    R = spec["num_rays"]
    cfg = make_config(spec.get("cfg"))
    model, state = build_oracle(cfg, num_images, seed, spec["table_scale"])
    rays = synthetic.make_rays(R, seed=1, num_cameras=num_images)
    targets = synthetic.make_targets(R, seed=3)
    out: Dict[str, np.ndarray] = {}
    if spec["training"]:
        model.train()
        feed = synthetic.JitterFeed(synthetic.make_jitter(R, 3, seed=2))
        model.proposal_sampler.initial_sampler.rand_fn = feed
        model.proposal_sampler.pdf_sampler.rand_fn = feed
        model.set_anneal(500)  # mid-anneal so weights**anneal is exercised
    else:
        model.eval()
    outputs = model(oracle_bundle(rays))
    for k in ("rgb", "depth", "accumulation", "semantics", "prop_depth_0", "prop_depth_1", "semantics_colormap"):
        out[k] = outputs[k].detach().numpy()
    out["pdf_inds_last"] = model.proposal_sampler.pdf_sampler.last_inds.numpy().astype(np.int32)
    out["median_index"] = model.renderer_depth.last_median_index.numpy().astype(np.int32)
    if spec["training"]:
        loss_dict = model.get_loss_dict(outputs, targets)
        metrics = model.get_metrics_dict(outputs, targets)
        loss = sum(loss_dict.values())
        loss.backward()
        for k, v in loss_dict.items():
            out["loss_" + k] = v.detach().numpy()
        out["metric_distortion"] = metrics["distortion"].detach().numpy()
        out["metric_psnr"] = metrics["psnr"].detach().numpy()
        # gradient fingerprints: norms + a strided sample (full tables are too big for a fixture)
        for pname, p in model.named_parameters():
            if p.grad is None or pname.startswith("field.mlp_base.0") or ".mlp_base.0." in pname:
                continue
            g = p.grad.detach().reshape(-1)
            out["gradnorm/" + pname] = np.asarray(g.double().norm().item())
            if g.numel() <= 4096:
                out["grad/" + pname] = g.numpy()
    return out
