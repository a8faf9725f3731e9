"""Seeded synthetic scenes for tests and benchmarks.

A 3DCotton-shaped scene: ``num_cameras`` pinhole cameras (1920x1080, fx=fy=1442.48, the intrinsics of the reference's
``utils/transforms.json`` template) on the unit sphere looking at the origin, poses scaled so max|t| = 1
(``cotton_nerf_dataparser.py:199-204``), scene box +-1 (``:215-220``); rays through uniformly drawn pixels; targets
``image ~ U(0,1)``, ``fruit_mask ~ Bernoulli(0.1)`` (binary masks, ``cotton_dataset.py:34-39``).  Everything is
generated on the host with explicit generators so the oracle and the CUDA path consume identical numbers.
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import numpy as np
import torch
from torch import Tensor

IMAGE_W, IMAGE_H = 1920, 1080
FOCAL = 1442.4757
NUM_CAMERAS = 300


def make_cameras(num_cameras: int = NUM_CAMERAS, seed: int = 1) -> Tensor:
    """camera-to-world [N,3,4] (OpenGL convention: camera looks down -z, +y up), centres on the unit sphere."""
    g = torch.Generator().manual_seed(seed)
    c = torch.randn((num_cameras, 3), generator=g)
    c = c / c.norm(dim=-1, keepdim=True)
    fwd = -c  # towards the origin
    up = torch.tensor([0.0, 0.0, 1.0]).expand_as(c)
    right = torch.cross(fwd, up, dim=-1)
    bad = right.norm(dim=-1, keepdim=True) < 1e-3
    right = torch.where(bad, torch.tensor([1.0, 0.0, 0.0]).expand_as(c), right)
    right = right / right.norm(dim=-1, keepdim=True)
    true_up = torch.cross(right, fwd, dim=-1)
    rot = torch.stack([right, true_up, -fwd], dim=-1)  # columns: x, y, z axes of the camera
    return torch.cat([rot, c[..., None]], dim=-1)


def make_rays(num_rays: int, seed: int = 1, num_cameras: int = NUM_CAMERAS, c2w: Optional[Tensor] = None,
              camera_indices: Optional[Tensor] = None, pixels: Optional[Tensor] = None) -> Dict[str, Tensor]:
    """Host tensors: origins [R,3], directions [R,3] (unit), pixel_area [R,1], camera_indices [R,1] int64."""
    g = torch.Generator().manual_seed(seed + 1000)
    if c2w is None:
        c2w = make_cameras(num_cameras, seed=1)
    n_cam = c2w.shape[0]
    if camera_indices is None:
        camera_indices = torch.randint(0, n_cam, (num_rays,), generator=g)
    if pixels is None:
        u = torch.rand((num_rays,), generator=g) * IMAGE_W
        v = torch.rand((num_rays,), generator=g) * IMAGE_H
    else:
        u, v = pixels[:, 0].float(), pixels[:, 1].float()
    d_cam = torch.stack([(u - IMAGE_W / 2) / FOCAL, -(v - IMAGE_H / 2) / FOCAL, -torch.ones_like(u)], dim=-1)
    rot = c2w[camera_indices, :, :3]
    d = torch.einsum("rij,rj->ri", rot, d_cam)
    norm = d.norm(dim=-1, keepdim=True)
    d = d / norm
    o = c2w[camera_indices, :, 3]
    pixel_area = (1.0 / FOCAL**2) / norm**3
    return {
        "origins": o.contiguous().float(),
        "directions": d.contiguous().float(),
        "pixel_area": pixel_area.contiguous().float(),
        "camera_indices": camera_indices[:, None].contiguous(),
    }


def make_targets(num_rays: int, seed: int = 3) -> Dict[str, Tensor]:
    g = torch.Generator().manual_seed(seed)
    image = torch.rand((num_rays, 3), generator=g)
    mask = (torch.rand((num_rays, 1), generator=g) < 0.1).float()
    return {"image": image, "fruit_mask": mask}


def make_jitter(num_rays: int, num_levels: int = 3, seed: int = 2) -> Tensor:
    """single-jitter random numbers for (piecewise sampler, pdf level 1, pdf level 2): [levels, R, 1]."""
    g = torch.Generator().manual_seed(seed)
    return torch.rand((num_levels, num_rays, 1), generator=g)


class JitterFeed:
    """Callable replacement for ``torch.rand`` inside the samplers: hands out pre-drawn jitter in call order."""

    def __init__(self, jitter: Tensor):
        self.jitter = jitter
        self.i = 0

    def reset(self) -> None:
        self.i = 0

    def __call__(self, shape, dtype=None, device=None):
        t = self.jitter[self.i % self.jitter.shape[0]]
        self.i += 1
        assert tuple(t.shape) == tuple(shape), f"jitter shape {tuple(t.shape)} != requested {tuple(shape)}"
        return t.to(device=device, dtype=dtype or torch.float32)


def randomize_state(state: Dict[str, Tensor], seed: int = 0, table_scale: float = 0.5, sem_bias: Optional[float] = None) -> Dict[str, Tensor]:
    """"Trained-like" weights: hash tables ~ U(-1,1)*table_scale (nerfstudio's init scale 1e-3 makes all densities ~1 and
    the proposal resampling trivial), everything else as initialised.  Done on the host so that oracle and product load
    the very same tensors."""
    g = torch.Generator().manual_seed(seed + 77)
    out = {}
    shared = {}  # aliases (mlp_base.0.hash_table / mlp_base_grid.hash_table) must receive the same tensor
    for k, v in state.items():
        if k.endswith("hash_table"):
            key = (v.data_ptr(), tuple(v.shape))
            if key not in shared:
                shared[key] = (torch.rand(v.shape, generator=g) * 2 - 1) * table_scale
            out[k] = shared[key]
        elif sem_bias is not None and k.endswith("field_head_semantics.net.bias"):
            out[k] = torch.full_like(v, sem_bias)
        else:
            out[k] = v.clone()
    return out
