"""Export / projection render loops (callers of the hot path), sharded by ray range across ranks with no data-path
collective.

* :func:`generate_point_cloud` -- the render loop of ``export/exporter_utils_nerfacto.py:126-183`` (``ns-export
  pointcloud`` as patched by ``debug/exporter_nerfacto.py``): render a batch of training rays, back-project the median
  depth, keep rays whose semantic label is "fruit" (``sigmoid(sem) > 0.9``) and that fall inside the crop OBB, until
  ``num_points`` are collected.  The open3d outlier removal / normal estimation that follows in the reference is CPU
  geometry-library work and out of scope.
* :func:`render_cluster_projection` -- the per-(camera, cluster AABB) step of ``FruitModel.get_outputs_for_projections``
  (``fruit_nerf.py:283-315``): un-occluded semantic render of the rays hitting the box, and the opacity accumulated in
  front of the box (occluded where >= 0.5).
* :func:`write_ply` -- ``semantics_pc.ply`` in the layout ``segmentation/segmenter.py:210`` reads (float xyz, uchar rgb).
"""
from __future__ import annotations

import struct
from dataclasses import dataclass
from typing import Callable, Dict, Optional, Tuple

import numpy as np
import torch
from torch import Tensor

from .rays import RayBundle


def shard_range(n: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous, disjoint, covering split of ``range(n)`` -- how rays / cameras are divided across ranks."""
    return (n * rank) // world_size, (n * (rank + 1)) // world_size


@dataclass
class OrientedBox:
    """nerfstudio ``OrientedBox`` (R, T, S) with ``within`` as used at exporter_utils_nerfacto.py:170-176."""

    R: Tensor  # [3,3]
    T: Tensor  # [3]
    S: Tensor  # [3]

    @staticmethod
    def from_params(pos, rpy, scale) -> "OrientedBox":
        r, p, y = rpy
        cr, sr, cp, sp, cy, sy = np.cos(r), np.sin(r), np.cos(p), np.sin(p), np.cos(y), np.sin(y)
        rx = np.array([[1, 0, 0], [0, cr, -sr], [0, sr, cr]])
        ry = np.array([[cp, 0, sp], [0, 1, 0], [-sp, 0, cp]])
        rz = np.array([[cy, -sy, 0], [sy, cy, 0], [0, 0, 1]])
        return OrientedBox(torch.tensor(rz @ ry @ rx, dtype=torch.float32), torch.tensor(pos, dtype=torch.float32), torch.tensor(scale, dtype=torch.float32))

    def within(self, pts: Tensor) -> Tensor:
        R, T, S = self.R.to(pts), self.T.to(pts), self.S.to(pts)
        local = (pts - T) @ R  # == (R^T (p - T))^T
        return ((local > -S / 2) & (local < S / 2)).all(dim=-1)


def generate_point_cloud(model, next_rays: Callable[[int], RayBundle], num_points: int, crop_obb: Optional[OrientedBox] = None,
                         only_semantics: bool = True, rank: int = 0, world_size: int = 1, max_batches: int = 1 << 30) -> Dict[str, Tensor]:
    """Rank-local share of the export: collects ``num_points / world_size`` kept points (over-produces at most one batch
    and trims).  ``next_rays(i)`` returns the i-th ray batch of this rank (ranks draw disjoint ray streams)."""
    lo, hi = shard_range(num_points, rank, world_size)
    want = hi - lo
    points, rgbs, dirs = [], [], []
    have = 0
    rays_rendered = 0
    i = 0
    while have < want and i < max_batches:
        with torch.no_grad():
            ray_bundle = next_rays(i)
            outputs = model(ray_bundle)
        rays_rendered += len(ray_bundle)
        point = ray_bundle.origins + ray_bundle.directions * outputs["depth"]
        mask = outputs["semantics_colormap"][:, 0] > 0 if only_semantics else torch.ones_like(outputs["depth"][:, 0], dtype=torch.bool)
        if crop_obb is not None:
            mask = mask & crop_obb.within(point)
        points.append(point[mask])
        rgbs.append(outputs["rgb"][mask])
        dirs.append(ray_bundle.directions[mask])
        have += int(mask.sum().item())
        i += 1
    cat = lambda xs, c: torch.cat(xs, dim=0)[:want] if xs else torch.empty((0, c))  # noqa: E731
    return {"points": cat(points, 3), "rgbs": cat(rgbs, 3), "view_directions": cat(dirs, 3), "rays_rendered": rays_rendered}


def write_ply(path: str, points: Tensor, rgbs: Tensor) -> None:
    """Binary little-endian PLY: float x y z, uchar red green blue (debug/exporter_nerfacto.py:140-146)."""
    pts = points.detach().cpu().numpy().astype("<f4")
    col = (rgbs.detach().cpu().clamp(0, 1).numpy() * 255).astype(np.uint8)
    n = pts.shape[0]
    header = ("ply\nformat binary_little_endian 1.0\n" f"element vertex {n}\n" "property float x\nproperty float y\nproperty float z\n"
              "property uchar red\nproperty uchar green\nproperty uchar blue\nend_header\n")
    rec = np.empty(n, dtype=[("p", "<f4", 3), ("c", "u1", 3)])
    rec["p"], rec["c"] = pts, col
    with open(path, "wb") as f:
        f.write(header.encode("ascii"))
        f.write(rec.tobytes())


def aabb_near_far(origins: Tensor, directions: Tensor, aabb: Tensor, invalid: float = 1e10) -> Tuple[Tensor, Tensor]:
    """Slab test of rays against an AABB [2,3] (what ``cam.generate_rays(aabb_box=...)`` sets, fruit_nerf.py:283);
    misses get nears = fars = ``invalid``."""
    inv = 1.0 / directions
    t0 = (aabb[0].to(origins) - origins) * inv
    t1 = (aabb[1].to(origins) - origins) * inv
    tmin = torch.minimum(t0, t1).amax(dim=-1, keepdim=True)
    tmax = torch.maximum(t0, t1).amin(dim=-1, keepdim=True)
    tmin = tmin.clamp_min(0.0)
    miss = tmax < tmin
    return torch.where(miss, torch.full_like(tmin, invalid), tmin), torch.where(miss, torch.full_like(tmax, invalid), tmax)


def render_cluster_projection(model, rays: RayBundle, aabb: Tensor) -> Dict[str, Tensor]:
    """One (camera, cluster) step of fruit_nerf.py:283-315 on a flat ray bundle [N]: returns per-ray ``valid``,
    ``semantics`` (un-occluded render between the box's near/far) and ``front_opacity`` (sum of weights on [0, near])."""
    dev = model.device
    o, d = rays.origins.to(dev), rays.directions.to(dev)
    nears, fars = aabb_near_far(o, d, aabb.to(dev))
    valid = nears[:, 0] < 1e10
    n = o.shape[0]
    sem = torch.zeros((n, 3), device=dev)
    front = torch.zeros((n,), device=dev)
    if int(valid.sum()) >= 10:  # fruit_nerf.py:293: fewer than 10 valid rays -> black images
        sub = RayBundle(origins=o[valid], directions=d[valid], pixel_area=None if rays.pixel_area is None else rays.pixel_area.to(dev)[valid],
                        camera_indices=None if rays.camera_indices is None else rays.camera_indices.to(dev)[valid],
                        nears=nears[valid], fars=fars[valid])
        with torch.no_grad():
            out = model.get_outputs_for_camera_jagged_ray_bundle(sub)
            sem[valid] = out["semantics"].to(dev).expand(-1, 3)
            sub.fars = sub.nears
            sub.nears = torch.zeros_like(sub.nears)
            front[valid] = model.get_density_for_camera_ray_bundle(sub).to(dev)
    visible = sem.clone()
    visible[front >= 0.5] = 0.0
    return {"valid": valid, "semantics": sem, "front_opacity": front, "visible": visible}
