"""Ray data carriers with nerfstudio's field names (``nerfstudio/cameras/rays.py``: RayBundle, RaySamples, Frustums).

The reference constructs these at ``components/ray_generators.py:59-64``, ``components/ray_samplers.py:96-102`` and
``scripts/semantic_projection.py:79-85`` and reads them throughout ``fruit_field.py`` / ``fruit_nerf.py``.  nerfstudio
itself is not a dependency of this package; objects of nerfstudio's own classes are accepted wherever these are
(only attributes are read).  All arithmetic on samples happens in the CUDA kernels; these classes only carry
tensors (mostly views into the bin-edge arrays the sampler kernels write).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Callable, Dict, Optional, Tuple

import torch
from torch import Tensor


@dataclass
class Frustums:
    origins: Tensor  # [..., 3]
    directions: Tensor  # [..., 3]
    starts: Tensor  # [..., 1]
    ends: Tensor  # [..., 1]
    pixel_area: Optional[Tensor]  # [..., 1]
    offsets: Optional[Tensor] = None

    @property
    def shape(self) -> torch.Size:
        return self.starts.shape[:-1]

    def get_positions(self) -> Tensor:
        """Sample centres (world frame).  Tiny elementwise expression kept for API parity (BayesRays,
        ``bayesrays/uncertainty.py:109``); the kernels recompute positions from rays + bin edges themselves."""
        pos = self.origins + self.directions * (self.starts + self.ends) / 2
        if self.offsets is not None:
            pos = pos + self.offsets
        return pos

    def set_offsets(self, offsets: Tensor) -> None:
        self.offsets = offsets


@dataclass
class RaySamples:
    frustums: Frustums
    camera_indices: Optional[Tensor] = None  # [..., 1]
    deltas: Optional[Tensor] = None  # [..., 1]
    spacing_starts: Optional[Tensor] = None
    spacing_ends: Optional[Tensor] = None
    spacing_to_euclidean_fn: Optional[Callable] = None
    metadata: Optional[Dict[str, Tensor]] = None
    times: Optional[Tensor] = None

    @property
    def shape(self) -> torch.Size:
        return self.frustums.shape

    def get_weights(self, densities: Tensor) -> Tensor:
        """``RaySamples.get_weights`` (called at fruit_nerf.py:556,508,442,341) on the compositing kernel."""
        from .ops import ray_weights

        return ray_weights(densities, self.frustums.starts, self.frustums.ends)


@dataclass
class RayBundle:
    origins: Tensor  # [R, 3]
    directions: Tensor  # [R, 3]
    pixel_area: Optional[Tensor] = None  # [R, 1]
    camera_indices: Optional[Tensor] = None  # [R, 1] int
    nears: Optional[Tensor] = None  # [R, 1]
    fars: Optional[Tensor] = None  # [R, 1]
    metadata: Dict[str, Tensor] = field(default_factory=dict)
    times: Optional[Tensor] = None

    _TENSORS = ("origins", "directions", "pixel_area", "camera_indices", "nears", "fars", "times")

    def __len__(self) -> int:
        n = 1
        for d in self.origins.shape[:-1]:
            n *= int(d)
        return n

    @property
    def shape(self) -> torch.Size:
        return self.origins.shape[:-1]

    def _map(self, fn) -> "RayBundle":
        kw = {k: (None if getattr(self, k) is None else fn(getattr(self, k))) for k in self._TENSORS}
        kw["metadata"] = {k: fn(v) for k, v in self.metadata.items()}
        return RayBundle(**kw)

    def flatten(self) -> "RayBundle":
        return self._map(lambda t: t.reshape(-1, t.shape[-1]))

    def get_row_major_sliced_ray_bundle(self, start_idx: int, end_idx: int) -> "RayBundle":
        return self.flatten()._map(lambda t: t[start_idx:end_idx])

    def __getitem__(self, idx) -> "RayBundle":
        return self._map(lambda t: t[idx])

    def to(self, device, non_blocking: bool = False) -> "RayBundle":
        return self._map(lambda t: t.to(device, non_blocking=non_blocking))

    def get_ray_samples(
        self,
        bin_starts: Tensor,
        bin_ends: Tensor,
        spacing_starts: Optional[Tensor] = None,
        spacing_ends: Optional[Tensor] = None,
        spacing_to_euclidean_fn: Optional[Callable] = None,
    ) -> RaySamples:
        S = bin_starts.shape[-2]
        lead = tuple(bin_starts.shape[:-2])
        ex = lambda t, c: None if t is None else t[..., None, :].expand(*lead, S, c)  # noqa: E731
        frustums = Frustums(
            origins=ex(self.origins, 3),
            directions=ex(self.directions, 3),
            starts=bin_starts,
            ends=bin_ends,
            pixel_area=ex(self.pixel_area, 1),
        )
        return RaySamples(
            frustums=frustums,
            camera_indices=ex(self.camera_indices, 1),
            deltas=bin_ends - bin_starts,
            spacing_starts=spacing_starts,
            spacing_ends=spacing_ends,
            spacing_to_euclidean_fn=spacing_to_euclidean_fn,
            metadata=None,
            times=None,
        )


def samples_from_edges(ray_bundle: RayBundle, euclid_edges: Tensor, spacing_edges: Tensor, spacing_to_euclidean_fn) -> RaySamples:
    """RaySamples whose starts/ends are views of the [R, S+1] bin-edge arrays written by the sampler kernels."""
    rs = ray_bundle.get_ray_samples(
        bin_starts=euclid_edges[..., :-1, None],
        bin_ends=euclid_edges[..., 1:, None],
        spacing_starts=spacing_edges[..., :-1, None],
        spacing_ends=spacing_edges[..., 1:, None],
        spacing_to_euclidean_fn=spacing_to_euclidean_fn,
    )
    rs.metadata = {"_euclid_edges": euclid_edges, "_spacing_edges": spacing_edges}
    return rs


def ray_layout(ray_samples) -> Tuple[Tensor, Tensor, Tensor, Tensor, Optional[Tensor], int, int, int]:
    """Flatten a RaySamples into what ``cnb_samples`` wants.

    Returns (origins [R,3], directions [R,3], starts, ends, camera_indices [R] int32 or None, R, S, row_stride)
    where element (r, s) of starts/ends lives at ``data_ptr + (r * row_stride + s) * 4``.  Ray-constant origins /
    directions (the expanded views ``get_ray_samples`` builds) are passed per ray; anything else (e.g.
    ``Field.density_fn(positions)``: one zero-length frustum per point) degrades to R = N rays of one sample.
    """
    fr = ray_samples.frustums
    shape = tuple(fr.starts.shape[:-1])
    o, d = fr.origins, fr.directions
    per_ray = len(shape) == 2 and o.dim() == 3 and o.stride(1) == 0 and d.stride(1) == 0
    cam = ray_samples.camera_indices
    if per_ray:
        R, S = shape
        origins = o[:, 0, :].float().contiguous()
        directions = d[:, 0, :].float().contiguous()
        starts, ends = fr.starts[..., 0], fr.ends[..., 0]
        ok = (
            starts.dtype == torch.float32
            and ends.dtype == torch.float32
            and (S == 1 or (starts.stride(1) == 1 and ends.stride(1) == 1))
            and starts.stride(0) == ends.stride(0)
            and starts.stride(0) >= S
        )
        if not ok:
            starts = starts.float().contiguous()
            ends = ends.float().contiguous()
        row_stride = starts.stride(0) if R > 1 else max(S, starts.stride(0))
        if cam is not None:
            cam = cam[:, 0, 0] if cam.dim() == 3 else cam.reshape(R, -1)[:, 0]
            cam = cam.to(torch.int32).contiguous()
        return origins, directions, starts, ends, cam, R, S, row_stride
    n = 1
    for v in shape:
        n *= int(v)
    origins = o.reshape(n, 3).float().contiguous()
    directions = d.expand(*shape, 3).reshape(n, 3).float().contiguous()
    starts = fr.starts.reshape(n).float().contiguous()
    ends = fr.ends.reshape(n).float().contiguous()
    if cam is not None:
        cam = cam.expand(*shape, 1).reshape(n).to(torch.int32).contiguous()
    return origins, directions, starts, ends, cam, n, 1, 1
