"""Restatement of the nerfstudio 1.1.3 torch primitives the FruitNeRF hot path imports.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``) -- parity unpinned: nerfstudio is not present in
``/root/reference`` nor installable here, so every function below restates the published nerfstudio 1.1.3
algorithm (module path given per function) as it is *used* by the cited reference call site
(paths relative to ``/root/reference/crop_nerf/fruit_nerf``).  SURVEY.md Appendix A is the spec.

Everything is plain torch, device/dtype agnostic (fp32 for parity, fp64 for gradient checks).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field, replace
from typing import Callable, Dict, List, Optional, Tuple

import numpy as np
import torch
from torch import Tensor, nn

# --------------------------------------------------------------------------------------------------------
# A.0 data carriers -- nerfstudio/cameras/rays.py (constructed at components/ray_generators.py:59-64,
#     components/ray_samplers.py:96-102, scripts/semantic_projection.py:79-85)
# --------------------------------------------------------------------------------------------------------


@dataclass
class Frustums:
    origins: Tensor  # [..., 3]
    directions: Tensor  # [..., 3]
    starts: Tensor  # [..., 1]
    ends: Tensor  # [..., 1]
    pixel_area: Tensor  # [..., 1]
    offsets: Optional[Tensor] = None

    @property
    def shape(self):
        return self.starts.shape[:-1]

    def get_positions(self) -> Tensor:
        # rays.py Frustums.get_positions: origins + directions * (starts + ends) / 2 (+ offsets)
        pos = self.origins + self.directions * (self.starts + self.ends) / 2
        if self.offsets is not None:
            pos = pos + self.offsets
        return pos

    def set_offsets(self, offsets: Tensor) -> None:
        self.offsets = offsets


@dataclass
class RaySamples:
    frustums: Frustums
    camera_indices: Optional[Tensor] = None  # [..., 1] int
    deltas: Optional[Tensor] = None  # [..., 1]
    spacing_starts: Optional[Tensor] = None  # [..., S, 1]
    spacing_ends: Optional[Tensor] = None
    spacing_to_euclidean_fn: Optional[Callable] = None
    metadata: Optional[Dict[str, Tensor]] = None

    @property
    def shape(self):
        return self.frustums.shape

    def get_weights(self, densities: Tensor) -> Tensor:
        """rays.py RaySamples.get_weights (called fruit_nerf.py:556,508,442,341)."""
        delta_density = self.deltas * densities
        alphas = 1 - torch.exp(-delta_density)
        transmittance = torch.cumsum(delta_density[..., :-1, :], dim=-2)
        transmittance = torch.cat(
            [torch.zeros((*transmittance.shape[:1], 1, 1), dtype=densities.dtype, device=densities.device), transmittance],
            dim=-2,
        )
        transmittance = torch.exp(-transmittance)
        weights = alphas * transmittance
        weights = torch.nan_to_num(weights)
        return weights


@dataclass
class RayBundle:
    origins: Tensor  # [R, 3]
    directions: Tensor  # [R, 3]
    pixel_area: Tensor  # [R, 1]
    camera_indices: Optional[Tensor] = None  # [R, 1] int
    nears: Optional[Tensor] = None  # [R, 1]
    fars: Optional[Tensor] = None  # [R, 1]
    metadata: Dict[str, Tensor] = field(default_factory=dict)

    def __len__(self) -> int:
        return int(np.prod(self.origins.shape[:-1]))

    def _map(self, fn):
        kw = {}
        for name in ("origins", "directions", "pixel_area", "camera_indices", "nears", "fars"):
            v = getattr(self, name)
            kw[name] = None if v is None else fn(v)
        kw["metadata"] = {k: fn(v) for k, v in self.metadata.items()}
        return RayBundle(**kw)

    def flatten(self) -> "RayBundle":
        return self._map(lambda t: t.reshape(-1, t.shape[-1]))

    def get_row_major_sliced_ray_bundle(self, start_idx: int, end_idx: int) -> "RayBundle":
        # rays.py: flatten then slice (used by the chunk loops fruit_nerf.py:333,360,391)
        return self.flatten()._map(lambda t: t[start_idx:end_idx])

    def __getitem__(self, idx) -> "RayBundle":
        return self._map(lambda t: t[idx])

    def to(self, device) -> "RayBundle":
        return self._map(lambda t: t.to(device))

    def get_ray_samples(
        self,
        bin_starts: Tensor,
        bin_ends: Tensor,
        spacing_starts: Optional[Tensor] = None,
        spacing_ends: Optional[Tensor] = None,
        spacing_to_euclidean_fn: Optional[Callable] = None,
    ) -> RaySamples:
        """rays.py RayBundle.get_ray_samples."""
        deltas = bin_ends - bin_starts
        camera_indices = self.camera_indices[..., None, :] if self.camera_indices is not None else None
        S = bin_starts.shape[-2]
        frustums = Frustums(
            origins=self.origins[..., None, :].expand(*bin_starts.shape[:-2], S, 3),
            directions=self.directions[..., None, :].expand(*bin_starts.shape[:-2], S, 3),
            starts=bin_starts,
            ends=bin_ends,
            pixel_area=self.pixel_area[..., None, :].expand(*bin_starts.shape[:-2], S, 1),
        )
        if camera_indices is not None:
            camera_indices = camera_indices.expand(*bin_starts.shape[:-2], S, 1)
        # TensorDataclass.__post_init__ broadcasts every field to the common batch shape (eval-mode bins are [1,S+1])
        if spacing_starts is not None:
            spacing_starts = spacing_starts.expand(bin_starts.shape)
        if spacing_ends is not None:
            spacing_ends = spacing_ends.expand(bin_ends.shape)
        return RaySamples(
            frustums=frustums,
            camera_indices=camera_indices,
            deltas=deltas,
            spacing_starts=spacing_starts,
            spacing_ends=spacing_ends,
            spacing_to_euclidean_fn=spacing_to_euclidean_fn,
            metadata=None,
        )


# --------------------------------------------------------------------------------------------------------
# A.1 HashEncoding -- nerfstudio/field_components/encodings.py (ctor call fruit_field.py:125-132,
#     proposal grids via HashMLPDensityField fruit_nerf.py:124-141)
# --------------------------------------------------------------------------------------------------------

HASH_PRIMES = (1, 2654435761, 805459861)


def hash_scalings(num_levels: int, min_res: int, max_res: int) -> Tensor:
    """encodings.py HashEncoding.__init__: float64 growth factor, float32 pow + floor.

    ``growth ** levels`` is python-float ** int64-tensor, which torch evaluates in float32; the field's top
    level therefore comes out as 2047, not 2048.
    """
    levels = torch.arange(num_levels)
    growth = np.exp((np.log(max_res) - np.log(min_res)) / (num_levels - 1)) if num_levels > 1 else 1.0
    return torch.floor(min_res * growth**levels)


class HashEncoding(nn.Module):
    def __init__(
        self,
        num_levels: int = 16,
        min_res: int = 16,
        max_res: int = 1024,
        log2_hashmap_size: int = 19,
        features_per_level: int = 2,
        hash_init_scale: float = 0.001,
    ) -> None:
        super().__init__()
        self.num_levels = num_levels
        self.min_res = min_res
        self.features_per_level = features_per_level
        self.hash_init_scale = hash_init_scale
        self.log2_hashmap_size = log2_hashmap_size
        self.hash_table_size = 2**log2_hashmap_size
        levels = torch.arange(num_levels)
        self.register_buffer("scalings", hash_scalings(num_levels, min_res, max_res))
        self.hash_offset = levels * self.hash_table_size
        table = torch.rand(size=(self.hash_table_size * num_levels, features_per_level)) * 2 - 1
        table *= hash_init_scale
        self.hash_table = nn.Parameter(table)

    def get_out_dim(self) -> int:
        return self.num_levels * self.features_per_level

    def hash_fn(self, in_tensor: Tensor) -> Tensor:
        """encodings.py HashEncoding.hash_fn: int32 coords * int64 primes -> int64, xor, mod, + level offset."""
        in_tensor = in_tensor * torch.tensor(HASH_PRIMES, device=in_tensor.device)
        x = torch.bitwise_xor(in_tensor[..., 0], in_tensor[..., 1])
        x = torch.bitwise_xor(x, in_tensor[..., 2])
        x %= self.hash_table_size
        x += self.hash_offset.to(x.device)
        return x

    def corner_indices(self, in_tensor: Tensor) -> Tuple[Tensor, Tensor]:
        """Returns ([..., L, 8] int64 table rows in h0..h7 order, offsets [..., L, 3])."""
        in_tensor = in_tensor[..., None, :]
        scaled = in_tensor * self.scalings.view(-1, 1).to(in_tensor)
        scaled_c = torch.ceil(scaled).type(torch.int32)
        scaled_f = torch.floor(scaled).type(torch.int32)
        offset = scaled - scaled_f
        c, f = scaled_c, scaled_f
        cat = lambda a, b, d: torch.cat([a[..., 0:1], b[..., 1:2], d[..., 2:3]], dim=-1)  # noqa: E731
        hashed = [
            self.hash_fn(c),  # 0: c c c
            self.hash_fn(cat(c, f, c)),  # 1: c f c
            self.hash_fn(cat(f, f, c)),  # 2: f f c
            self.hash_fn(cat(f, c, c)),  # 3: f c c
            self.hash_fn(cat(c, c, f)),  # 4: c c f
            self.hash_fn(cat(c, f, f)),  # 5: c f f
            self.hash_fn(f),  # 6: f f f
            self.hash_fn(cat(f, c, f)),  # 7: f c f
        ]
        return torch.stack(hashed, dim=-1), offset

    def forward(self, in_tensor: Tensor) -> Tensor:
        """encodings.py HashEncoding.pytorch_fwd."""
        assert in_tensor.shape[-1] == 3
        hashed, offset = self.corner_indices(in_tensor)
        f = [self.hash_table[hashed[..., i]] for i in range(8)]
        ox, oy, oz = offset[..., 0:1], offset[..., 1:2], offset[..., 2:3]
        f_03 = f[0] * ox + f[3] * (1 - ox)
        f_12 = f[1] * ox + f[2] * (1 - ox)
        f_56 = f[5] * ox + f[6] * (1 - ox)
        f_47 = f[4] * ox + f[7] * (1 - ox)
        f0312 = f_03 * oy + f_12 * (1 - oy)
        f4756 = f_47 * oy + f_56 * (1 - oy)
        encoded_value = f0312 * oz + f4756 * (1 - oz)
        return torch.flatten(encoded_value, start_dim=-2, end_dim=-1)


# --------------------------------------------------------------------------------------------------------
# A.2 MLP -- nerfstudio/field_components/mlp.py (ctor calls fruit_field.py:133-141,146-154,159-167)
# --------------------------------------------------------------------------------------------------------


class MLP(nn.Module):
    def __init__(
        self,
        in_dim: int,
        num_layers: int,
        layer_width: int,
        out_dim: Optional[int] = None,
        activation: Optional[nn.Module] = nn.ReLU(),
        out_activation: Optional[nn.Module] = None,
    ) -> None:
        super().__init__()
        self.in_dim = in_dim
        self.out_dim = out_dim if out_dim is not None else layer_width
        self.num_layers = num_layers
        self.layer_width = layer_width
        self.activation = activation
        self.out_activation = out_activation
        layers = []
        if num_layers == 1:
            layers.append(nn.Linear(in_dim, self.out_dim))
        else:
            for i in range(num_layers - 1):
                layers.append(nn.Linear(in_dim if i == 0 else layer_width, layer_width))
            layers.append(nn.Linear(layer_width, self.out_dim))
        self.layers = nn.ModuleList(layers)

    def get_out_dim(self) -> int:
        return self.out_dim

    def forward(self, in_tensor: Tensor) -> Tensor:
        x = in_tensor
        for i, layer in enumerate(self.layers):
            x = layer(x)
            if self.activation is not None and i < len(self.layers) - 1:
                x = self.activation(x)
        if self.out_activation is not None:
            x = self.out_activation(x)
        return x


# --------------------------------------------------------------------------------------------------------
# A.4 misc field components
# --------------------------------------------------------------------------------------------------------


class _TruncExp(torch.autograd.Function):
    """nerfstudio/field_components/activations.py: exp forward (fp32), backward g*exp(clamp(x,-15,15))."""

    @staticmethod
    def forward(ctx, x):
        ctx.save_for_backward(x)
        return torch.exp(x)

    @staticmethod
    def backward(ctx, g):
        x = ctx.saved_tensors[0]
        return g * torch.exp(torch.clamp(x, min=-15, max=15))


trunc_exp = _TruncExp.apply


class SceneContraction(nn.Module):
    """nerfstudio/field_components/spatial_distortions.py, order=inf (fruit_nerf.py:95)."""

    def __init__(self, order=float("inf")) -> None:
        super().__init__()
        self.order = order

    def forward(self, positions: Tensor) -> Tensor:
        mag = torch.linalg.norm(positions, ord=self.order, dim=-1)[..., None]
        return torch.where(mag < 1, positions, (2 - (1 / mag)) * (positions / mag))


def get_normalized_positions(positions: Tensor, aabb: Tensor) -> Tensor:
    """nerfstudio/data/scene_box.py SceneBox.get_normalized_positions (fruit_field.py:176)."""
    aabb_lengths = aabb[1] - aabb[0]
    return (positions - aabb[0]) / aabb_lengths


def get_normalized_directions(directions: Tensor) -> Tensor:
    """nerfstudio/fields/base_field.py (imported as shift_directions_for_tcnn, fruit_field.py:39,209,244)."""
    return (directions + 1.0) / 2.0


def components_from_spherical_harmonics(degree: int, directions: Tensor) -> Tensor:
    """nerfstudio/utils/math.py; degree = levels-1 = 3 for SHEncoding(levels=4) (fruit_field.py:116-119)."""
    num_components = (degree + 1) ** 2
    components = torch.zeros((*directions.shape[:-1], num_components), device=directions.device, dtype=directions.dtype)
    assert 0 <= degree <= 4
    x = directions[..., 0]
    y = directions[..., 1]
    z = directions[..., 2]
    xx = x**2
    yy = y**2
    zz = z**2
    components[..., 0] = 0.28209479177387814
    if degree > 0:
        components[..., 1] = 0.4886025119029199 * y
        components[..., 2] = 0.4886025119029199 * z
        components[..., 3] = 0.4886025119029199 * x
    if degree > 1:
        components[..., 4] = 1.0925484305920792 * x * y
        components[..., 5] = 1.0925484305920792 * y * z
        components[..., 6] = 0.9461746957575601 * zz - 0.31539156525251999
        components[..., 7] = 1.0925484305920792 * x * z
        components[..., 8] = 0.5462742152960396 * (xx - yy)
    if degree > 2:
        components[..., 9] = 0.5900435899266435 * y * (3 * xx - yy)
        components[..., 10] = 2.890611442640554 * x * y * z
        components[..., 11] = 0.4570457994644658 * y * (5 * zz - 1)
        components[..., 12] = 0.3731763325901154 * z * (5 * zz - 3)
        components[..., 13] = 0.4570457994644658 * x * (5 * zz - 1)
        components[..., 14] = 1.445305721320277 * z * (xx - yy)
        components[..., 15] = 0.5900435899266435 * x * (xx - 3 * yy)
    return components


class SHEncoding(nn.Module):
    """encodings.py SHEncoding.pytorch_fwd: applied directly to its input under no_grad (App. A.4 [verify])."""

    def __init__(self, levels: int = 4) -> None:
        super().__init__()
        self.levels = levels

    def get_out_dim(self) -> int:
        return self.levels**2

    @torch.no_grad()
    def forward(self, in_tensor: Tensor) -> Tensor:
        return components_from_spherical_harmonics(degree=self.levels - 1, directions=in_tensor)


class Embedding(nn.Module):
    """nerfstudio/field_components/embedding.py (fruit_field.py:106,220,257)."""

    def __init__(self, in_dim: int, out_dim: int) -> None:
        super().__init__()
        self.embedding = nn.Embedding(in_dim, out_dim)

    def mean(self, dim=0):
        return self.embedding.weight.mean(dim)

    def forward(self, in_tensor: Tensor) -> Tensor:
        return self.embedding(in_tensor)


# --------------------------------------------------------------------------------------------------------
# A.3 HashMLPDensityField -- nerfstudio/fields/density_fields.py (built fruit_nerf.py:118-142)
# --------------------------------------------------------------------------------------------------------


class HashMLPDensityField(nn.Module):
    def __init__(
        self,
        aabb: Tensor,
        num_layers: int = 2,
        hidden_dim: int = 64,
        spatial_distortion: Optional[nn.Module] = None,
        use_linear: bool = False,
        num_levels: int = 8,
        max_res: int = 1024,
        base_res: int = 16,
        log2_hashmap_size: int = 18,
        features_per_level: int = 2,
        average_init_density: float = 1.0,
    ) -> None:
        super().__init__()
        self.register_buffer("aabb", aabb)
        self.spatial_distortion = spatial_distortion
        self.use_linear = use_linear
        self.average_init_density = average_init_density
        self.register_buffer("max_res", torch.tensor(max_res))
        self.register_buffer("num_levels", torch.tensor(num_levels))
        self.register_buffer("log2_hashmap_size", torch.tensor(log2_hashmap_size))
        self.encoding = HashEncoding(
            num_levels=num_levels,
            min_res=base_res,
            max_res=max_res,
            log2_hashmap_size=log2_hashmap_size,
            features_per_level=features_per_level,
        )
        if not use_linear:
            network = MLP(
                in_dim=self.encoding.get_out_dim(),
                num_layers=num_layers,
                layer_width=hidden_dim,
                out_dim=1,
                activation=nn.ReLU(),
                out_activation=None,
            )
            self.mlp_base = nn.Sequential(self.encoding, network)
        else:
            self.linear = nn.Linear(self.encoding.get_out_dim(), 1)

    def normalised_positions(self, positions: Tensor) -> Tuple[Tensor, Tensor]:
        """Contraction / AABB normalisation + selector mask of density_fields.py get_density.  The reference tree restates the same
        steps itself ("according to density_feild.py in nerfstudio", bayesrays/utils.py:6-16); tests/test_reference_pin_cpu.py
        executes that function and requires this one to agree bit for bit."""
        if self.spatial_distortion is not None:
            positions = self.spatial_distortion(positions)
            positions = (positions + 2.0) / 4.0
        else:
            positions = get_normalized_positions(positions, self.aabb)
        selector = ((positions > 0.0) & (positions < 1.0)).all(dim=-1)
        positions = positions * selector[..., None]
        return positions, selector

    def get_density(self, ray_samples: RaySamples) -> Tuple[Tensor, None]:
        positions, selector = self.normalised_positions(ray_samples.frustums.get_positions())
        positions_flat = positions.view(-1, 3)
        if not self.use_linear:
            dba = self.mlp_base(positions_flat).view(*ray_samples.frustums.shape, -1).to(positions)
        else:
            x = self.encoding(positions_flat).to(positions)
            dba = self.linear(x).view(*ray_samples.frustums.shape, -1)
        density = self.average_init_density * trunc_exp(dba)
        density = density * selector[..., None]
        return density, None

    def density_fn(self, positions: Tensor) -> Tensor:
        """nerfstudio/fields/base_field.py Field.density_fn: wrap positions in zero-length frustums."""
        ray_samples = RaySamples(
            frustums=Frustums(
                origins=positions,
                directions=torch.ones_like(positions),
                starts=torch.zeros_like(positions[..., :1]),
                ends=torch.zeros_like(positions[..., :1]),
                pixel_area=torch.ones_like(positions[..., :1]),
            )
        )
        density, _ = self.get_density(ray_samples)
        return density


# --------------------------------------------------------------------------------------------------------
# A.6 samplers -- nerfstudio/model_components/ray_samplers.py
# --------------------------------------------------------------------------------------------------------


class SpacedSampler(nn.Module):
    """ray_samplers.py SpacedSampler; the reference's own copy of this routine is
    components/ray_samplers.py:54-104 (UniformSamplerWithNoise.generate_ray_samples)."""

    def __init__(self, spacing_fn, spacing_fn_inv, num_samples=None, train_stratified=True, single_jitter=False):
        super().__init__()
        self.num_samples = num_samples
        self.train_stratified = train_stratified
        self.single_jitter = single_jitter
        self.spacing_fn = spacing_fn
        self.spacing_fn_inv = spacing_fn_inv
        self.rand_fn = torch.rand  # injectable: tests feed the same jitter to oracle and kernels

    def forward(self, ray_bundle: RayBundle, num_samples: Optional[int] = None) -> RaySamples:
        assert ray_bundle.nears is not None and ray_bundle.fars is not None
        num_samples = num_samples or self.num_samples
        num_rays = ray_bundle.origins.shape[0]
        dt = ray_bundle.origins.dtype
        bins = torch.linspace(0.0, 1.0, num_samples + 1, dtype=dt).to(ray_bundle.origins.device)[None, ...]
        if self.train_stratified and self.training:
            if self.single_jitter:
                t_rand = self.rand_fn((num_rays, 1), dtype=bins.dtype, device=bins.device)
            else:
                t_rand = self.rand_fn((num_rays, num_samples + 1), dtype=bins.dtype, device=bins.device)
            bin_centers = (bins[..., 1:] + bins[..., :-1]) / 2.0
            bin_upper = torch.cat([bin_centers, bins[..., -1:]], -1)
            bin_lower = torch.cat([bins[..., :1], bin_centers], -1)
            bins = bin_lower + (bin_upper - bin_lower) * t_rand
        s_near, s_far = (self.spacing_fn(x) for x in (ray_bundle.nears, ray_bundle.fars))

        def spacing_to_euclidean_fn(x):
            return self.spacing_fn_inv(x * s_far + (1 - x) * s_near)

        euclidean_bins = spacing_to_euclidean_fn(bins)
        return ray_bundle.get_ray_samples(
            bin_starts=euclidean_bins[..., :-1, None],
            bin_ends=euclidean_bins[..., 1:, None],
            spacing_starts=bins[..., :-1, None],
            spacing_ends=bins[..., 1:, None],
            spacing_to_euclidean_fn=spacing_to_euclidean_fn,
        )


class UniformSampler(SpacedSampler):
    def __init__(self, num_samples=None, train_stratified=True, single_jitter=False):
        super().__init__(lambda x: x, lambda x: x, num_samples, train_stratified, single_jitter)


class UniformSamplerWithNoise(UniformSampler):
    """components/ray_samplers.py:31-104 -- identity spacing, otherwise SpacedSampler."""


class UniformLinDispPiecewiseSampler(SpacedSampler):
    def __init__(self, num_samples=None, train_stratified=True, single_jitter=False):
        super().__init__(
            spacing_fn=lambda x: torch.where(x < 1, x / 2, 1 - 1 / (2 * x)),
            spacing_fn_inv=lambda x: torch.where(x < 0.5, 2 * x, 1 / (2 - 2 * x)),
            num_samples=num_samples,
            train_stratified=train_stratified,
            single_jitter=single_jitter,
        )


class PDFSampler(nn.Module):
    def __init__(self, num_samples=None, train_stratified=True, single_jitter=False, include_original=True, histogram_padding=0.01):
        super().__init__()
        self.num_samples = num_samples
        self.train_stratified = train_stratified
        self.single_jitter = single_jitter
        self.include_original = include_original
        self.histogram_padding = histogram_padding
        self.rand_fn = torch.rand
        self.last_inds: Optional[Tensor] = None  # exposed for bit-exact bin-index parity tests

    def forward(self, ray_bundle: RayBundle, ray_samples: RaySamples, weights: Tensor, num_samples=None, eps: float = 1e-5) -> RaySamples:
        num_samples = num_samples or self.num_samples
        num_bins = num_samples + 1
        weights = weights[..., 0] + self.histogram_padding
        weights_sum = torch.sum(weights, dim=-1, keepdim=True)
        padding = torch.relu(eps - weights_sum)
        weights = weights + padding / weights.shape[-1]
        weights_sum = weights_sum + padding
        pdf = weights / weights_sum
        cdf = torch.min(torch.ones_like(pdf), torch.cumsum(pdf, dim=-1))
        cdf = torch.cat([torch.zeros_like(cdf[..., :1]), cdf], dim=-1)
        if self.train_stratified and self.training:
            u = torch.linspace(0.0, 1.0 - (1.0 / num_bins), steps=num_bins, device=cdf.device, dtype=cdf.dtype)
            u = u.expand(size=(*cdf.shape[:-1], num_bins))
            if self.single_jitter:
                rand = self.rand_fn((*cdf.shape[:-1], 1), device=cdf.device, dtype=cdf.dtype) / num_bins
            else:
                rand = self.rand_fn((*cdf.shape[:-1], num_samples + 1), device=cdf.device, dtype=cdf.dtype) / num_bins
            u = u + rand
        else:
            u = torch.linspace(0.0, 1.0 - (1.0 / num_bins), steps=num_bins, device=cdf.device, dtype=cdf.dtype)
            u = u + 1.0 / (2 * num_bins)
            u = u.expand(size=(*cdf.shape[:-1], num_bins))
        u = u.contiguous()
        existing_bins = torch.cat([ray_samples.spacing_starts[..., 0], ray_samples.spacing_ends[..., -1:, 0]], dim=-1)
        inds = torch.searchsorted(cdf, u, side="right")
        self.last_inds = inds
        below = torch.clamp(inds - 1, 0, existing_bins.shape[-1] - 1)
        above = torch.clamp(inds, 0, existing_bins.shape[-1] - 1)
        cdf_g0 = torch.gather(cdf, -1, below)
        bins_g0 = torch.gather(existing_bins, -1, below)
        cdf_g1 = torch.gather(cdf, -1, above)
        bins_g1 = torch.gather(existing_bins, -1, above)
        t = torch.clip(torch.nan_to_num((u - cdf_g0) / (cdf_g1 - cdf_g0), 0), 0, 1)
        bins = bins_g0 + t * (bins_g1 - bins_g0)
        if self.include_original:
            bins, _ = torch.sort(torch.cat([existing_bins, bins], -1), -1)
        bins = bins.detach()
        euclidean_bins = ray_samples.spacing_to_euclidean_fn(bins)
        return ray_bundle.get_ray_samples(
            bin_starts=euclidean_bins[..., :-1, None],
            bin_ends=euclidean_bins[..., 1:, None],
            spacing_starts=bins[..., :-1, None],
            spacing_ends=bins[..., 1:, None],
            spacing_to_euclidean_fn=ray_samples.spacing_to_euclidean_fn,
        )


class ProposalNetworkSampler(nn.Module):
    """ray_samplers.py ProposalNetworkSampler (built fruit_nerf.py:157-164; called :549,501,429,337)."""

    def __init__(
        self,
        num_proposal_samples_per_ray: Tuple[int, ...] = (64,),
        num_nerf_samples_per_ray: int = 32,
        num_proposal_network_iterations: int = 2,
        single_jitter: bool = False,
        update_sched: Callable = lambda x: 1,
        initial_sampler: Optional[nn.Module] = None,
    ) -> None:
        super().__init__()
        self.num_proposal_samples_per_ray = num_proposal_samples_per_ray
        self.num_nerf_samples_per_ray = num_nerf_samples_per_ray
        self.num_proposal_network_iterations = num_proposal_network_iterations
        self.update_sched = update_sched
        if self.num_proposal_network_iterations < 1:
            raise ValueError("num_proposal_network_iterations must be >= 1")
        if initial_sampler is None:
            self.initial_sampler = UniformLinDispPiecewiseSampler(single_jitter=single_jitter)
        else:
            self.initial_sampler = initial_sampler
        self.pdf_sampler = PDFSampler(include_original=False, single_jitter=single_jitter)
        self._anneal = 1.0
        self._steps_since_update = 0
        self._step = 0

    def set_anneal(self, anneal: float) -> None:
        self._anneal = anneal

    def step_cb(self, step):
        self._step = step
        self._steps_since_update += 1

    def forward(self, ray_bundle: RayBundle, density_fns: List[Callable]):
        weights_list, ray_samples_list = [], []
        n = self.num_proposal_network_iterations
        weights = None
        ray_samples = None
        updated = self._steps_since_update > self.update_sched(self._step) or self._step < 10
        for i_level in range(n + 1):
            is_prop = i_level < n
            num_samples = self.num_proposal_samples_per_ray[i_level] if is_prop else self.num_nerf_samples_per_ray
            if i_level == 0:
                ray_samples = self.initial_sampler(ray_bundle, num_samples=num_samples)
            else:
                annealed_weights = torch.pow(weights, self._anneal)
                ray_samples = self.pdf_sampler(ray_bundle, ray_samples, annealed_weights, num_samples=num_samples)
            if is_prop:
                if updated:
                    density = density_fns[i_level](ray_samples.frustums.get_positions())
                else:
                    with torch.no_grad():
                        density = density_fns[i_level](ray_samples.frustums.get_positions())
                weights = ray_samples.get_weights(density)
                weights_list.append(weights)
                ray_samples_list.append(ray_samples)
        if updated:
            self._steps_since_update = 0
        return ray_samples, weights_list, ray_samples_list


class NearFarCollider(nn.Module):
    """nerfstudio/model_components/scene_colliders.py (fruit_nerf.py:167): pass-through if nears/fars set
    (this is what lets get_outputs_for_projections inject AABB near/far, fruit_nerf.py:283,307-308);
    eval mode resets the near plane to 0 (reset_near_plane=True default) [verify]."""

    def __init__(self, near_plane: float, far_plane: float, reset_near_plane: bool = True):
        super().__init__()
        self.near_plane = near_plane
        self.far_plane = far_plane
        self.reset_near_plane = reset_near_plane

    def forward(self, ray_bundle: RayBundle) -> RayBundle:
        if ray_bundle.nears is not None and ray_bundle.fars is not None:
            return ray_bundle
        ones = torch.ones_like(ray_bundle.origins[..., 0:1])
        near_plane = self.near_plane if (self.training or not self.reset_near_plane) else 0
        return replace(ray_bundle, nears=ones * near_plane, fars=ones * self.far_plane)


# --------------------------------------------------------------------------------------------------------
# A.7 renderers -- nerfstudio/model_components/renderers.py (fruit_nerf.py:170-174,560-591)
# --------------------------------------------------------------------------------------------------------

BACKGROUND_COLOR_OVERRIDE: Optional[Tensor] = None


class background_color_override_context:  # noqa: N801  (same name as nerfstudio; semantic_projection.py:51,169)
    def __init__(self, color: Tensor):
        self.color = color

    def __enter__(self):
        global BACKGROUND_COLOR_OVERRIDE
        self.old = BACKGROUND_COLOR_OVERRIDE
        BACKGROUND_COLOR_OVERRIDE = self.color
        return self

    def __exit__(self, *a):
        global BACKGROUND_COLOR_OVERRIDE
        BACKGROUND_COLOR_OVERRIDE = self.old


class RGBRenderer(nn.Module):
    def __init__(self, background_color="random"):
        super().__init__()
        self.background_color = background_color

    @classmethod
    def combine_rgb(cls, rgb: Tensor, weights: Tensor, background_color="random") -> Tensor:
        comp_rgb = torch.sum(weights * rgb, dim=-2)
        accumulated_weight = torch.sum(weights, dim=-2)
        if BACKGROUND_COLOR_OVERRIDE is not None:
            background_color = BACKGROUND_COLOR_OVERRIDE
        if isinstance(background_color, str) and background_color == "random":
            return comp_rgb
        if isinstance(background_color, str) and background_color == "last_sample":
            background_color = rgb[..., -1, :]
        elif isinstance(background_color, str):
            background_color = {"black": torch.zeros(3), "white": torch.ones(3)}[background_color]
        background_color = background_color.to(comp_rgb).expand(comp_rgb.shape)
        return comp_rgb + background_color * (1.0 - accumulated_weight)

    def forward(self, rgb: Tensor, weights: Tensor) -> Tensor:
        if not self.training:
            rgb = torch.nan_to_num(rgb)
        rgb = self.combine_rgb(rgb, weights, background_color=self.background_color)
        if not self.training:
            rgb = torch.clamp(rgb, min=0.0, max=1.0)
        return rgb


class AccumulationRenderer(nn.Module):
    def forward(self, weights: Tensor) -> Tensor:
        return torch.sum(weights, dim=-2)


class DepthRenderer(nn.Module):
    def __init__(self, method: str = "median"):
        super().__init__()
        self.method = method
        self.last_median_index: Optional[Tensor] = None

    def forward(self, weights: Tensor, ray_samples: RaySamples) -> Tensor:
        if self.method == "median":
            steps = (ray_samples.frustums.starts + ray_samples.frustums.ends) / 2
            cumulative_weights = torch.cumsum(weights[..., 0], dim=-1)
            split = torch.ones((*weights.shape[:-2], 1), device=weights.device, dtype=weights.dtype) * 0.5
            median_index = torch.searchsorted(cumulative_weights, split, side="left")
            median_index = torch.clamp(median_index, 0, steps.shape[-2] - 1)
            self.last_median_index = median_index
            return torch.gather(steps[..., 0], dim=-1, index=median_index)
        if self.method == "expected":
            eps = 1e-10
            steps = (ray_samples.frustums.starts + ray_samples.frustums.ends) / 2
            depth = torch.sum(weights * steps, dim=-2) / (torch.sum(weights, -2) + eps)
            return torch.clip(depth, steps.min(), steps.max())
        raise NotImplementedError(self.method)


class SemanticRenderer(nn.Module):
    def forward(self, semantics: Tensor, weights: Tensor) -> Tensor:
        return torch.sum(weights * semantics, dim=-2)


# --------------------------------------------------------------------------------------------------------
# A.8 losses -- nerfstudio/model_components/losses.py (fruit_nerf.py:177-178,601-615,639-645)
# --------------------------------------------------------------------------------------------------------

EPSILON = 1e-7


def ray_samples_to_sdist(ray_samples: RaySamples) -> Tensor:
    starts = ray_samples.spacing_starts
    ends = ray_samples.spacing_ends
    return torch.cat([starts[..., 0], ends[..., -1:, 0]], dim=-1)


def outer(t0_starts, t0_ends, t1_starts, t1_ends, y1):
    cy1 = torch.cat([torch.zeros_like(y1[..., :1]), torch.cumsum(y1, dim=-1)], dim=-1)
    idx_lo = torch.searchsorted(t1_starts.contiguous(), t0_starts.contiguous(), side="right") - 1
    idx_lo = torch.clamp(idx_lo, min=0, max=y1.shape[-1] - 1)
    idx_hi = torch.searchsorted(t1_ends.contiguous(), t0_ends.contiguous(), side="right")
    idx_hi = torch.clamp(idx_hi, min=0, max=y1.shape[-1] - 1)
    cy1_lo = torch.take_along_dim(cy1[..., :-1], idx_lo, dim=-1)
    cy1_hi = torch.take_along_dim(cy1[..., 1:], idx_hi, dim=-1)
    return cy1_hi - cy1_lo


def lossfun_outer(t, w, t_env, w_env):
    w_outer = outer(t[..., :-1], t[..., 1:], t_env[..., :-1], t_env[..., 1:], w_env)
    return torch.clip(w - w_outer, min=0) ** 2 / (w + EPSILON)


def interlevel_loss(weights_list, ray_samples_list) -> Tensor:
    c = ray_samples_to_sdist(ray_samples_list[-1]).detach()
    w = weights_list[-1][..., 0].detach()
    loss_interlevel = 0.0
    for ray_samples, weights in zip(ray_samples_list[:-1], weights_list[:-1]):
        cp = ray_samples_to_sdist(ray_samples)
        wp = weights[..., 0]
        loss_interlevel = loss_interlevel + torch.mean(lossfun_outer(c, w, cp, wp))
    return loss_interlevel


def lossfun_distortion(t, w):
    ut = (t[..., 1:] + t[..., :-1]) / 2
    dut = torch.abs(ut[..., :, None] - ut[..., None, :])
    loss_inter = torch.sum(w * torch.sum(w[..., None, :] * dut, dim=-1), dim=-1)
    loss_intra = torch.sum(w**2 * (t[..., 1:] - t[..., :-1]), dim=-1) / 3
    return loss_inter + loss_intra


def distortion_loss(weights_list, ray_samples_list) -> Tensor:
    c = ray_samples_to_sdist(ray_samples_list[-1])
    w = weights_list[-1][..., 0]
    return torch.mean(lossfun_distortion(c, w))


# --------------------------------------------------------------------------------------------------------
# Field head -- nerfstudio/field_components/field_heads.py FieldHead (components/field_heads.py:29-40)
# --------------------------------------------------------------------------------------------------------


class FieldHead(nn.Module):
    def __init__(self, in_dim: int, out_dim: int, activation=None):
        super().__init__()
        self.net = nn.Linear(in_dim, out_dim)
        self.activation = activation

    def forward(self, in_tensor: Tensor) -> Tensor:
        out = self.net(in_tensor)
        if self.activation:
            out = self.activation(out)
        return out


def exponential_decay_lr(step: int, lr_init: float, lr_final: float, max_steps: int) -> float:
    """nerfstudio/engine/schedulers.py ExponentialDecayScheduler without warm-up (fruit_nerf_config.py:45-60)."""
    t = min(max(step / max_steps, 0.0), 1.0)
    return math.exp(math.log(lr_init) * (1 - t) + math.log(lr_final) * t)


# --------------------------------------------------------------------------------------------------------
# f1 ("next" row of SURVEY.md section 8): ray generation -- nerfstudio/cameras/cameras.py Cameras.generate_rays /
#     _generate_rays_from_coords (perspective, no distortion) and nerfstudio/utils/math.py intersect_aabb, as used by
#     cam.generate_rays(camera_indices=0, keep_shape=True, aabb_box=aabb) at fruit_nerf.py:283.  [verify]: restated from the
#     published nerfstudio 1.1.3 source, not line-checked (nerfstudio is absent here).
# --------------------------------------------------------------------------------------------------------


def generate_pinhole_rays(c2w: Tensor, fx: float, fy: float, cx: float, cy: float, coords_yx: Tensor):
    """coords_yx [..., 2] pixel coordinates INCLUDING the 0.5 pixel-centre offset -> (origins, directions, pixel_area)."""
    y, x = coords_yx[..., 0], coords_yx[..., 1]
    coord = torch.stack([(x - cx) / fx, -(y - cy) / fy], -1)
    coord_x_offset = torch.stack([(x - cx + 1) / fx, -(y - cy) / fy], -1)
    coord_y_offset = torch.stack([(x - cx) / fx, -(y - cy + 1) / fy], -1)
    coord_stack = torch.stack([coord, coord_x_offset, coord_y_offset], dim=0)
    directions_stack = torch.cat([coord_stack, -torch.ones_like(coord_stack[..., :1])], dim=-1)
    rotation = c2w[:3, :3]
    directions_stack = torch.sum(directions_stack[..., None, :] * rotation, dim=-1)
    norm = torch.linalg.norm(directions_stack, dim=-1, keepdim=True)
    directions_stack = directions_stack / norm
    directions = directions_stack[0]
    dx = torch.sqrt(torch.sum((directions - directions_stack[1]) ** 2, dim=-1))
    dy = torch.sqrt(torch.sum((directions - directions_stack[2]) ** 2, dim=-1))
    pixel_area = (dx * dy)[..., None]
    origins = c2w[:3, 3].expand(directions.shape)
    return origins, directions, pixel_area


def intersect_aabb(origins: Tensor, directions: Tensor, aabb: Tensor, max_bound: float = 1e10, invalid_value: float = 1e10):
    """aabb [6] = (min xyz, max xyz) -> (t_min, t_max), both `invalid_value` where the ray misses the box."""
    tx_min = (aabb[:3] - origins) / directions
    tx_max = (aabb[3:] - origins) / directions
    t_min = torch.stack((tx_min, tx_max)).amin(dim=0)
    t_max = torch.stack((tx_min, tx_max)).amax(dim=0)
    t_min = t_min.amax(dim=-1)
    t_max = t_max.amin(dim=-1)
    t_min = torch.clamp(t_min, min=0, max=max_bound)
    t_max = torch.clamp(t_max, min=0, max=max_bound)
    cond = t_max <= t_min
    t_min = torch.where(cond, torch.full_like(t_min, invalid_value), t_min)
    t_max = torch.where(cond, torch.full_like(t_max, invalid_value), t_max)
    return t_min, t_max


# --------------------------------------------------------------------------------------------------------
# A.10 training batch -- FruitDataManager.next_train (data/fruit_datamanager.py:188-197): nerfstudio
#      data/pixel_samplers.py PixelSampler.sample_method / collate_image_dataset_batch followed by
#      model_components/ray_generators.py RayGenerator.forward.  Published nerfstudio 1.1.3 source restated,
#      not line-checked (nerfstudio is absent here).
# --------------------------------------------------------------------------------------------------------


def pixel_sampler_indices(rand3: Tensor, num_images: int, height: int, width: int) -> Tensor:
    """PixelSampler.sample_method, mask=None branch: ``(torch.rand((R, 3)) * tensor([N, H, W])).long()`` -> [R,3] (camera, y, x)."""
    return (rand3 * torch.tensor([num_images, height, width], dtype=rand3.dtype)).long()


def next_train_batch(rand3: Tensor, images: Tensor, masks: Optional[Tensor], c2w: Tensor, fx: Tensor, fy: Tensor, cx: Tensor, cy: Tensor):
    """``batch = pixel_sampler.sample(image_batch); ray_bundle = ray_generator(batch["indices"])`` for images [N,H,W,3] float32,
    fruit masks [N,H,W] (0/1) and per-camera c2w [N,3,4] / intrinsics [N].  Returns (indices, origins, directions, pixel_area, image, mask)."""
    n, h, w = images.shape[:3]
    idx = pixel_sampler_indices(rand3, n, h, w)
    c, y, x = idx[:, 0], idx[:, 1], idx[:, 2]
    image = images[c, y, x]                                     # collate_image_dataset_batch: value[c, y, x]
    mask = masks[c, y, x][:, None].to(images.dtype) if masks is not None else torch.zeros((idx.shape[0], 1), dtype=images.dtype)
    coords = torch.stack([y.to(rand3.dtype) + 0.5, x.to(rand3.dtype) + 0.5], -1)   # Cameras.get_image_coords()[y, x]: pixel centres
    yy, xx = coords[:, 0], coords[:, 1]
    fxc, fyc, cxc, cyc = fx[c], fy[c], cx[c], cy[c]
    coord = torch.stack([(xx - cxc) / fxc, -(yy - cyc) / fyc], -1)
    coord_x_offset = torch.stack([(xx - cxc + 1) / fxc, -(yy - cyc) / fyc], -1)
    coord_y_offset = torch.stack([(xx - cxc) / fxc, -(yy - cyc + 1) / fyc], -1)
    coord_stack = torch.stack([coord, coord_x_offset, coord_y_offset], dim=0)
    directions_stack = torch.cat([coord_stack, -torch.ones_like(coord_stack[..., :1])], dim=-1)
    rotation = c2w[c][:, :3, :3]
    directions_stack = torch.sum(directions_stack[..., None, :] * rotation, dim=-1)
    directions_stack = directions_stack / torch.linalg.norm(directions_stack, dim=-1, keepdim=True)
    directions = directions_stack[0]
    dx = torch.sqrt(torch.sum((directions - directions_stack[1]) ** 2, dim=-1))
    dy = torch.sqrt(torch.sum((directions - directions_stack[2]) ** 2, dim=-1))
    return idx, c2w[c][:, :3, 3], directions, (dx * dy)[..., None], image, mask
