"""Ray samplers with nerfstudio's interfaces on the sampler kernels (rows a8 / a9).

* ``SpacedSampler`` / ``UniformSampler`` / ``UniformLinDispPiecewiseSampler`` / ``PDFSampler`` /
  ``ProposalNetworkSampler``: ``nerfstudio/model_components/ray_samplers.py`` as used at ``fruit_nerf.py:155-164``.
* ``UniformSamplerWithNoise``: the reference's own ``components/ray_samplers.py:31-104`` (installed by
  ``FruitModel.setup_inference``, ``fruit_nerf.py:185-189``).

Random jitter is drawn with ``torch.rand`` exactly where nerfstudio draws it (shape ``[R,1]`` for single jitter,
``[R,S+1]`` otherwise) through an injectable ``rand_fn`` so tests can feed the oracle the same numbers.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Tuple

import torch
from torch import Tensor, nn

from . import _lib as L
from . import ops
from .density_fields import HashMLPDensityField
from .rays import RayBundle, RaySamples, ray_layout, samples_from_edges


class _SpacingFn:
    """``spacing_to_euclidean_fn`` closure of SpacedSampler, parametric so the PDF kernel can evaluate it in-kernel."""

    def __init__(self, kind: int, nears: Tensor, fars: Tensor):
        self.kind = kind
        self.nears = nears
        self.fars = fars

    def _fn(self, x):
        return torch.where(x < 1, x / 2, 1 - 1 / (2 * x)) if self.kind == L.SPACING_LINDISP_PIECEWISE else x

    def _inv(self, x):
        return torch.where(x < 0.5, 2 * x, 1 / (2 - 2 * x)) if self.kind == L.SPACING_LINDISP_PIECEWISE else x

    def __call__(self, x: Tensor) -> Tensor:  # API parity for external callers; kernels never call this
        s_near, s_far = self._fn(self.nears), self._fn(self.fars)
        return self._inv(x * s_far + (1 - x) * s_near)


class SpacedSampler(nn.Module):
    spacing_kind = L.SPACING_UNIFORM

    def __init__(self, spacing_fn: Optional[Callable] = None, spacing_fn_inv: Optional[Callable] = None, num_samples: Optional[int] = None,
                 train_stratified: bool = True, single_jitter: bool = False) -> None:
        super().__init__()
        self.num_samples = num_samples
        self.train_stratified = train_stratified
        self.single_jitter = single_jitter
        # kept for subclasses that override generate_ray_samples with their own torch code (the reference's
        # components/ray_samplers.py:31-104 UniformSamplerWithNoise does); the kernels use spacing_kind
        self.spacing_fn = spacing_fn if spacing_fn is not None else (lambda x: x)
        self.spacing_fn_inv = spacing_fn_inv if spacing_fn_inv is not None else (lambda x: x)
        self.rand_fn = torch.rand
        self._lin = {}

    def _lin_bins(self, S: int, device) -> Tensor:
        key = (S, str(device))
        if key not in self._lin:  # torch.linspace on the host, uploaded once: bit-identical to the reference's bins
            self._lin[key] = torch.linspace(0.0, 1.0, S + 1).to(device)
        return self._lin[key]

    def generate_ray_samples(self, ray_bundle: Optional[RayBundle] = None, num_samples: Optional[int] = None) -> RaySamples:
        assert ray_bundle is not None and ray_bundle.nears is not None and ray_bundle.fars is not None
        num_samples = num_samples or self.num_samples
        assert num_samples is not None
        R = ray_bundle.origins.shape[0]
        dev = ray_bundle.origins.device
        t_rand = None
        if self.train_stratified and self.training:
            shape = (R, 1) if self.single_jitter else (R, num_samples + 1)
            t_rand = self.rand_fn(shape, dtype=torch.float32, device=dev)
        sp, eu = ops.sample_spaced(ray_bundle.nears, ray_bundle.fars, self._lin_bins(num_samples, dev), t_rand, self.spacing_kind)
        return samples_from_edges(ray_bundle, eu, sp, _SpacingFn(self.spacing_kind, ray_bundle.nears, ray_bundle.fars))

    def forward(self, *args, **kwargs) -> RaySamples:
        return self.generate_ray_samples(*args, **kwargs)


class UniformSampler(SpacedSampler):
    def __init__(self, num_samples: Optional[int] = None, train_stratified: bool = True, single_jitter: bool = False) -> None:
        super().__init__(num_samples=num_samples, train_stratified=train_stratified, single_jitter=single_jitter)


class UniformSamplerWithNoise(UniformSampler):
    """components/ray_samplers.py:31-104 (identity spacing; stratified jitter in training)."""


class UniformLinDispPiecewiseSampler(SpacedSampler):
    spacing_kind = L.SPACING_LINDISP_PIECEWISE

    def __init__(self, num_samples: Optional[int] = None, train_stratified: bool = True, single_jitter: bool = False) -> None:
        super().__init__(num_samples=num_samples, train_stratified=train_stratified, single_jitter=single_jitter)


class PDFSampler(nn.Module):
    def __init__(self, num_samples: Optional[int] = None, train_stratified: bool = True, single_jitter: bool = False,
                 include_original: bool = True, histogram_padding: float = 0.01) -> None:
        super().__init__()
        if include_original:
            raise ValueError("include_original=True is not used by ProposalNetworkSampler and is not compiled")
        self.num_samples = num_samples
        self.train_stratified = train_stratified
        self.single_jitter = single_jitter
        self.include_original = include_original
        self.histogram_padding = histogram_padding
        self.rand_fn = torch.rand
        self.keep_inds = False
        self.last_inds: Optional[Tensor] = None
        self._u = {}

    def _u_base(self, num_bins: int, device) -> Tensor:
        key = (num_bins, str(device))
        if key not in self._u:
            self._u[key] = torch.linspace(0.0, 1.0 - (1.0 / num_bins), steps=num_bins).to(device)
        return self._u[key]

    def generate_ray_samples(self, ray_bundle: Optional[RayBundle] = None, ray_samples: Optional[RaySamples] = None,
                             weights: Optional[Tensor] = None, num_samples: Optional[int] = None, eps: float = 1e-5,
                             anneal: float = 1.0) -> RaySamples:
        assert ray_bundle is not None and ray_samples is not None and weights is not None
        num_samples = num_samples or self.num_samples
        assert num_samples is not None
        fn = ray_samples.spacing_to_euclidean_fn
        if not isinstance(fn, _SpacingFn):
            raise RuntimeError("cropnerf_b200 PDFSampler needs ray samples produced by a cropnerf_b200 SpacedSampler/PDFSampler")
        dev = weights.device
        R = weights.shape[0]
        meta = ray_samples.metadata or {}
        prev = meta.get("_spacing_edges")
        if prev is None:
            prev = torch.cat([ray_samples.spacing_starts[..., 0], ray_samples.spacing_ends[..., -1:, 0]], dim=-1).contiguous()
        rand = None
        if self.train_stratified and self.training:
            shape = (R, 1) if self.single_jitter else (R, num_samples + 1)
            rand = self.rand_fn(shape, device=dev, dtype=torch.float32)
        sp, eu, inds = ops.sample_pdf(weights, anneal, prev, fn.nears, fn.fars, fn.kind, self._u_base(num_samples + 1, dev), rand, num_samples,
                                      self.histogram_padding, eps, want_inds=self.keep_inds)
        self.last_inds = inds
        return samples_from_edges(ray_bundle, eu, sp, fn)

    def forward(self, *args, **kwargs) -> RaySamples:
        return self.generate_ray_samples(*args, **kwargs)


class ProposalNetworkSampler(nn.Module):
    """nerfstudio ProposalNetworkSampler (built fruit_nerf.py:157-164; called :549,501,429,337)."""

    def __init__(self, num_proposal_samples_per_ray: Tuple[int, ...] = (64,), num_nerf_samples_per_ray: int = 32,
                 num_proposal_network_iterations: int = 2, single_jitter: bool = False, update_sched: Callable = lambda x: 1,
                 initial_sampler: Optional[nn.Module] = None, pdf_sampler: Optional[PDFSampler] = None) -> None:
        super().__init__()
        self.num_proposal_samples_per_ray = num_proposal_samples_per_ray
        self.num_nerf_samples_per_ray = num_nerf_samples_per_ray
        self.num_proposal_network_iterations = num_proposal_network_iterations
        self.update_sched = update_sched
        if self.num_proposal_network_iterations < 1:
            raise ValueError("num_proposal_network_iterations must be >= 1")
        self.initial_sampler = initial_sampler if initial_sampler is not None else UniformLinDispPiecewiseSampler(single_jitter=single_jitter)
        self.pdf_sampler = pdf_sampler if pdf_sampler is not None else PDFSampler(include_original=False, single_jitter=single_jitter)
        self._anneal = 1.0
        self._steps_since_update = 0
        self._step = 0

    def set_anneal(self, anneal: float) -> None:
        self._anneal = anneal

    def step_cb(self, step) -> None:
        self._step = step
        self._steps_since_update += 1

    @staticmethod
    def _density(fn: Callable, ray_samples: RaySamples) -> Tensor:
        owner = getattr(fn, "__self__", None)
        if isinstance(owner, HashMLPDensityField) and getattr(fn, "__func__", None) is HashMLPDensityField.density_fn:
            # fast path: the fused kernel derives the positions from rays + bin edges itself
            return owner.density_from_layout(ray_layout(ray_samples)).view(*ray_samples.frustums.shape, 1)
        return fn(ray_samples.frustums.get_positions())  # arbitrary density_fn (e.g. BayesRays' wrapped ones)

    def generate_ray_samples(self, ray_bundle: Optional[RayBundle] = None, density_fns: Optional[List[Callable]] = None):
        assert ray_bundle is not None and density_fns is not None
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
                # weights ** anneal is folded into the resampling kernel
                ray_samples = self.pdf_sampler(ray_bundle, ray_samples, weights, num_samples=num_samples, anneal=self._anneal)
            if is_prop:
                if updated:
                    density = self._density(density_fns[i_level], ray_samples)
                else:
                    with torch.no_grad():
                        density = self._density(density_fns[i_level], ray_samples)
                weights = ray_samples.get_weights(density)
                weights_list.append(weights)
                ray_samples_list.append(ray_samples)
        if updated:
            self._steps_since_update = 0
        return ray_samples, weights_list, ray_samples_list

    def forward(self, *args, **kwargs):
        return self.generate_ray_samples(*args, **kwargs)
