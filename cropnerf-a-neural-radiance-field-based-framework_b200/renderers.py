"""Renderers and losses with nerfstudio's call signatures (rows a11-a14, a16) on the compositing kernels.

``nerfstudio/model_components/renderers.py``: RGBRenderer, AccumulationRenderer, DepthRenderer(method="median"),
SemanticRenderer, ``background_color_override_context`` (``fruit_nerf.py:170-174,560-591``;
``scripts/semantic_projection.py:51,169``).  ``nerfstudio/model_components/losses.py``: interlevel_loss,
distortion_loss, MSELoss (``fruit_nerf.py:177-178,601-615,639-645``).
"""
from __future__ import annotations

from typing import List, Optional, Union

import torch
from torch import Tensor, nn

from . import _lib as L
from . import ops
from .rays import RaySamples

BACKGROUND_COLOR_OVERRIDE: Optional[Tensor] = None


class background_color_override_context:  # noqa: N801 -- nerfstudio's name
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


def resolve_background(background_color: Union[str, Tensor]):
    """-> (bg_mode, host colour or None) honouring the global override (renderers.py RGBRenderer.combine_rgb)."""
    if BACKGROUND_COLOR_OVERRIDE is not None:
        background_color = BACKGROUND_COLOR_OVERRIDE
    if isinstance(background_color, str):
        if background_color == "random":
            return L.BG_NONE, None
        if background_color == "last_sample":
            return L.BG_LAST_SAMPLE, None
        if background_color == "black":
            return L.BG_CONSTANT, (0.0, 0.0, 0.0)
        if background_color == "white":
            return L.BG_CONSTANT, (1.0, 1.0, 1.0)
        raise ValueError(f"unknown background colour {background_color}")
    return L.BG_CONSTANT, tuple(float(v) for v in background_color.detach().flatten().tolist())


class RGBRenderer(nn.Module):
    def __init__(self, background_color: Union[str, Tensor] = "random") -> None:
        super().__init__()
        self.background_color = background_color

    def forward(self, rgb: Tensor, weights: Tensor, ray_indices=None, num_rays=None, background_color=None) -> Tensor:
        mode, color = resolve_background(background_color if background_color is not None else self.background_color)
        out, _, _ = ops.render(weights, rgb, None, mode, color, eval_mode=not self.training)
        return out


class AccumulationRenderer(nn.Module):
    def forward(self, weights: Tensor, ray_indices=None, num_rays=None) -> Tensor:
        _, acc, _ = ops.render(weights, None, None, L.BG_NONE, None, False)
        return acc


class SemanticRenderer(nn.Module):
    def forward(self, semantics: Tensor, weights: Tensor, ray_indices=None, num_rays=None) -> Tensor:
        _, _, sem = ops.render(weights, None, semantics, L.BG_NONE, None, False)
        return sem


class DepthRenderer(nn.Module):
    def __init__(self, method: str = "median") -> None:
        super().__init__()
        if method != "median":
            raise ValueError('only DepthRenderer(method="median") is used by FruitModel (fruit_nerf.py:172) and compiled')
        self.method = method

    def forward(self, weights: Tensor, ray_samples: RaySamples, ray_indices=None, num_rays=None) -> Tensor:
        return ops.render_median_depth(weights, ray_samples.frustums.starts, ray_samples.frustums.ends)


# ---------------------------------------------------------------------------------------------------------------


def ray_samples_to_sdist(ray_samples: RaySamples) -> Tensor:
    meta = ray_samples.metadata or {}
    if "_spacing_edges" in meta:
        return meta["_spacing_edges"]
    return torch.cat([ray_samples.spacing_starts[..., 0], ray_samples.spacing_ends[..., -1:, 0]], dim=-1).contiguous()


def interlevel_loss(weights_list: List[Tensor], ray_samples_list: List[RaySamples]) -> Tensor:
    c = ray_samples_to_sdist(ray_samples_list[-1]).detach()
    w = weights_list[-1][..., 0].detach()
    loss = None
    for ray_samples, weights in zip(ray_samples_list[:-1], weights_list[:-1]):
        cp = ray_samples_to_sdist(ray_samples)
        term = ops.interlevel_term(c, w, cp, weights[..., 0])
        loss = term if loss is None else loss + term
    return loss


def distortion_loss(weights_list: List[Tensor], ray_samples_list: List[RaySamples]) -> Tensor:
    c = ray_samples_to_sdist(ray_samples_list[-1])
    w = weights_list[-1][..., 0]
    return ops.distortion(c, w)


class PixelLosses(nn.Module):
    """``MSELoss()(image, rgb)`` and ``semantic_loss_weight * BCEWithLogitsLoss(mean)(sem, fruit_mask)`` in one kernel
    (fruit_nerf.py:177-178,604-608)."""

    def forward(self, rgb: Tensor, sem: Tensor, image: Tensor, mask: Tensor, sem_weight: float = 1.0):
        return ops.pixel_losses(rgb, sem, image, mask, sem_weight)
