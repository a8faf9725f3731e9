"""Field components with nerfstudio's constructor signatures and state-dict names, backed by the CUDA ops.

Mirrors the pieces of ``nerfstudio/field_components`` the reference imports at ``fruit_field.py:24-39``:
``HashEncoding`` (encodings.py), ``MLP`` (mlp.py), ``SHEncoding`` (encodings.py), ``Embedding`` (embedding.py),
``FieldHead`` / ``SemanticFieldHead`` (field_heads.py, ``components/field_heads.py:29-40``), ``SceneContraction``
(spatial_distortions.py).  Only ``implementation="torch"`` *semantics* exist (hashed coarse levels, biased Linear
layers, SH on the shifted direction -- SURVEY.md App. B-1); the arithmetic runs in ``libcropnerf_b200.so``.
"""
from __future__ import annotations

from enum import Enum
from typing import List, Optional

import numpy as np
import torch
from torch import Tensor, nn

from . import _lib as L
from . import ops

try:  # use nerfstudio's enum when it is installed so dict keys compare equal
    from nerfstudio.field_components.field_heads import FieldHeadNames  # type: ignore
except Exception:  # pragma: no cover - nerfstudio is absent in this image

    class FieldHeadNames(Enum):
        RGB = "rgb"
        SH = "sh"
        DENSITY = "density"
        NORMALS = "normals"
        PRED_NORMALS = "pred_normals"
        UNCERTAINTY = "uncertainty"
        BACKGROUND_RGB = "background_rgb"
        TRANSIENT_RGB = "transient_rgb"
        TRANSIENT_DENSITY = "transient_density"
        SEMANTICS = "semantics"
        SDF = "sdf"
        ALPHA = "alpha"
        GRADIENT = "gradient"


def hash_scalings(num_levels: int, min_res: int, max_res: int) -> Tensor:
    """``floor(min_res * growth**levels)`` evaluated the way nerfstudio's HashEncoding.__init__ does: float64 growth
    factor, float32 power (python float ** int64 tensor) -- the field's top level is 2047, not 2048 (App. B-2)."""
    levels = torch.arange(num_levels)
    growth = np.exp((np.log(max_res) - np.log(min_res)) / (num_levels - 1)) if num_levels > 1 else 1.0
    return torch.floor(min_res * growth**levels)


class HashEncoding(nn.Module):
    """Multiresolution hash grid (row a2).  ``forward`` = ``cnb_hashgrid_fwd`` with autograd to the table."""

    def __init__(self, num_levels: int = 16, min_res: int = 16, max_res: int = 1024, log2_hashmap_size: int = 19,
                 features_per_level: int = 2, hash_init_scale: float = 0.001, implementation: str = "b200",
                 interpolation: Optional[str] = None) -> None:
        super().__init__()
        if features_per_level != 2:
            raise ValueError("cropnerf_b200 HashEncoding is compiled for features_per_level=2 (fruit_field.py:92)")
        if num_levels > L.MAX_LEVELS:
            raise ValueError(f"num_levels {num_levels} > {L.MAX_LEVELS}")
        self.in_dim = 3
        self.num_levels = num_levels
        self.min_res = min_res
        self.features_per_level = features_per_level
        self.hash_init_scale = hash_init_scale
        self.log2_hashmap_size = log2_hashmap_size
        self.hash_table_size = 2**log2_hashmap_size
        self.register_buffer("scalings", hash_scalings(num_levels, min_res, max_res))
        self.hash_offset = torch.arange(num_levels) * self.hash_table_size
        table = torch.rand(size=(self.hash_table_size * num_levels, features_per_level)) * 2 - 1
        table *= hash_init_scale
        self.hash_table = nn.Parameter(table)
        self._scalings_host = [float(v) for v in self.scalings.tolist()]

    def get_out_dim(self) -> int:
        return self.num_levels * self.features_per_level

    def grid_cfg(self):
        return (self.num_levels, self.log2_hashmap_size, tuple(self._scalings_host))

    def forward(self, in_tensor: Tensor) -> Tensor:
        return ops.hashgrid_encode(in_tensor, self.hash_table, self.num_levels, self.log2_hashmap_size, self._scalings_host)

    def corner_indices(self, in_tensor: Tensor) -> Tensor:
        return ops.hashgrid_indices(in_tensor, self.hash_table, self.num_levels, self.log2_hashmap_size, self._scalings_host)[1]


_ACT = {type(None): L.ACT_NONE, nn.ReLU: L.ACT_RELU, nn.Sigmoid: L.ACT_SIGMOID}


class MLP(nn.Module):
    """``nn.Linear`` stack with nerfstudio's naming (``layers.{i}.weight/bias``), ReLU between layers (row a3)."""

    def __init__(self, in_dim: int, num_layers: int, layer_width: int, out_dim: Optional[int] = None, skip_connections=None,
                 activation: Optional[nn.Module] = nn.ReLU(), out_activation: Optional[nn.Module] = None, implementation: str = "b200") -> None:
        super().__init__()
        if skip_connections:
            raise ValueError("skip connections are not used by FruitField and are not compiled")
        if activation is not None and not isinstance(activation, nn.ReLU):
            raise ValueError("hidden activation must be ReLU")
        if type(out_activation) not in _ACT:
            raise ValueError(f"unsupported out_activation {out_activation}")
        self.in_dim = in_dim
        self.out_dim = out_dim if out_dim is not None else layer_width
        self.num_layers = num_layers
        self.layer_width = layer_width
        self.out_activation = out_activation
        self.out_act_code = _ACT[type(out_activation)]
        layers: List[nn.Module] = []
        if num_layers == 1:
            layers.append(nn.Linear(in_dim, self.out_dim))
        else:
            for i in range(num_layers - 1):
                layers.append(nn.Linear(in_dim if i == 0 else layer_width, layer_width))
            layers.append(nn.Linear(layer_width, self.out_dim))
        self.layers = nn.ModuleList(layers)

    def get_out_dim(self) -> int:
        return self.out_dim

    def weights(self) -> List[Tensor]:
        return [l.weight for l in self.layers]

    def biases(self) -> List[Tensor]:
        return [l.bias for l in self.layers]

    def forward(self, in_tensor: Tensor) -> Tensor:
        return ops.mlp_forward(in_tensor, self.weights(), self.biases(), self.out_act_code)


class FieldHead(nn.Module):
    def __init__(self, out_dim: int, field_head_name, in_dim: Optional[int] = None, activation=None) -> None:
        super().__init__()
        self.out_dim = out_dim
        self.in_dim = in_dim
        self.field_head_name = field_head_name
        self.activation = activation
        if type(activation) not in _ACT:
            raise ValueError(f"unsupported head activation {activation}")
        self.net = nn.Linear(in_dim, out_dim)

    def forward(self, in_tensor: Tensor) -> Tensor:
        return ops.mlp_forward(in_tensor, [self.net.weight], [self.net.bias], _ACT[type(self.activation)])


class SemanticFieldHead(FieldHead):
    """``components/field_heads.py:29-40``: Linear(in_dim -> num_classes), no activation."""

    def __init__(self, num_classes: int, in_dim: Optional[int] = None, activation=None) -> None:
        super().__init__(in_dim=in_dim, out_dim=num_classes, field_head_name=FieldHeadNames.SEMANTICS, activation=activation)


class SHEncoding(nn.Module):
    """Parameter-free; the degree-3 real SH basis on the shifted direction is evaluated inside the field kernels
    (``csrc/field_common.cuh`` cnb_sh16).  Kept for constructor / ``get_out_dim`` parity (fruit_field.py:116-119)."""

    def __init__(self, levels: int = 4, implementation: str = "b200") -> None:
        super().__init__()
        if levels != 4:
            raise ValueError("cropnerf_b200 evaluates SHEncoding(levels=4) only")
        self.levels = levels

    def get_out_dim(self) -> int:
        return self.levels**2


class Embedding(nn.Module):
    def __init__(self, in_dim: int, out_dim: int) -> None:
        super().__init__()
        self.in_dim = in_dim
        self.out_dim = out_dim
        self.embedding = nn.Embedding(in_dim, out_dim)

    def mean(self, dim=0):
        return self.embedding.weight.mean(dim)


class SceneContraction(nn.Module):
    """Marker for the L-inf scene contraction (fruit_nerf.py:95); the warp itself runs in-kernel (cnb_warp_position)."""

    def __init__(self, order=float("inf")) -> None:
        super().__init__()
        if order != float("inf"):
            raise ValueError("only order=inf (nerfacto default) is compiled")
        self.order = order
