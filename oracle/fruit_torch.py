"""Oracle restatement of the reference's own wiring of the hot path (FruitField + FruitModel).

TEST INFRASTRUCTURE (see ``oracle/__init__.py``) -- parity unpinned.

Follows, in behaviour (not in text), ``crop_nerf/fruit_nerf/fruit_field.py:71-302`` (module wiring, the
detach / mask / ordering rules) and ``crop_nerf/fruit_nerf/fruit_nerf.py:87-183,476-645`` (model wiring, output
dict, losses), with hyper-parameters from ``fruit_nerf_config.py:29-65`` and the nerfacto defaults listed in
SURVEY.md section 8.  Parameter names match the reference state-dict so the same
``state_dict`` loads into the product modules.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch
from torch import Tensor, nn

from . import nerfstudio_torch as ns


@dataclass
class FruitNerfModelConfig:
    """Fields of FruitNerfModelConfig(NerfactoModelConfig) that reach the hot path (fruit_nerf.py:59-68 plus
    inherited nerfacto defaults, SURVEY.md section 8 preamble)."""

    near_plane: float = 0.05
    far_plane: float = 1000.0
    background_color: str = "last_sample"
    num_levels: int = 16
    base_res: int = 16
    max_res: int = 2048
    log2_hashmap_size: int = 19
    features_per_level: int = 2
    num_proposal_samples_per_ray: Tuple[int, ...] = (256, 96)
    num_nerf_samples_per_ray: int = 48
    proposal_update_every: int = 5
    proposal_warmup: int = 5000
    num_proposal_iterations: int = 2
    use_same_proposal_network: bool = False
    proposal_net_args_list: List[Dict] = field(
        default_factory=lambda: [
            {"hidden_dim": 16, "log2_hashmap_size": 17, "num_levels": 5, "max_res": 128, "use_linear": False},
            {"hidden_dim": 16, "log2_hashmap_size": 17, "num_levels": 5, "max_res": 256, "use_linear": False},
        ]
    )
    interlevel_loss_mult: float = 1.0
    use_proposal_weight_anneal: bool = True
    use_average_appearance_embedding: bool = True
    proposal_weights_anneal_slope: float = 10.0
    proposal_weights_anneal_max_num_iters: int = 1000
    use_single_jitter: bool = True
    disable_scene_contraction: bool = False
    eval_num_rays_per_chunk: int = 1 << 15
    # FruitNerfModelConfig additions (fruit_nerf.py:63-68)
    semantic_loss_weight: float = 1.0
    pass_semantic_gradients: bool = False
    num_layers_semantic: int = 2
    hidden_dim_semantics: int = 64
    geo_feat_dim: int = 15


class FruitField(nn.Module):
    """fruit_field.py:44-302 -- hash grid -> base MLP -> (density, geo) ; semantic MLP + head ; SH+geo+emb -> RGB."""

    def __init__(
        self,
        aabb: Tensor,
        num_images: int,
        num_layers: int = 2,
        hidden_dim: int = 64,
        geo_feat_dim: int = 15,
        num_levels: int = 16,
        base_res: int = 16,
        max_res: int = 2048,
        log2_hashmap_size: int = 19,
        num_layers_color: int = 3,
        num_layers_semantic: int = 2,
        features_per_level: int = 2,
        hidden_dim_color: int = 64,
        hidden_dim_semantics: int = 64,
        hidden_dim_transient: int = 64,
        appearance_embedding_dim: int = 32,
        use_semantics: bool = True,
        test_mode: Optional[str] = None,
        num_semantic_classes: int = 1,
        pass_semantic_gradients: bool = False,
        use_average_appearance_embedding: bool = False,
        spatial_distortion: Optional[nn.Module] = None,
    ) -> None:
        super().__init__()
        self.register_buffer("aabb", aabb)
        self.register_buffer("max_res", torch.tensor(max_res))
        self.register_buffer("num_levels", torch.tensor(num_levels))
        self.register_buffer("log2_hashmap_size", torch.tensor(log2_hashmap_size))
        self.geo_feat_dim = geo_feat_dim
        self.spatial_distortion = spatial_distortion
        self.num_images = num_images
        self.appearance_embedding_dim = appearance_embedding_dim
        self.use_average_appearance_embedding = use_average_appearance_embedding
        self.use_semantics = use_semantics
        self.test_mode = test_mode
        self.pass_semantic_gradients = pass_semantic_gradients
        # construction order follows fruit_field.py:106-167 so torch's RNG stream hands out the same inits
        self.embedding_appearance = ns.Embedding(num_images, appearance_embedding_dim)
        self.direction_encoding = ns.SHEncoding(levels=4)
        self.mlp_base_grid = ns.HashEncoding(
            num_levels=num_levels,
            min_res=base_res,
            max_res=max_res,
            log2_hashmap_size=log2_hashmap_size,
            features_per_level=features_per_level,
        )
        self.mlp_base_mlp = ns.MLP(
            in_dim=self.mlp_base_grid.get_out_dim(),
            num_layers=num_layers,
            layer_width=hidden_dim,
            out_dim=1 + geo_feat_dim,
            activation=nn.ReLU(),
            out_activation=None,
        )
        self.mlp_base = nn.Sequential(self.mlp_base_grid, self.mlp_base_mlp)
        if use_semantics:
            self.mlp_semantics = ns.MLP(
                in_dim=geo_feat_dim,
                num_layers=num_layers_semantic,
                layer_width=hidden_dim_semantics,
                out_dim=hidden_dim_transient,
                activation=nn.ReLU(),
                out_activation=None,
            )
            self.field_head_semantics = ns.FieldHead(in_dim=self.mlp_semantics.get_out_dim(), out_dim=num_semantic_classes)
        self.mlp_head = ns.MLP(
            in_dim=self.direction_encoding.get_out_dim() + geo_feat_dim + appearance_embedding_dim,
            num_layers=num_layers_color,
            layer_width=hidden_dim_color,
            out_dim=3,
            activation=nn.ReLU(),
            out_activation=nn.Sigmoid(),
        )

    # fruit_field.py:169-194
    def get_density(self, ray_samples: ns.RaySamples) -> Tuple[Tensor, Tensor]:
        pos = ray_samples.frustums.get_positions()
        if self.spatial_distortion is not None:
            pos = (self.spatial_distortion(pos) + 2.0) / 4.0
        else:
            pos = ns.get_normalized_positions(pos, self.aabb)
        selector = ((pos > 0.0) & (pos < 1.0)).all(dim=-1)
        pos = pos * selector[..., None]
        self._sample_locations = pos
        h = self.mlp_base(pos.view(-1, 3)).view(*ray_samples.frustums.shape, -1)
        dba, geo = torch.split(h, [1, self.geo_feat_dim], dim=-1)
        self._density_before_activation = dba
        density = ns.trunc_exp(dba.to(pos)) * selector[..., None]
        return density, geo

    def _heads(self, ray_samples: ns.RaySamples, geo: Tensor, appearance: Tensor) -> Dict[str, Tensor]:
        shape = ray_samples.frustums.directions.shape[:-1]
        d = self.direction_encoding(ns.get_normalized_directions(ray_samples.frustums.directions).reshape(-1, 3))
        sem_in = geo.reshape(-1, self.geo_feat_dim)
        if not self.pass_semantic_gradients:
            sem_in = sem_in.detach()
        x = self.mlp_semantics(sem_in).view(*shape, -1)
        out = {"semantics": self.field_head_semantics(x)}
        h = torch.cat([d, geo.reshape(-1, self.geo_feat_dim), appearance.reshape(-1, self.appearance_embedding_dim)], dim=-1)
        out["rgb"] = self.mlp_head(h).view(*shape, -1)
        return out

    # fruit_field.py:235-282
    def get_outputs(self, ray_samples: ns.RaySamples, density_embedding: Tensor) -> Dict[str, Tensor]:
        if ray_samples.camera_indices is None:
            raise AttributeError("Camera indices are not provided.")
        shape = ray_samples.frustums.directions.shape[:-1]
        if self.training:
            app = self.embedding_appearance(ray_samples.camera_indices.squeeze(-1))
        elif self.use_average_appearance_embedding:
            app = torch.ones((*shape, self.appearance_embedding_dim), dtype=density_embedding.dtype) * self.embedding_appearance.mean(dim=0)
        else:
            app = torch.zeros((*shape, self.appearance_embedding_dim), dtype=density_embedding.dtype)
        return self._heads(ray_samples, density_embedding, app)

    # fruit_field.py:196-233 (always the mean embedding)
    def get_inference_outputs(self, ray_samples: ns.RaySamples, density_embedding: Tensor) -> Dict[str, Tensor]:
        shape = ray_samples.frustums.directions.shape[:-1]
        app = torch.ones((*shape, self.appearance_embedding_dim), dtype=density_embedding.dtype) * self.embedding_appearance.mean(dim=0)
        return self._heads(ray_samples, density_embedding, app)

    # fruit_field.py:284-302
    def forward(self, ray_samples: ns.RaySamples) -> Dict[str, Tensor]:
        density, geo = self.get_density(ray_samples)
        if self.test_mode in ("inference", "export"):
            out = self.get_inference_outputs(ray_samples, geo)
        else:
            out = self.get_outputs(ray_samples, geo)
        out["density"] = density
        return out


class FruitModel(nn.Module):
    """fruit_nerf.py:71-645 restricted to the ray-render path (camera optimizer = identity: mode "off")."""

    def __init__(self, config: FruitNerfModelConfig, num_train_data: int, aabb: Optional[Tensor] = None, test_mode: str = "val"):
        super().__init__()
        self.config = config
        self.test_mode = test_mode
        self.num_train_data = num_train_data
        aabb = aabb if aabb is not None else torch.tensor([[-1.0, -1.0, -1.0], [1.0, 1.0, 1.0]])
        contraction = None if config.disable_scene_contraction else ns.SceneContraction(order=float("inf"))
        # fruit_nerf.py:97-112 -- only these kwargs are forwarded; the rest are FruitField defaults
        self.field = FruitField(
            aabb,
            num_levels=config.num_levels,
            max_res=config.max_res,
            num_layers_semantic=config.num_layers_semantic,
            hidden_dim_semantics=config.hidden_dim_semantics,
            log2_hashmap_size=config.log2_hashmap_size,
            spatial_distortion=contraction,
            num_images=num_train_data,
            geo_feat_dim=config.geo_feat_dim,
            use_average_appearance_embedding=config.use_average_appearance_embedding,
            use_semantics=True,
            test_mode=test_mode,
            num_semantic_classes=1,
            pass_semantic_gradients=config.pass_semantic_gradients,
        )
        # fruit_nerf.py:118-142
        self.proposal_networks = nn.ModuleList()
        n = config.num_proposal_iterations
        if config.use_same_proposal_network:
            net = ns.HashMLPDensityField(aabb, spatial_distortion=contraction, **config.proposal_net_args_list[0])
            self.proposal_networks.append(net)
            self.density_fns = [net.density_fn for _ in range(n)]
        else:
            for i in range(n):
                args = config.proposal_net_args_list[min(i, len(config.proposal_net_args_list) - 1)]
                self.proposal_networks.append(ns.HashMLPDensityField(aabb, spatial_distortion=contraction, **args))
            self.density_fns = [net.density_fn for net in self.proposal_networks]

        def update_schedule(step):  # fruit_nerf.py:144-149
            return np.clip(np.interp(step, [0, config.proposal_warmup], [0, config.proposal_update_every]), 1, config.proposal_update_every)

        self.proposal_sampler = ns.ProposalNetworkSampler(
            num_nerf_samples_per_ray=config.num_nerf_samples_per_ray,
            num_proposal_samples_per_ray=config.num_proposal_samples_per_ray,
            num_proposal_network_iterations=config.num_proposal_iterations,
            single_jitter=config.use_single_jitter,
            update_sched=update_schedule,
        )
        self.collider = ns.NearFarCollider(near_plane=config.near_plane, far_plane=config.far_plane)
        self.renderer_rgb = ns.RGBRenderer(background_color=config.background_color)
        self.renderer_accumulation = ns.AccumulationRenderer()
        self.renderer_depth = ns.DepthRenderer(method="median")
        self.renderer_semantics = ns.SemanticRenderer()
        self.rgb_loss = nn.MSELoss()
        self.binary_cross_entropy_loss = nn.BCEWithLogitsLoss(reduction="mean")
        self.register_buffer("colormap", torch.tensor([0.0, 1.0]))  # cotton_nerf_dataparser.py:248-255

    # fruit_nerf.py:185-189
    def setup_inference(self, render_rgb: bool, num_inference_samples: int) -> None:
        self.proposal_sampler = ns.UniformSamplerWithNoise(num_samples=num_inference_samples, single_jitter=False)
        self.field.spatial_distortion = None

    def get_param_groups(self):
        return {"proposal_networks": list(self.proposal_networks.parameters()), "fields": list(self.field.parameters())}

    def set_anneal(self, step: int) -> None:  # fruit_nerf.py:206-216
        N = self.config.proposal_weights_anneal_max_num_iters
        frac = float(np.clip(step / N, 0, 1))
        b = self.config.proposal_weights_anneal_slope
        self.proposal_sampler.set_anneal(b * frac / ((b - 1) * frac + 1))

    def _semantic_colormap(self, sem: Tensor) -> Tensor:  # fruit_nerf.py:594-597
        labels = torch.heaviside(torch.sigmoid(sem.detach()) - 0.9, torch.tensor(0.0, dtype=sem.dtype)).to(torch.long)
        return self.colormap[labels].repeat(1, 3)

    def _render(self, ray_bundle: ns.RayBundle, depth_no_grad: bool, keep_lists: bool) -> Dict:
        ray_samples, weights_list, ray_samples_list = self.proposal_sampler(ray_bundle, density_fns=self.density_fns)
        fo = self.field.forward(ray_samples)
        weights = ray_samples.get_weights(fo["density"])
        weights_list.append(weights)
        ray_samples_list.append(ray_samples)
        rgb = self.renderer_rgb(rgb=fo["rgb"], weights=weights)
        if depth_no_grad:
            with torch.no_grad():
                depth = self.renderer_depth(weights=weights, ray_samples=ray_samples)
        else:
            depth = self.renderer_depth(weights=weights, ray_samples=ray_samples)
        out = {"rgb": rgb, "accumulation": self.renderer_accumulation(weights=weights), "depth": depth}
        if keep_lists:
            out["weights_list"] = weights_list
            out["ray_samples_list"] = ray_samples_list
        for i in range(self.config.num_proposal_iterations):
            out[f"prop_depth_{i}"] = self.renderer_depth(weights=weights_list[i], ray_samples=ray_samples_list[i])
        sw = weights if self.config.pass_semantic_gradients else weights.detach()
        out["semantics"] = self.renderer_semantics(fo["semantics"], weights=sw)
        out["semantics_colormap"] = self._semantic_colormap(out["semantics"])
        return out

    # fruit_nerf.py:543-599
    def get_outputs(self, ray_bundle: ns.RayBundle) -> Dict:
        return self._render(ray_bundle, depth_no_grad=True, keep_lists=self.training)

    # fruit_nerf.py:497-541
    def get_inference_outputs(self, ray_bundle: ns.RayBundle) -> Dict:
        return self._render(ray_bundle, depth_no_grad=False, keep_lists=True)

    # fruit_nerf.py:476-494
    def get_export_outputs(self, ray_bundle: ns.RayBundle) -> Dict:
        ray_samples = self.proposal_sampler(ray_bundle)
        fo = self.field.forward(ray_samples)
        out = {
            "rgb": fo["rgb"],
            "point_location": ray_samples.frustums.get_positions(),
            "semantics": fo["semantics"][..., 0],
            "density": fo["density"][..., 0],
        }
        out["semantics_colormap"] = torch.heaviside(
            torch.sigmoid(out["semantics"]) - 0.9, torch.tensor(0.0, dtype=out["semantics"].dtype)
        ).to(torch.long)
        return out

    # fruit_nerf.py:320-344 (one chunk)
    def get_density_for_ray_bundle(self, ray_bundle: ns.RayBundle) -> Tensor:
        ray_samples, _, _ = self.proposal_sampler(ray_bundle, density_fns=self.density_fns)
        fo = self.field.forward(ray_samples)
        weights = ray_samples.get_weights(fo["density"])
        return weights.squeeze(-1).sum(-1)

    # fruit_nerf.py:617-637
    def forward(self, ray_bundle: ns.RayBundle) -> Dict:
        ray_bundle = self.collider(ray_bundle)
        if self.test_mode == "inference":
            return self.get_inference_outputs(ray_bundle)
        if self.test_mode == "export":
            return self.get_export_outputs(ray_bundle)
        return self.get_outputs(ray_bundle)

    # fruit_nerf.py:601-615
    def get_loss_dict(self, outputs: Dict, batch: Dict) -> Dict[str, Tensor]:
        loss = {"rgb_loss": self.rgb_loss(batch["image"][:, :3], outputs["rgb"])}
        loss["semantics_loss"] = self.config.semantic_loss_weight * self.binary_cross_entropy_loss(
            outputs["semantics"], batch["fruit_mask"]
        )
        if self.training:
            loss["interlevel_loss"] = self.config.interlevel_loss_mult * ns.interlevel_loss(
                outputs["weights_list"], outputs["ray_samples_list"]
            )
        return loss

    # fruit_nerf.py:639-645 (PSNR via torchmetrics = 10*log10(1/mse) at data_range 1)
    def get_metrics_dict(self, outputs: Dict, batch: Dict) -> Dict[str, Tensor]:
        mse = torch.mean((outputs["rgb"] - batch["image"][:, :3]) ** 2)
        return {
            "psnr": 10.0 * torch.log10(1.0 / mse),
            "distortion": ns.distortion_loss(outputs["weights_list"], outputs["ray_samples_list"]),
        }
