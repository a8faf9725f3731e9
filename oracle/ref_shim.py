"""Run the reference's OWN ``FruitField`` / ``FruitModel`` code with nerfstudio's primitives shimmed by this oracle.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  The reference (``/root/reference/crop_nerf/fruit_nerf``) cannot be
imported as is: every file imports ``nerfstudio`` (1.1.3, absent from the tree and from this image), ``torchmetrics``,
``nerfacc`` and ``segmentation.segmenter`` (open3d).  This module installs stand-in modules under those names whose hot-path
symbols are the oracle's restated primitives (``oracle/nerfstudio_torch.py``) and whose off-path symbols are inert stubs,
then imports the reference's ``fruit_field.py``, ``components/*.py`` and ``fruit_nerf.py`` FROM WHERE THEY LIE and runs them.

What this pins: the reference's *wiring* -- FruitField.get_density / get_outputs / get_inference_outputs / forward
(fruit_field.py:169-302), SemanticFieldHead (components/field_heads.py:29-40), UniformSamplerWithNoise
(components/ray_samplers.py:31-104), FruitModel.populate_modules / get_outputs / get_inference_outputs /
get_export_outputs / get_loss_dict / get_metrics_dict / setup_inference (fruit_nerf.py:87-645) -- is executed verbatim;
``oracle/fruit_torch.py`` (the restated wiring the CUDA path is tested against) must reproduce its outputs bit for bit.
What it cannot pin: the nerfstudio primitives themselves (still the restatement of SURVEY.md Appendix A).

``python -m oracle.ref_shim`` regenerates ``tests/golden/ref_*.npz`` (needs /root/reference; the fixtures travel, the
reference does not).
"""
from __future__ import annotations

import enum
import os
import sys
import types
from dataclasses import dataclass, field
from typing import Dict, List, Optional

import numpy as np
import torch
from torch import nn

from . import nerfstudio_torch as ns

REFERENCE_ROOT = "/root/reference/crop_nerf"
_installed = False


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "fruit_nerf", "fruit_field.py"))


def _module(name: str, **attrs) -> types.ModuleType:
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    parent, _, child = name.rpartition(".")
    if parent:
        if parent not in sys.modules:
            _module(parent)
        setattr(sys.modules[parent], child, m)
    return m


class _Inert:
    """Stand-in for off-path classes (metrics, viewers, ...): constructible, callable, attribute-tolerant."""

    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        raise RuntimeError("off-path stub called")

    def to(self, *a, **k):
        return self


def install_shims(backend: str = "oracle") -> None:
    """``backend="oracle"``: nerfstudio's names are bound to the restated primitives (pins the oracle's wiring to the reference's code).
    ``backend="product"``: the SAME names are bound to the cropnerf_b200 classes and ``fruit_nerf.fruit_field`` to the product's FruitField
    -- the drop-in demonstration: the reference's unmodified ``fruit_nerf.py`` builds and drives the B200 modules
    (``tests/ref_dropin.py``; one backend per process)."""
    global _installed
    if _installed:
        if _installed != backend:
            raise RuntimeError(f"shims already installed for backend {_installed!r}; run the other backend in a fresh process")
        return
    _installed = backend
    if backend == "product":
        _install_product_shims()
        return

    # ---- hot-path primitives = the oracle's restatement, with nerfstudio's constructor signatures ------------------
    class HashEncoding(ns.HashEncoding):
        def __init__(self, num_levels=16, min_res=16, max_res=1024, log2_hashmap_size=19, features_per_level=2, hash_init_scale=0.001,
                     implementation="torch", interpolation=None):
            super().__init__(num_levels, min_res, max_res, log2_hashmap_size, features_per_level, hash_init_scale)

    class MLP(ns.MLP):
        def __init__(self, in_dim, num_layers, layer_width, out_dim=None, skip_connections=None, activation=nn.ReLU(), out_activation=None,
                     implementation="torch"):
            assert skip_connections is None
            super().__init__(in_dim, num_layers, layer_width, out_dim, activation, out_activation)

    class SHEncoding(ns.SHEncoding):
        def __init__(self, levels=4, implementation="torch"):
            super().__init__(levels)

    class NeRFEncoding(nn.Module):  # constructed at fruit_field.py:121-123, never called
        def __init__(self, *a, **k):
            super().__init__()

    class FieldHeadNames(enum.Enum):
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

    class FieldComponent(nn.Module):
        pass

    class FieldHead(FieldComponent):
        """nerfstudio/field_components/field_heads.py FieldHead: Linear(in_dim, out_dim) then the optional activation."""

        def __init__(self, out_dim, field_head_name, in_dim=None, activation=None):
            super().__init__()
            self.out_dim, self.in_dim, self.activation, self.field_head_name = out_dim, in_dim, activation, field_head_name
            self.net = nn.Linear(in_dim, out_dim)

        def forward(self, in_tensor):
            out = self.net(in_tensor)
            return self.activation(out) if self.activation else out

    class Field(nn.Module):
        def __init__(self):
            super().__init__()
            self._sample_locations = None
            self._density_before_activation = None

        def density_fn(self, positions, times=None):
            raise RuntimeError("off-path")

    class SceneBox:
        def __init__(self, aabb):
            self.aabb = aabb

        @staticmethod
        def get_normalized_positions(positions, aabb):
            return ns.get_normalized_positions(positions, aabb)

    class SpatialDistortion(nn.Module):
        pass

    _module("nerfstudio")
    _module("nerfstudio.cameras.rays", RaySamples=ns.RaySamples, Frustums=ns.Frustums, RayBundle=ns.RayBundle)
    _module("nerfstudio.cameras.cameras", Cameras=_Inert)

    @dataclass
    class CameraOptimizerConfig:
        mode: str = "off"

        def setup(self, num_cameras, device):
            return CameraOptimizer(self, num_cameras, device)

    class CameraOptimizer(nn.Module):  # mode "off": identity (headline configuration)
        def __init__(self, config, num_cameras, device):
            super().__init__()
            self.config = config

        def apply_to_raybundle(self, ray_bundle):
            return None

        def get_loss_dict(self, d):
            return None

        def get_metrics_dict(self, d):
            return None

        def get_param_groups(self, param_groups):
            return None

    _module("nerfstudio.cameras.camera_optimizers", CameraOptimizer=CameraOptimizer, CameraOptimizerConfig=CameraOptimizerConfig)
    _module("nerfstudio.data.scene_box", SceneBox=SceneBox, OrientedBox=_Inert)

    @dataclass
    class Semantics:
        filenames: list
        classes: list
        colors: torch.Tensor
        mask_classes: list = field(default_factory=list)

    _module("nerfstudio.data.dataparsers.base_dataparser", Semantics=Semantics)
    _module("nerfstudio.engine.callbacks", TrainingCallback=_Callback, TrainingCallbackAttributes=_Inert, TrainingCallbackLocation=_Loc)
    _module("nerfstudio.field_components.activations", trunc_exp=ns.trunc_exp)
    _module("nerfstudio.field_components.base_field_component", FieldComponent=FieldComponent)
    _module("nerfstudio.field_components.encodings", Encoding=nn.Module, Identity=nn.Identity, HashEncoding=HashEncoding, NeRFEncoding=NeRFEncoding,
            SHEncoding=SHEncoding)
    _module("nerfstudio.field_components.embedding", Embedding=ns.Embedding)
    _module("nerfstudio.field_components.field_heads", FieldHeadNames=FieldHeadNames, FieldHead=FieldHead, DensityFieldHead=_Inert,
            SemanticFieldHead=_Inert, RGBFieldHead=_Inert)
    _module("nerfstudio.field_components.mlp", MLP=MLP)
    _module("nerfstudio.field_components.spatial_distortions", SpatialDistortion=SpatialDistortion, SceneContraction=ns.SceneContraction)
    _module("nerfstudio.fields.base_field", Field=Field, get_normalized_directions=ns.get_normalized_directions)

    class HashMLPDensityField(ns.HashMLPDensityField):
        def __init__(self, aabb, implementation="torch", **kw):
            super().__init__(aabb, **kw)

    _module("nerfstudio.fields.density_fields", HashMLPDensityField=HashMLPDensityField)
    _module("nerfstudio.fields.semantic_nerf_field", SemanticNerfField=_Inert)
    _module("nerfstudio.model_components.losses", MSELoss=nn.MSELoss, distortion_loss=ns.distortion_loss, interlevel_loss=ns.interlevel_loss,
            scale_gradients_by_distance_squared=None)
    _module("nerfstudio.model_components.renderers", AccumulationRenderer=ns.AccumulationRenderer, DepthRenderer=ns.DepthRenderer,
            RGBRenderer=ns.RGBRenderer, SemanticRenderer=ns.SemanticRenderer, UncertaintyRenderer=_InertModule)
    _module("nerfstudio.model_components.ray_samplers", ProposalNetworkSampler=ns.ProposalNetworkSampler, UniformSampler=ns.UniformSampler,
            SpacedSampler=ns.SpacedSampler)
    _module("nerfstudio.model_components.scene_colliders", NearFarCollider=ns.NearFarCollider)

    class Model(nn.Module):
        """nerfstudio/models/base_model.py Model.__init__: stores config / scene_box / num_train_data / kwargs, then
        populate_modules()."""

        def __init__(self, config, scene_box, num_train_data, **kwargs):
            super().__init__()
            self.config = config
            self.scene_box = scene_box
            self.render_aabb = None
            self.num_train_data = num_train_data
            self.kwargs = kwargs
            self.collider = None
            self.populate_modules()
            self.device_indicator_param = nn.Parameter(torch.empty(0))

        @property
        def device(self):
            return self.device_indicator_param.device

        def populate_modules(self):
            return None

    _module("nerfstudio.models.base_model", Model=Model)

    @dataclass
    class NerfactoModelConfig:
        """nerfstudio/models/nerfacto.py NerfactoModelConfig defaults (1.1.3) restricted to what FruitModel reads
        (SURVEY.md section 8 preamble)."""

        _target: type = None
        near_plane: float = 0.05
        far_plane: float = 1000.0
        background_color: str = "last_sample"
        hidden_dim: int = 64
        hidden_dim_color: int = 64
        hidden_dim_transient: int = 64
        num_levels: int = 16
        base_res: int = 16
        max_res: int = 2048
        log2_hashmap_size: int = 19
        features_per_level: int = 2
        num_proposal_samples_per_ray: tuple = (256, 96)
        num_nerf_samples_per_ray: int = 48
        proposal_update_every: int = 5
        proposal_warmup: int = 5000
        num_proposal_iterations: int = 2
        use_same_proposal_network: bool = False
        proposal_net_args_list: list = field(default_factory=lambda: [
            {"hidden_dim": 16, "log2_hashmap_size": 17, "num_levels": 5, "max_res": 128, "use_linear": False},
            {"hidden_dim": 16, "log2_hashmap_size": 17, "num_levels": 5, "max_res": 256, "use_linear": False},
        ])
        proposal_initial_sampler: str = "piecewise"
        interlevel_loss_mult: float = 1.0
        distortion_loss_mult: float = 0.002
        use_proposal_weight_anneal: bool = True
        use_average_appearance_embedding: bool = True
        proposal_weights_anneal_slope: float = 10.0
        proposal_weights_anneal_max_num_iters: int = 1000
        use_single_jitter: bool = True
        predict_normals: bool = False
        disable_scene_contraction: bool = False
        use_gradient_scaling: bool = False
        implementation: str = "torch"
        appearance_embed_dim: int = 32
        camera_optimizer: CameraOptimizerConfig = field(default_factory=CameraOptimizerConfig)
        eval_num_rays_per_chunk: int = 1 << 15

    _module("nerfstudio.models.nerfacto", NerfactoModelConfig=NerfactoModelConfig)
    _module("nerfstudio.utils.colormaps")
    _module("nerfstudio.utils", colormaps=sys.modules["nerfstudio.utils.colormaps"])
    # ---- unrelated third-party imports of the reference files ---------------------------------------------------------
    if "torchmetrics" not in sys.modules:
        try:
            import torchmetrics  # noqa: F401
        except ImportError:
            _module("torchmetrics", PeakSignalNoiseRatio=_Psnr, JaccardIndex=_Inert)
            _module("torchmetrics.functional", structural_similarity_index_measure=None)
            _module("torchmetrics.image.lpip", LearnedPerceptualImagePatchSimilarity=_Inert)
    try:
        import nerfacc  # noqa: F401
    except ImportError:
        _module("nerfacc", OccGridEstimator=_Inert)
    if "segmentation" not in sys.modules:
        _module("segmentation.segmenter")
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)


def _install_product_shims() -> None:
    """nerfstudio's module names -> cropnerf_b200 classes (what a nerfstudio install would provide is replaced class for class by the
    product; INTEGRATION.md lists the same table).  Off-path names stay inert stubs."""
    from cropnerf_b200 import density_fields as p_df
    from cropnerf_b200 import field_components as p_fc
    from cropnerf_b200 import fruit_field as p_ff
    from cropnerf_b200 import fruit_nerf as p_fn
    from cropnerf_b200 import ray_samplers as p_rs
    from cropnerf_b200 import rays as p_rays
    from cropnerf_b200 import renderers as p_rd

    class SceneBox:
        def __init__(self, aabb):
            self.aabb = aabb

    @dataclass
    class Semantics:
        filenames: list
        classes: list
        colors: torch.Tensor
        mask_classes: list = field(default_factory=list)

    class Model(nn.Module):  # nerfstudio/models/base_model.py Model.__init__
        def __init__(self, config, scene_box, num_train_data, **kwargs):
            super().__init__()
            self.config, self.scene_box, self.render_aabb, self.num_train_data, self.kwargs = config, scene_box, None, num_train_data, kwargs
            self.collider = None
            self.populate_modules()
            self.device_indicator_param = nn.Parameter(torch.empty(0))

        @property
        def device(self):
            return self.device_indicator_param.device

        def populate_modules(self):
            return None

    @dataclass
    class NerfactoModelConfig:  # nerfacto.py defaults FruitModel reads (SURVEY.md section 8 preamble)
        _target: type = None
        near_plane: float = 0.05
        far_plane: float = 1000.0
        background_color: str = "last_sample"
        hidden_dim: int = 64
        hidden_dim_color: int = 64
        hidden_dim_transient: int = 64
        num_levels: int = 16
        base_res: int = 16
        max_res: int = 2048
        log2_hashmap_size: int = 19
        features_per_level: int = 2
        num_proposal_samples_per_ray: tuple = (256, 96)
        num_nerf_samples_per_ray: int = 48
        proposal_update_every: int = 5
        proposal_warmup: int = 5000
        num_proposal_iterations: int = 2
        use_same_proposal_network: bool = False
        proposal_net_args_list: list = field(default_factory=lambda: [
            {"hidden_dim": 16, "log2_hashmap_size": 17, "num_levels": 5, "max_res": 128, "use_linear": False},
            {"hidden_dim": 16, "log2_hashmap_size": 17, "num_levels": 5, "max_res": 256, "use_linear": False},
        ])
        proposal_initial_sampler: str = "piecewise"
        interlevel_loss_mult: float = 1.0
        distortion_loss_mult: float = 0.002
        use_proposal_weight_anneal: bool = True
        use_average_appearance_embedding: bool = True
        proposal_weights_anneal_slope: float = 10.0
        proposal_weights_anneal_max_num_iters: int = 1000
        use_single_jitter: bool = True
        predict_normals: bool = False
        disable_scene_contraction: bool = False
        use_gradient_scaling: bool = False
        implementation: str = "torch"
        appearance_embed_dim: int = 32
        camera_optimizer: object = field(default_factory=p_fn.CameraOptimizerConfig)
        eval_num_rays_per_chunk: int = 1 << 15

    _module("nerfstudio")
    _module("nerfstudio.cameras.rays", RaySamples=p_rays.RaySamples, Frustums=p_rays.Frustums, RayBundle=p_rays.RayBundle)
    _module("nerfstudio.cameras.cameras", Cameras=_Inert)
    _module("nerfstudio.cameras.camera_optimizers", CameraOptimizer=p_fn.CameraOptimizer, CameraOptimizerConfig=p_fn.CameraOptimizerConfig)
    _module("nerfstudio.data.scene_box", SceneBox=SceneBox, OrientedBox=_Inert)
    _module("nerfstudio.data.dataparsers.base_dataparser", Semantics=Semantics)
    _module("nerfstudio.engine.callbacks", TrainingCallback=_Callback, TrainingCallbackAttributes=_Inert, TrainingCallbackLocation=_Loc)
    _module("nerfstudio.field_components.activations", trunc_exp=None)
    _module("nerfstudio.field_components.base_field_component", FieldComponent=nn.Module)
    _module("nerfstudio.field_components.encodings", Encoding=nn.Module, Identity=nn.Identity, HashEncoding=p_fc.HashEncoding, NeRFEncoding=_InertModule,
            SHEncoding=p_fc.SHEncoding)
    _module("nerfstudio.field_components.embedding", Embedding=p_fc.Embedding)
    _module("nerfstudio.field_components.field_heads", FieldHeadNames=p_fc.FieldHeadNames, FieldHead=p_fc.FieldHead, DensityFieldHead=_Inert,
            SemanticFieldHead=p_fc.SemanticFieldHead, RGBFieldHead=_Inert)
    _module("nerfstudio.field_components.mlp", MLP=p_fc.MLP)
    _module("nerfstudio.field_components.spatial_distortions", SpatialDistortion=nn.Module, SceneContraction=p_fc.SceneContraction)
    _module("nerfstudio.fields.base_field", Field=nn.Module, get_normalized_directions=None)
    _module("nerfstudio.fields.density_fields", HashMLPDensityField=p_df.HashMLPDensityField)
    _module("nerfstudio.fields.semantic_nerf_field", SemanticNerfField=_Inert)
    _module("nerfstudio.model_components.losses", MSELoss=nn.MSELoss, distortion_loss=p_rd.distortion_loss, interlevel_loss=p_rd.interlevel_loss,
            scale_gradients_by_distance_squared=None)
    _module("nerfstudio.model_components.renderers", AccumulationRenderer=p_rd.AccumulationRenderer, DepthRenderer=p_rd.DepthRenderer,
            RGBRenderer=p_rd.RGBRenderer, SemanticRenderer=p_rd.SemanticRenderer, UncertaintyRenderer=_InertModule)
    _module("nerfstudio.model_components.ray_samplers", ProposalNetworkSampler=p_rs.ProposalNetworkSampler, UniformSampler=p_rs.UniformSampler,
            SpacedSampler=p_rs.SpacedSampler)
    _module("nerfstudio.model_components.scene_colliders", NearFarCollider=p_fn.NearFarCollider)
    _module("nerfstudio.models.base_model", Model=Model)
    _module("nerfstudio.models.nerfacto", NerfactoModelConfig=NerfactoModelConfig)
    _module("nerfstudio.utils.colormaps")
    _module("nerfstudio.utils", colormaps=sys.modules["nerfstudio.utils.colormaps"])
    if "torchmetrics" not in sys.modules:
        try:
            import torchmetrics  # noqa: F401
        except ImportError:
            _module("torchmetrics", PeakSignalNoiseRatio=_Psnr, JaccardIndex=_Inert)
            _module("torchmetrics.functional", structural_similarity_index_measure=None)
            _module("torchmetrics.image.lpip", LearnedPerceptualImagePatchSimilarity=_Inert)
    try:
        import nerfacc  # noqa: F401
    except ImportError:
        _module("nerfacc", OccGridEstimator=_Inert)
    if "segmentation" not in sys.modules:
        _module("segmentation.segmenter")
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    # the Field itself is replaced as a module (north_star: "the FruitField field ... replaced behind the same nerfstudio Field interface"):
    # `from fruit_nerf.fruit_field import FruitField, SemanticNeRFField` in the reference's fruit_nerf.py resolves to the product's class
    import fruit_nerf  # noqa: F401  (the reference's package, empty __init__)

    _module("fruit_nerf.fruit_field", FruitField=p_ff.FruitField, SemanticNeRFField=_Inert)


class _InertModule(nn.Module):
    def __init__(self, *a, **k):
        super().__init__()


class _Psnr(nn.Module):
    """torchmetrics.PeakSignalNoiseRatio(data_range=1.0) for one batch: 10 log10(1 / mse)."""

    def __init__(self, data_range=1.0):
        super().__init__()
        self.data_range = data_range

    def forward(self, preds, target):
        return 10.0 * torch.log10(self.data_range**2 / torch.mean((preds - target) ** 2))


@dataclass
class _Callback:
    where_to_run: list
    func: object
    update_every_num_iters: int = 1


class _Loc(enum.Enum):
    BEFORE_TRAIN_ITERATION = "BEFORE_TRAIN_ITERATION"
    AFTER_TRAIN_ITERATION = "AFTER_TRAIN_ITERATION"


# -------------------------------------------------------------------------------------------------------------------
def load_reference(backend: str = "oracle"):
    """-> (reference fruit_field module, reference fruit_nerf module), imported from /root/reference."""
    if not reference_available():
        raise RuntimeError("/root/reference is not present on this machine (GPU box): use the committed tests/golden/ref_*.npz")
    install_shims(backend)
    import fruit_nerf.fruit_field as ref_field  # noqa: E402
    import fruit_nerf.fruit_nerf as ref_model  # noqa: E402

    return ref_field, ref_model


def build_reference_model(cfg, num_images: int, state: Dict[str, torch.Tensor], test_mode: str = "val", backend: str = "oracle"):
    """The reference's FruitModel configured like the oracle's ``cfg`` (oracle/fruit_torch.FruitNerfModelConfig) and
    loaded with the same state dict."""
    _, ref_model = load_reference(backend)
    sem_cls = sys.modules["nerfstudio.data.dataparsers.base_dataparser"].Semantics
    rc = ref_model.FruitNerfModelConfig()
    for k in cfg.__dataclass_fields__:
        if hasattr(rc, k):
            setattr(rc, k, getattr(cfg, k))
    scene_box = sys.modules["nerfstudio.data.scene_box"].SceneBox(torch.tensor([[-1.0, -1.0, -1.0], [1.0, 1.0, 1.0]]))
    # cotton_nerf_dataparser.py:248-255: classes ["fruit"], colors [0, 1]
    semantics = sem_cls(filenames=[], classes=["fruit"], colors=torch.tensor([0.0, 1.0]), mask_classes=[])
    model = ref_model.FruitModel(rc, metadata={"semantics": semantics}, scene_box=scene_box, num_train_data=num_images, test_mode=test_mode)
    missing, unexpected = model.load_state_dict(state, strict=False)
    bad = [m for m in missing if "hash_table" in m or "weight" in m or "bias" in m]
    assert not bad, f"reference model did not receive: {bad}"
    return model


def run_reference_case(name: str, num_images: int = 20, seed: int = 0) -> Dict[str, np.ndarray]:
    """``oracle.cases.run_case`` with the reference's own FruitModel in place of the restated one."""
    from . import cases
    from cropnerf_b200 import synthetic

    spec = cases.CASES[name]
    cfg = cases.make_config(spec.get("cfg"))
    oracle, state = cases.build_oracle(cfg, num_images, seed, spec["table_scale"])
    model = build_reference_model(cfg, num_images, state)
    R = spec["num_rays"]
    rays = synthetic.make_rays(R, seed=1, num_cameras=num_images)
    model.train(spec["training"])
    if spec["training"]:
        feed = synthetic.JitterFeed(synthetic.make_jitter(R, 3, seed=2))
        model.proposal_sampler.initial_sampler.rand_fn = feed
        model.proposal_sampler.pdf_sampler.rand_fn = feed
        for cb in model.get_training_callbacks(None):
            if any(getattr(w, "value", w) == "BEFORE_TRAIN_ITERATION" for w in cb.where_to_run):
                cb.func(500)
    bundle = cases.oracle_bundle(rays, with_near_far=spec.get("near_far"))
    res: Dict[str, np.ndarray] = {}
    if spec["training"]:
        out = model(bundle)
        targets = synthetic.make_targets(R, seed=3)
        loss_dict = model.get_loss_dict(out, targets)
        metrics = model.get_metrics_dict(out, targets)
        for k, v in loss_dict.items():
            res["loss_" + k] = v.detach().numpy()
        res["metric_distortion"] = metrics["distortion"].detach().numpy()
        res["metric_psnr"] = metrics["psnr"].detach().numpy()
        sum(loss_dict.values()).backward()
        for pname, p in model.named_parameters():
            if p.grad is not None and p.numel() > 0:
                res["gradnorm/" + pname] = p.grad.double().norm().numpy()
    else:
        with torch.no_grad():
            out = model(bundle)
    for k in ("rgb", "accumulation", "depth", "prop_depth_0", "prop_depth_1", "semantics", "semantics_colormap"):
        res[k] = out[k].detach().numpy()
    return res


MODE_CASES = {
    # test_mode "inference": FruitField.get_inference_outputs + FruitModel.get_inference_outputs (fruit_nerf.py:497-541)
    "inference": dict(num_rays=64, test_mode="inference", near_far=None, samples=None),
    # test_mode "export" after setup_inference: the reference's UniformSamplerWithNoise + get_export_outputs (fruit_nerf.py:185-189,476-494)
    "export": dict(num_rays=32, test_mode="export", near_far=(0.0, 1.5), samples=64),
}


def run_mode_case(kind: str, which: str, num_images: int = 20, seed: int = 0) -> Dict[str, np.ndarray]:
    """``which`` = "reference" (the reference's FruitModel through the shims) or "oracle" (oracle/fruit_torch.py)."""
    from . import cases
    from cropnerf_b200 import synthetic

    spec = MODE_CASES[kind]
    cfg = cases.make_config(dict(log2_hashmap_size=14))
    oracle, state = cases.build_oracle(cfg, num_images, seed, 0.5)
    if which == "reference":
        model = build_reference_model(cfg, num_images, state, test_mode=spec["test_mode"])
    else:
        model = oracle
        model.test_mode = spec["test_mode"]
        model.field.test_mode = spec["test_mode"]
    if spec["samples"] is not None:
        model.setup_inference(True, spec["samples"])
    model.eval()
    rays = synthetic.make_rays(spec["num_rays"], seed=5, num_cameras=num_images)
    with torch.no_grad():
        out = model(cases.oracle_bundle(rays, with_near_far=spec["near_far"]))
    keys = ("rgb", "semantics", "semantics_colormap", "density", "point_location", "accumulation", "depth", "prop_depth_0", "prop_depth_1")
    return {k: out[k].detach().numpy() for k in keys if k in out}


VOLUME_BOXES = {
    # scripts/exporter.py:70-73 defaults; a box entirely above z = 0 (exercises the reference's sign()*|z_min| + |z_max| ray length);
    # an anisotropic one (grid counts int(dx/dz*n), int(dy/dz*n))
    "exporter_default": ((-1.0, -1.0, -1.0 + 0.318), (1.0, 1.0, 1.0 + 0.318)),
    "positive_z": ((-0.5, -0.25, 0.125), (0.5, 0.75, 0.625)),
    "anisotropic": ((-0.3, -0.1, -0.4), (0.6, 0.2, 0.1)),
}


def run_volume_rays_case(n: int = 13, batch: int = 64) -> Dict[str, np.ndarray]:
    """The reference's volumetric-export ray source executed verbatim: ``get_corners_of_aabb`` + ``sample_surface_points``
    (data/fruit_datamanager.py:42-120; extracted with ``ast`` because that file imports the nerfstudio data stack) feeding
    ``OrthographicRayGenerator`` (components/ray_generators.py, imported through the shims), all batches concatenated."""
    import ast

    if not reference_available():
        raise RuntimeError("/root/reference is not present on this machine")
    install_shims()
    src_path = os.path.join(REFERENCE_ROOT, "fruit_nerf", "data", "fruit_datamanager.py")
    tree = ast.parse(open(src_path).read())
    wanted = [node for node in tree.body if isinstance(node, ast.FunctionDef) and node.name in ("get_corners_of_aabb", "sample_surface_points")]
    ns_exec: Dict[str, object] = {"torch": torch, "np": np}
    exec(compile(ast.Module(body=wanted, type_ignores=[]), src_path, "exec"), ns_exec)
    from fruit_nerf.components.ray_generators import OrthographicRayGenerator  # noqa: E402

    out: Dict[str, np.ndarray] = {}
    for name, box in VOLUME_BOXES.items():
        corners = ns_exec["get_corners_of_aabb"](aabb=box, device="cpu")
        pts, plane = ns_exec["sample_surface_points"](corners, n=n, device="cpu", noise=False)
        gen = OrthographicRayGenerator(surface_points=pts, plane_normal=plane, ray_batch_size=batch, device="cpu", aabb=box)
        o, d, f, count = [], [], [], 0
        while sum(x.shape[0] for x in o) < pts.shape[0]:
            count += 1
            rb = gen(count)
            o.append(rb.origins); d.append(rb.directions); f.append(rb.fars)
        out[name + "_origins"] = torch.cat(o).numpy()
        out[name + "_directions"] = torch.cat(d).numpy()
        out[name + "_fars"] = torch.cat(f).numpy()
        out[name + "_aabb"] = np.asarray(box, dtype=np.float32)
    out["n"] = np.asarray(n)
    return out


def density_normalisation_inputs() -> Dict[str, torch.Tensor]:
    """Points inside, on the faces of, and far outside the scene box (contracted and AABB-normalised variants)."""
    g = torch.Generator().manual_seed(7)
    pts = torch.cat([torch.rand((64, 3), generator=g) * 2 - 1, (torch.rand((64, 3), generator=g) * 2 - 1) * 6,
                     torch.tensor([[1.0, 0.0, 0.0], [-1.0, -1.0, -1.0], [0.0, 0.0, 0.0], [2.0, 2.0, 2.0], [1e3, -5.0, 0.5], [0.999999, 0.5, -0.25]])])
    return {"points": pts, "aabb": torch.tensor([[-1.0, -1.0, -1.0], [1.0, 1.0, 1.0]])}


def run_density_normalisation_case() -> Dict[str, np.ndarray]:
    """``normalize_point_coords`` of the reference (bayesrays/utils.py:6-16: "coordinate normalization process according to
    density_feild.py in nerfstudio") executed verbatim -- extracted with ``ast`` because the file imports nerfstudio.utils.math --
    with the scene contraction / SceneBox of the shims.  Pins the normalisation + selector wiring of the restated HashMLPDensityField."""
    import ast

    if not reference_available():
        raise RuntimeError("/root/reference is not present on this machine")
    install_shims()
    from . import nerfstudio_torch as ns

    src_path = os.path.join(REFERENCE_ROOT, "fruit_nerf", "bayesrays", "utils.py")
    tree = ast.parse(open(src_path).read())
    wanted = [node for node in tree.body if isinstance(node, ast.FunctionDef) and node.name == "normalize_point_coords"]
    ns_exec: Dict[str, object] = {"torch": torch, "SceneBox": sys.modules["nerfstudio.data.scene_box"].SceneBox}
    exec(compile(ast.Module(body=wanted, type_ignores=[]), src_path, "exec"), ns_exec)
    inp = density_normalisation_inputs()
    out: Dict[str, np.ndarray] = {"points": inp["points"].numpy(), "aabb": inp["aabb"].numpy()}
    for name, distortion in (("contract", ns.SceneContraction()), ("aabb", None)):
        pos, sel = ns_exec["normalize_point_coords"](inp["points"], inp["aabb"], distortion)
        out[name + "_pos"] = pos.numpy()
        out[name + "_selector"] = sel.numpy()
    return out


def main() -> None:
    from . import cases

    out_dir = os.path.join(cases.ROOT, "tests", "golden")
    for name in cases.CASES:
        res = run_reference_case(name)
        path = os.path.join(out_dir, "ref_" + name + ".npz")
        np.savez_compressed(path, **res)
        print("reference-executed", name, "->", path)
    for kind in MODE_CASES:
        res = run_mode_case(kind, "reference")
        path = os.path.join(out_dir, "ref_mode_" + kind + ".npz")
        np.savez_compressed(path, **res)
        print("reference-executed", kind, "->", path, {k: v.shape for k, v in res.items()})
    res = run_volume_rays_case()
    path = os.path.join(out_dir, "ref_volume_rays.npz")
    np.savez_compressed(path, **res)
    print("reference-executed volume rays ->", path, {k: v.shape for k, v in res.items()})
    res = run_density_normalisation_case()
    path = os.path.join(out_dir, "ref_density_normalisation.npz")
    np.savez_compressed(path, **res)
    print("reference-executed density-field normalisation ->", path, {k: v.shape for k, v in res.items()})


if __name__ == "__main__":
    main()
