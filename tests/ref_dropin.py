"""Drop-in demonstration (script, run by tests/test_dropin_cpu.py in a fresh process): the reference's UNMODIFIED ``fruit_nerf.py``
(``/root/reference/crop_nerf/fruit_nerf/fruit_nerf.py``: FruitNerfModelConfig, FruitModel.populate_modules / get_param_groups /
get_training_callbacks / setup_inference / get_outputs / get_loss_dict / get_metrics_dict, :59-645) is imported from where it lies with
every ``nerfstudio.*`` name it imports bound to the cropnerf_b200 class of the same name and ``fruit_nerf.fruit_field.FruitField`` bound to
the product's FruitField (``oracle.ref_shim.install_shims("product")``) -- i.e. exactly what a plugin install would do.

Without a GPU (this container) it checks everything up to the first kernel launch: the reference's ``populate_modules`` builds B200
modules, their state dict takes the reference-named tensors, param groups / callbacks / ``setup_inference`` work, and ``forward`` fails
loudly (no CPU fallback).  Where the reference tree AND a CUDA device are both present it also runs the reference's
``get_outputs`` -> ``get_loss_dict`` -> ``get_metrics_dict`` -> backward on the kernels and compares with the oracle (the driver's GPU boxes
have no /root/reference, so that branch is for a maintainer's machine).  Prints one JSON object."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from cropnerf_b200 import density_fields, fruit_field, ray_samplers, renderers, synthetic  # noqa: E402
from cropnerf_b200.fruit_nerf import FruitModel as ProductModel  # noqa: E402
from cropnerf_b200.fruit_nerf import FruitNerfModelConfig as ProductConfig  # noqa: E402
from cropnerf_b200.rays import RayBundle  # noqa: E402
from oracle import cases, ref_shim  # noqa: E402


def main() -> None:
    precision = os.environ.get("CNB_PRECISION", "fp32")
    num_images, R = 12, 128
    cfg = cases.make_config(dict(log2_hashmap_size=12))
    oracle, state = cases.build_oracle(cfg, num_images, seed=0, table_scale=0.5)
    model = ref_shim.build_reference_model(cfg, num_images, state, backend="product")
    ref_mod = sys.modules[type(model).__module__]
    rep = {"model_class": f"{type(model).__module__}.{type(model).__qualname__}", "model_file": ref_mod.__file__, "precision": precision}
    assert ref_mod.__file__.startswith(ref_shim.REFERENCE_ROOT), ref_mod.__file__
    # ---- the reference's populate_modules built product modules ------------------------------------------------------------------
    assert type(model.field) is fruit_field.FruitField
    assert all(type(n) is density_fields.HashMLPDensityField for n in model.proposal_networks) and len(model.proposal_networks) == 2
    assert type(model.proposal_sampler) is ray_samplers.ProposalNetworkSampler
    assert type(model.renderer_rgb) is renderers.RGBRenderer and type(model.renderer_depth) is renderers.DepthRenderer
    assert type(model.renderer_accumulation) is renderers.AccumulationRenderer and type(model.renderer_semantics) is renderers.SemanticRenderer
    assert [fn.__self__ for fn in model.density_fns] == list(model.proposal_networks)
    # ---- state dict: same learnable tensors, names and shapes as the product's own FruitModel and as the reference-named state --------
    kw = {k: getattr(cfg, k) for k in cfg.__dataclass_fields__ if k in ProductConfig.__dataclass_fields__}
    own = ProductModel(ProductConfig(**kw), num_train_data=num_images)
    sd, sd_own = model.state_dict(), own.state_dict()
    learn = {n: tuple(p.shape) for n, p in model.named_parameters() if p.numel()}
    learn_own = {n: tuple(p.shape) for n, p in own.named_parameters()}
    assert learn == learn_own, sorted(set(learn) ^ set(learn_own))
    for k, v in state.items():
        if k in sd:
            assert torch.equal(sd[k], v), k
    rep["parameters"] = len(learn)
    rep["state_keys_shared_with_reference_state"] = len([k for k in state if k in sd])
    # ---- param groups / callbacks / setup_inference (fruit_nerf.py:185-232) ----------------------------------------------------------
    groups = model.get_param_groups()
    assert set(groups) == {"proposal_networks", "fields"} and sum(p.numel() for p in groups["fields"]) == sum(p.numel() for p in own.get_param_groups()["fields"])
    cbs = model.get_training_callbacks(None)
    assert len(cbs) == 2
    cbs[0].func(500)
    own.get_training_callbacks()[0].func(500)
    assert model.proposal_sampler._anneal == own.proposal_sampler._anneal and 0 < model.proposal_sampler._anneal < 1
    cbs[1].func(500)
    assert model.proposal_sampler._step == 500
    rep["anneal_at_500"] = float(model.proposal_sampler._anneal)
    # ---- forward: into the kernels, or loudly nowhere ------------------------------------------------------------------------------
    rays = synthetic.make_rays(R, seed=1, num_cameras=num_images)
    targets = synthetic.make_targets(R, seed=3)
    if not torch.cuda.is_available():
        try:
            model(RayBundle(rays["origins"], rays["directions"], rays["pixel_area"], rays["camera_indices"]))
            raise AssertionError("forward on CPU tensors must fail: there is no CPU fallback")
        except RuntimeError as e:
            assert "CUDA" in str(e) or "cuda" in str(e), e
            rep["cpu_forward"] = "raises: " + str(e)[:80]
    else:
        dev = torch.device("cuda:0")
        for net in [model.field, *model.proposal_networks]:
            net.precision = precision
        model = model.to(dev).train()
        oracle.train()
        jit = synthetic.make_jitter(R, 3, seed=2)
        for m in (model, oracle):
            feed = synthetic.JitterFeed(jit)
            m.proposal_sampler.initial_sampler.rand_fn = feed
            m.proposal_sampler.pdf_sampler.rand_fn = feed
        oracle.set_anneal(500)
        out = model(RayBundle(rays["origins"].to(dev), rays["directions"].to(dev), rays["pixel_area"].to(dev), rays["camera_indices"].to(dev)))
        tg = {k: v.to(dev) for k, v in targets.items()}
        loss = model.get_loss_dict(out, tg)
        metrics = model.get_metrics_dict(out, tg)
        sum(loss.values()).backward()
        ref_out = oracle(cases.oracle_bundle(rays))
        ref_loss = oracle.get_loss_dict(ref_out, targets)
        ref_metrics = oracle.get_metrics_dict(ref_out, targets)
        sum(ref_loss.values()).backward()
        tol = 1e-4 if precision == "fp32" else 2e-3
        rep["losses"] = {}
        for k in ref_loss:
            a, b = float(loss[k].detach()), float(ref_loss[k].detach())
            rep["losses"][k] = [a, b]
            assert abs(a - b) <= tol * abs(b) + 1e-7, (k, a, b)
        assert abs(float(metrics["psnr"]) - float(ref_metrics["psnr"])) <= 1e-2
        gtol = 2e-3 if precision == "fp32" else 4e-2
        ref_params = dict(oracle.named_parameters())
        for n, p in model.named_parameters():
            if n in ref_params and ref_params[n].grad is not None and p.grad is not None:
                gr = ref_params[n].grad.double()
                err = (p.grad.cpu().double() - gr).norm().item() / (gr.norm().item() + 1e-30)
                assert err <= gtol, (n, err)
        rep["gpu_forward"] = "losses and gradients match the oracle"
    # ---- setup_inference installs the reference's own UniformSamplerWithNoise (a subclass of the product's SpacedSampler) -------------
    model.setup_inference(True, 64)
    assert type(model.proposal_sampler).__module__ == "fruit_nerf.components.ray_samplers" and isinstance(model.proposal_sampler, ray_samplers.SpacedSampler)
    assert model.field.spatial_distortion is None
    rep["ok"] = True
    print(json.dumps(rep))


if __name__ == "__main__":
    main()
