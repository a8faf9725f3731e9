"""The one-call pipeline (csrc/pipeline.cu: cnb_render_rays / cnb_train_step) against the per-module operator path and
against the oracle's golden vectors: same kernels, so per-ray outputs must agree to the last bit and gradients to fp32
atomic-ordering noise."""
import os

import numpy as np
import pytest
import torch

from helpers import GOLDEN, assert_close, product_bundle, product_model

from cropnerf_b200 import engine, synthetic
from oracle import cases

pytestmark = pytest.mark.gpu


def _models(dev, training, precision="fp32", num_images=20, small=True):
    cfg = cases.make_config({}, small=small)
    oracle, state = cases.build_oracle(cfg, num_images, 0, 0.5)
    a = product_model(cfg, state, num_images, dev, training, precision=precision)
    b = product_model(cfg, state, num_images, dev, training, precision=precision)
    return a, b


@pytest.mark.parametrize("precision", ["fp32", "mixed"])
def test_fused_render_equals_modular(dev, precision):
    fused, modular = _models(dev, False, precision, small=(precision == "fp32"))
    modular.fused_render = False
    rays = synthetic.make_rays(777, seed=4, num_cameras=20)
    with torch.no_grad():
        a = fused(product_bundle(rays, dev))
        b = modular(product_bundle(rays, dev))
    for k in ("rgb", "accumulation", "depth", "semantics", "prop_depth_0", "prop_depth_1", "semantics_colormap"):
        assert torch.equal(a[k], b[k]), f"{precision} {k}: max diff {(a[k] - b[k]).abs().max().item()}"


def test_fused_render_against_golden(dev):
    name = "tiny_eval"
    ref = dict(np.load(os.path.join(GOLDEN, name + ".npz")))
    spec = cases.CASES[name]
    cfg = cases.make_config(spec.get("cfg"))
    oracle, state = cases.build_oracle(cfg, 20, 0, spec["table_scale"])
    model = product_model(cfg, state, 20, dev, False)
    model.proposal_sampler.pdf_sampler.keep_inds = True
    rays = synthetic.make_rays(spec["num_rays"], seed=1, num_cameras=20)
    with torch.no_grad():
        out = model(product_bundle(rays, dev))
    assert_close(out["rgb"], ref["rgb"], 1e-4, "rgb")
    assert_close(out["accumulation"], ref["accumulation"], 1e-4, "accumulation")
    assert_close(out["depth"], ref["depth"], 1e-4, "depth", frac=0.999)
    inds = model.proposal_sampler.pdf_sampler.last_inds.cpu().numpy()
    assert inds.shape == ref["pdf_inds_last"].shape
    assert (inds == ref["pdf_inds_last"]).mean() >= 0.9995
    assert (out["semantics_colormap"].cpu().numpy() == ref["semantics_colormap"]).mean() >= 0.999


def test_near_far_injection_and_black_background(dev):
    """get_outputs_for_projections feeds AABB near/far and a black background override (fruit_nerf.py:283,307-308;
    scripts/semantic_projection.py:169): both reach the fused call."""
    from cropnerf_b200.renderers import background_color_override_context

    fused, modular = _models(dev, False)
    modular.fused_render = False
    rays = synthetic.make_rays(300, seed=9, num_cameras=20)
    with torch.no_grad(), background_color_override_context(torch.zeros(3)):
        a = fused(product_bundle(rays, dev, near_far=(0.3, 1.7)))
        b = modular(product_bundle(rays, dev, near_far=(0.3, 1.7)))
    for k in ("rgb", "accumulation", "depth", "semantics"):
        assert torch.equal(a[k], b[k]), k


@pytest.mark.parametrize("precision", ["fp32", "mixed"])
def test_fused_train_step_equals_modular(dev, precision):
    """Two identical models, same rays / targets / jitter: one step through cnb_train_step, one through the per-module
    autograd operators.  Losses, parameters after Adam and the (pre-Adam) gradients must agree."""
    R = 256
    fused, modular = _models(dev, True, precision, small=(precision == "fp32"))
    rays = synthetic.make_rays(R, seed=6, num_cameras=20)
    targets = {k: v.to(dev) for k, v in synthetic.make_targets(R, seed=3).items()}
    jit = synthetic.make_jitter(R, 3, seed=2)
    results = []
    for model, use_fused in ((fused, True), (modular, False)):
        feed = synthetic.JitterFeed(jit)
        model.proposal_sampler.initial_sampler.rand_fn = feed
        model.proposal_sampler.pdf_sampler.rand_fn = feed
        tr = engine.Trainer(model, fused=use_fused)
        grads = {}
        orig = tr.optimizer_step

        def spy(step, tr=tr, grads=grads, orig=orig, **kw):
            for name, g in tr.groups.items():
                grads[name] = g.grad.clone()
            orig(step, **kw)

        tr.optimizer_step = spy
        stats = tr.train_iteration(500, product_bundle(rays, dev), targets)
        results.append((stats, grads, {n: g.flat.clone() for n, g in tr.groups.items()}))
    (sa, ga, pa), (sb, gb, pb) = results
    for k in ("rgb_loss", "semantics_loss", "interlevel_loss", "distortion", "psnr"):
        a, b = float(sa[k]), float(sb[k])
        assert abs(a - b) <= 1e-5 * abs(b) + 1e-7, f"{k}: fused {a} vs modular {b}"
    tol = 1e-4 if precision == "fp32" else 2e-3
    for name in ga:
        scale = gb[name].abs().max().item() + 1e-20
        err = (ga[name] - gb[name]).abs().max().item() / scale
        assert err < tol, f"{precision} grad group {name}: {err:.3e}"
        assert gb[name].abs().max().item() > 0
    for name in pa:
        assert (pa[name] - pb[name]).abs().max().item() < 2e-2 + 0.0, name  # Adam steps of lr=1e-2 on agreeing gradients


def test_adam_zero_matches_torch_adam(dev):
    from cropnerf_b200 import ops

    g = torch.Generator().manual_seed(0)
    n = 4096 * 3
    p0 = torch.randn((n,), generator=g)
    ref = torch.nn.Parameter(p0.clone())
    opt = torch.optim.Adam([ref], lr=1e-2, eps=1e-15)
    p = p0.clone().to(dev)
    m, v = torch.zeros_like(p), torch.zeros_like(p)
    for step in range(1, 4):
        grad = torch.randn((n,), generator=g) * 0.1
        ref.grad = grad.clone()
        opt.step()
        gd = grad.to(dev)
        ops.adam_step(p, gd, m, v, 1e-2, step, zero_grad=True)
        assert float(gd.abs().max()) == 0.0
    assert_close(p, ref.detach(), 1e-5, "adam params")


@pytest.mark.parametrize("optimizer", ["adam", "radam"])
def test_cuda_graph_step_equals_eager(dev, optimizer):
    """The captured-graph training step replays exactly the kernels of the eager one-call step.  radam: the optimizer of the big / huge presets
    (engine.BIG_PRESET_OPTIMIZERS), whose update is the same kernel with other per-step scalars -- in-graph from the device scalar slots, eagerly
    through cnb_adam_step_zero_dev."""
    R = 256
    optimizers = engine.BIG_PRESET_OPTIMIZERS if optimizer == "radam" else None
    a, b = _models(dev, True, "mixed", small=False)
    rays = synthetic.make_rays(R, seed=6, num_cameras=20)
    targets = {k: v.to(dev) for k, v in synthetic.make_targets(R, seed=3).items()}
    jit = synthetic.make_jitter(R, 3, seed=2)
    stats, params = [], []
    for model, graph in ((a, True), (b, False)):
        feed = synthetic.JitterFeed(jit)
        model.proposal_sampler.initial_sampler.rand_fn = feed
        model.proposal_sampler.pdf_sampler.rand_fn = feed
        tr = engine.Trainer(model, optimizers=optimizers, cuda_graph=graph, force_proposal_update=True)
        for step in (3000, 3001, 3002):
            feed.reset()
            st = tr.train_iteration(step, product_bundle(rays, dev), targets)
        stats.append({k: float(v) for k, v in st.items()})
        params.append({n: g.flat.clone() for n, g in tr.groups.items()})
        if graph:
            assert len(tr._graphs) == 1
    for k in ("rgb_loss", "semantics_loss", "interlevel_loss", "distortion"):
        assert abs(stats[0][k] - stats[1][k]) <= 2e-4 * abs(stats[1][k]) + 1e-7, (k, stats[0][k], stats[1][k])
    for n in params[0]:
        assert (params[0][n] - params[1][n]).abs().max().item() < 5e-2, n


def test_grad_scaler_scaled_step_and_skip_on_inf(dev):
    """GradScaler row of f2: a power-of-two loss scale leaves the update unchanged (the scale is folded out inside Adam); an
    inf in any gradient skips the whole step on the device (parameters and moments untouched, gradients cleared, Adam's step
    count not advanced) and halves the scale, like torch.amp.GradScaler.step / update."""
    from cropnerf_b200 import ops

    R = 256
    a, b = _models(dev, True, "mixed", small=False)
    rays = synthetic.make_rays(R, seed=6, num_cameras=20)
    targets = {k: v.to(dev) for k, v in synthetic.make_targets(R, seed=3).items()}
    jit = synthetic.make_jitter(R, 3, seed=2)
    trainers = []
    for model, scaler in ((a, engine.GradScaler(init_scale=1024.0, growth_interval=2)), (b, None)):
        feed = synthetic.JitterFeed(jit)
        model.proposal_sampler.initial_sampler.rand_fn = feed
        model.proposal_sampler.pdf_sampler.rand_fn = feed
        tr = engine.Trainer(model, force_proposal_update=True, grad_scaler=scaler)
        for step in (3000, 3001):
            feed.reset()
            tr.train_iteration(step, product_bundle(rays, dev), targets)
        trainers.append((tr, feed))
    (ta, fa), (tb, _) = trainers
    assert ta.opt_step == tb.opt_step == 2
    assert ta.grad_scaler.scale == 2048.0 and ta.grad_scaler.skipped_steps == 0  # grew after growth_interval clean steps
    for n in ta.groups:
        # gradients differ only by the order of the fp32 atomics; Adam normalises them, so compare loosely
        assert (ta.groups[n].flat - tb.groups[n].flat).abs().max().item() < 5e-2, n
        assert float(ta.groups[n].grad.abs().max()) == 0.0
    # ---- poison one gradient element: the next step must be skipped everywhere
    before = {n: (g.flat.clone(), g.exp_avg.clone(), g.exp_avg_sq.clone()) for n, g in ta.groups.items()}
    ta.groups["proposal_networks"].grad[12345] = float("inf")
    ta._grads_clean = True  # keep the poisoned value (train_iteration would otherwise clear a "dirty" gradient first)
    fa.reset()
    ta.train_iteration(3002, product_bundle(rays, dev), targets)
    assert ta.opt_step == 2 and ta.grad_scaler.skipped_steps == 1 and ta.grad_scaler.scale == 1024.0
    for n, g in ta.groups.items():
        assert torch.equal(g.flat, before[n][0]) and torch.equal(g.exp_avg, before[n][1]) and torch.equal(g.exp_avg_sq, before[n][2]), n
        assert float(g.grad.abs().max()) == 0.0
    # ---- and the step after that runs normally again
    fa.reset()
    ta.train_iteration(3003, product_bundle(rays, dev), targets)
    assert ta.opt_step == 3 and not torch.equal(ta.groups["fields"].flat, before["fields"][0])
    # the check kernel itself, on a ragged length with the bad value in the tail
    flag = torch.zeros((1,), device=dev, dtype=torch.int32)
    gbuf = torch.zeros((1027 + 1,), device=dev)[:1027]
    ops.grad_check_finite(gbuf, flag)
    assert int(flag.item()) == 0
    gbuf[1026] = float("nan")
    ops.grad_check_finite(gbuf, flag)
    assert int(flag.item()) == 1
