"""Kernel-level parity: every C-ABI operator against the oracle's restatement of the same nerfstudio primitive,
on identical seeded inputs.  Bit-exact for integer/index work, tolerances written per test for floating point."""
import numpy as np
import pytest
import torch

from helpers import assert_close, assert_close_scaled, rel_err

from cropnerf_b200 import _lib as L
from cropnerf_b200 import ops
from cropnerf_b200.field_components import MLP as BMLP
from cropnerf_b200.field_components import HashEncoding as BHash
from oracle import nerfstudio_torch as ns

pytestmark = pytest.mark.gpu


def _positions(n, seed=0, edge_cases=True):
    g = torch.Generator().manual_seed(seed)
    p = torch.rand((n, 3), generator=g)
    if edge_cases:
        p[0] = 0.0  # masked samples land exactly here (App. B-3)
        p[1] = torch.tensor([0.5, 0.25, 0.125])  # exact grid nodes on power-of-two levels: ceil == floor
        p[2] = torch.tensor([1.0 - 2**-24, 1.0 - 2**-24, 1.0 - 2**-24])
        p[3] = torch.tensor([2**-20, 0.999, 0.5])
    return p


@pytest.mark.parametrize("levels,min_res,max_res,log2T", [(16, 16, 2048, 19), (5, 16, 128, 17), (5, 16, 256, 17), (7, 16, 2048, 13), (1, 16, 16, 8)])
def test_hashgrid_indices_and_features_bit_exact(dev, levels, min_res, max_res, log2T):
    torch.manual_seed(0)
    ref = ns.HashEncoding(levels, min_res, max_res, log2T, hash_init_scale=0.5)
    mine = BHash(levels, min_res, max_res, log2T)
    mine.load_state_dict(ref.state_dict())
    mine = mine.to(dev)
    assert torch.equal(mine.scalings.cpu(), ref.scalings)
    if levels == 16:
        assert ref.scalings[-1].item() == 2047.0  # App. B-2
    pos = _positions(20000, seed=levels)
    idx_ref, _ = ref.corner_indices(pos)
    feat_ref = ref(pos)
    feat, idx = ops.hashgrid_indices(pos.to(dev), mine.hash_table, levels, log2T, mine._scalings_host)
    assert torch.equal(idx.cpu().to(torch.int64), idx_ref), "hash indices must be bit-exact"
    # the blend is evaluated with the reference's operation order and separately rounded ops: bit-exact too
    assert torch.equal(feat.cpu(), feat_ref.detach()), f"max diff {(feat.cpu()-feat_ref).abs().max()}"


def test_hashgrid_backward(dev):
    torch.manual_seed(1)
    ref = ns.HashEncoding(16, 16, 2048, 15, hash_init_scale=0.5)
    mine = BHash(16, 16, 2048, 15)
    mine.load_state_dict(ref.state_dict())
    mine = mine.to(dev)
    pos = _positions(4096, seed=5)
    g = torch.Generator().manual_seed(2)
    dout = torch.randn((4096, 32), generator=g)
    ref(pos).backward(dout)
    mine(pos.to(dev)).backward(dout.to(dev))
    # scatter-add order differs (atomics): fp32 sums of <= a few thousand terms
    gr = ref.hash_table.grad
    err = (mine.hash_table.grad.cpu() - gr).abs().max().item() / gr.abs().max().item()
    assert err < 1e-5, err
    nz = ref.hash_table.grad.abs().sum(-1) > 0
    assert torch.equal((mine.hash_table.grad.cpu().abs().sum(-1) > 0), nz), "touched rows must coincide"


def test_hashgrid_empty_and_errors(dev):
    mine = BHash(5, 16, 128, 10).to(dev)
    out = mine(torch.empty((0, 3), device=dev))
    assert out.shape == (0, 10)
    with pytest.raises(RuntimeError):
        mine.to("cpu")(torch.rand(4, 3))  # no CPU fallback


@pytest.mark.parametrize("dims,act", [((32, 64, 16), None), ((15, 64, 64), None), ((64, 1), None), ((63, 64, 64, 3), "sigmoid"), ((10, 16, 1), None),
                                      # the big / huge presets' layers (above 64 wide: csrc/mlp_wide.cu, layer by layer)
                                      ((30, 128, 128, 128), None), ((128, 1), None), ((78, 64, 64, 3), "sigmoid"), ((32, 64, 31), None)])
def test_mlp_forward_backward(dev, dims, act):
    torch.manual_seed(3)
    nl = len(dims) - 1
    out_act = torch.nn.Sigmoid() if act == "sigmoid" else None
    ref = ns.MLP(dims[0], nl, dims[1] if nl > 1 else dims[-1], dims[-1], out_activation=out_act)
    mine = BMLP(dims[0], nl, dims[1] if nl > 1 else dims[-1], dims[-1], out_activation=out_act)
    mine.load_state_dict(ref.state_dict())
    mine = mine.to(dev)
    for n in (1, 127, 128, 1000):
        g = torch.Generator().manual_seed(n)
        x = torch.randn((n, dims[0]), generator=g)
        dy = torch.randn((n, dims[-1]), generator=g)
        xr = x.clone().requires_grad_(True)
        xm = x.to(dev).requires_grad_(True)
        ref.zero_grad()
        mine.zero_grad()
        yr = ref(xr)
        ym = mine(xm)
        # dot products of O(1) terms: compare against the output scale (fp32 FMA order differs from oneDNN's)
        assert_close_scaled(ym, yr, 1e-5, f"mlp{dims} fwd n={n}")
        yr.backward(dy)
        ym.backward(dy.to(dev))
        assert_close_scaled(xm.grad, xr.grad, 1e-5, f"mlp{dims} dx n={n}")
        for (k, pr), (_, pm) in zip(ref.named_parameters(), mine.named_parameters()):
            scale = pr.grad.abs().max().item() + 1e-12
            err = (pm.grad.cpu() - pr.grad).abs().max().item() / scale
            assert err < 2e-5, f"mlp{dims} grad {k} n={n}: {err}"


def _ray_case(R, S, seed, near=0.05, far=1000.0):
    from cropnerf_b200 import synthetic

    rays = synthetic.make_rays(R, seed=seed, num_cameras=20)
    nears = torch.full((R, 1), near)
    fars = torch.full((R, 1), far)
    return rays, nears, fars


@pytest.mark.parametrize("kind", ["piecewise", "uniform"])
@pytest.mark.parametrize("jitter", [None, "single", "full"])
def test_spaced_sampler_bit_exact(dev, kind, jitter):
    R, S = 257, 256
    rays, nears, fars = _ray_case(R, S, 1)
    ref = (ns.UniformLinDispPiecewiseSampler if kind == "piecewise" else ns.UniformSampler)(num_samples=S, single_jitter=(jitter == "single"))
    ref.train(jitter is not None)
    g = torch.Generator().manual_seed(9)
    t_rand = None
    if jitter is not None:
        t_rand = torch.rand((R, 1) if jitter == "single" else (R, S + 1), generator=g)
        ref.rand_fn = lambda shape, dtype=None, device=None: t_rand
    rb = ns.RayBundle(rays["origins"], rays["directions"], rays["pixel_area"], rays["camera_indices"], nears, fars)
    rs = ref(rb)
    sp_ref = torch.cat([rs.spacing_starts[..., 0], rs.spacing_ends[..., -1:, 0]], -1)
    eu_ref = torch.cat([rs.frustums.starts[..., 0], rs.frustums.ends[..., -1:, 0]], -1)
    sp, eu = ops.sample_spaced(nears.to(dev), fars.to(dev), torch.linspace(0.0, 1.0, S + 1).to(dev), None if t_rand is None else t_rand.to(dev),
                               L.SPACING_LINDISP_PIECEWISE if kind == "piecewise" else L.SPACING_UNIFORM)
    assert torch.equal(sp.cpu(), sp_ref.expand(R, S + 1)), (sp.cpu() - sp_ref).abs().max()
    assert torch.equal(eu.cpu(), eu_ref), (eu.cpu() - eu_ref).abs().max()


@pytest.mark.parametrize("training", [False, True])
@pytest.mark.parametrize("Sp,S,anneal", [(256, 96, 1.0), (96, 48, 0.37), (256, 96, 0.0), (33, 64, 1.0), (512, 512, 0.5)])
def test_pdf_sampler(dev, training, Sp, S, anneal):
    R = 300
    rays, nears, fars = _ray_case(R, Sp, 2)
    g = torch.Generator().manual_seed(Sp + S)
    weights = torch.rand((R, Sp, 1), generator=g) ** 4
    weights[0] = 0.0          # all-zero histogram: the eps padding branch
    weights[1, : Sp // 2] = 0.0
    weights[2] = 1e-9
    initial = ns.UniformLinDispPiecewiseSampler(num_samples=Sp, single_jitter=True)
    initial.eval()
    rb = ns.RayBundle(rays["origins"], rays["directions"], rays["pixel_area"], rays["camera_indices"], nears, fars)
    rs0 = initial(rb)
    pdf = ns.PDFSampler(include_original=False, single_jitter=True)
    pdf.train(training)
    rand = torch.rand((R, 1), generator=g)
    pdf.rand_fn = lambda shape, device=None, dtype=None: rand
    rs1 = pdf(rb, rs0, torch.pow(weights, anneal), num_samples=S)
    sp_ref = torch.cat([rs1.spacing_starts[..., 0], rs1.spacing_ends[..., -1:, 0]], -1)
    eu_ref = torch.cat([rs1.frustums.starts[..., 0], rs1.frustums.ends[..., -1:, 0]], -1)
    prev = torch.cat([rs0.spacing_starts[..., 0], rs0.spacing_ends[..., -1:, 0]], -1).expand(R, Sp + 1).contiguous()
    u_base = torch.linspace(0.0, 1.0 - 1.0 / (S + 1), steps=S + 1)
    sp, eu, inds = ops.sample_pdf(weights.to(dev), anneal, prev.to(dev), nears.to(dev), fars.to(dev), L.SPACING_LINDISP_PIECEWISE, u_base.to(dev),
                                  rand.to(dev) if training else None, S, want_inds=True)
    same = (inds.cpu().to(torch.int64) == pdf.last_inds).float().mean().item()
    # the searchsorted bin is exact whenever the cdf is; the cdf can differ in the last ulp (powf / fp32 sum order)
    # (the oracle's fp32 sum order depends on the host's thread count, so the handful of last-ulp ties is not identical from box to box)
    assert same >= 0.9999, f"searchsorted bins agree on {same*100:.3f}%"
    assert_close(sp, sp_ref, 1e-5, "pdf spacing bins", floor=1e-2, frac=0.999)
    assert_close(eu, eu_ref, 2e-4, "pdf euclid bins", floor=1e-2, frac=0.999)
    assert (sp[:, 1:] >= sp[:, :-1]).all(), "bins must be sorted"


@pytest.mark.parametrize("training", [False, True])
@pytest.mark.parametrize("Sp,S", [(256, 96), (96, 48), (64, 256)])
def test_pdf_sampler_bit_exact_given_identical_cdf(dev, training, Sp, S):
    """PDFSampler is bit-exact given an identical cdf.  The only places where the kernel and torch may round differently are the fp32
    sum of the histogram (torch: vectorised pairwise order, thread-count dependent), `pow(w, anneal)` (libm vs CUDA) and the cumsum
    order; everything after the cdf (u, searchsorted side="right", clamp, the two gathers, the lerp, the spacing -> euclidean map) is
    op-for-op the reference's.  Here the histogram is made of dyadic rationals with a power-of-two sum (weights k/2^16, padding 2^-7,
    anneal 1), so every sum and every division of the cdf is EXACT in any order: both sides hold the identical cdf, and the bins, the
    new spacing bins and the euclidean bins must then be equal bit for bit."""
    R = 512
    rays, nears, fars = _ray_case(R, Sp, 5)
    g = torch.Generator().manual_seed(1000 + Sp + S)
    pad = 2.0 ** -7
    total = 2 ** 18 if Sp * 512 < 2 ** 17 + 1 else 2 ** 19      # (sum k_j + Sp * 512) / 2^16 is a power of two
    budget = total - Sp * 512
    assert budget > 0
    k = torch.rand((R, Sp), generator=g) ** 6                     # peaky histograms
    k[0] = 1.0                                                    # flat
    k[1, : Sp // 2] = 0.0                                         # empty first half
    k = torch.floor(k / k.sum(-1, keepdim=True) * (budget - Sp)).to(torch.int64)
    k[:, -1] += budget - k.sum(-1)                                # exact integer total
    assert (k >= 0).all() and (k.sum(-1) == budget).all()
    weights = (k.double() / 2.0 ** 16).float()[..., None]
    assert torch.equal(weights.double()[..., 0] * 2.0 ** 16, k.double())
    initial = ns.UniformLinDispPiecewiseSampler(num_samples=Sp, single_jitter=True)
    initial.eval()
    rb = ns.RayBundle(rays["origins"], rays["directions"], rays["pixel_area"], rays["camera_indices"], nears, fars)
    rs0 = initial(rb)
    pdf = ns.PDFSampler(include_original=False, single_jitter=True, histogram_padding=pad)
    pdf.train(training)
    rand = torch.rand((R, 1), generator=g)
    pdf.rand_fn = lambda shape, device=None, dtype=None: rand
    rs1 = pdf(rb, rs0, weights, num_samples=S)
    sp_ref = torch.cat([rs1.spacing_starts[..., 0], rs1.spacing_ends[..., -1:, 0]], -1)
    eu_ref = torch.cat([rs1.frustums.starts[..., 0], rs1.frustums.ends[..., -1:, 0]], -1)
    prev = torch.cat([rs0.spacing_starts[..., 0], rs0.spacing_ends[..., -1:, 0]], -1).expand(R, Sp + 1).contiguous()
    u_base = torch.linspace(0.0, 1.0 - 1.0 / (S + 1), steps=S + 1)
    sp, eu, inds = ops.sample_pdf(weights.to(dev), 1.0, prev.to(dev), nears.to(dev), fars.to(dev), L.SPACING_LINDISP_PIECEWISE, u_base.to(dev),
                                  rand.to(dev) if training else None, S, histogram_padding=pad, want_inds=True)
    assert torch.equal(inds.cpu().to(torch.int64), pdf.last_inds), "searchsorted bins differ although the cdf is identical"
    assert torch.equal(sp.cpu(), sp_ref), (sp.cpu() - sp_ref).abs().max()
    assert torch.equal(eu.cpu(), eu_ref), (eu.cpu() - eu_ref).abs().max()


@pytest.mark.parametrize("S", [48, 96, 256, 1, 31, 500])
def test_weights_and_renderers(dev, S):
    R = 333
    g = torch.Generator().manual_seed(S)
    edges = torch.sort(torch.rand((R, S + 1), generator=g) * 6.0, dim=-1).values
    density = torch.exp(torch.randn((R, S, 1), generator=g) * 2.0)
    density[0] = 0.0
    density[1] = 1e4
    rgb = torch.rand((R, S, 3), generator=g)
    sem = torch.randn((R, S, 1), generator=g)
    fr = ns.Frustums(torch.zeros(R, S, 3), torch.zeros(R, S, 3), edges[:, :-1, None], edges[:, 1:, None], torch.ones(R, S, 1))
    rs = ns.RaySamples(fr, deltas=edges[:, 1:, None] - edges[:, :-1, None])
    dr = density.clone().requires_grad_(True)
    rr = rgb.clone().requires_grad_(True)
    sr = sem.clone().requires_grad_(True)
    w_ref = rs.get_weights(dr)
    rgb_ref = ns.RGBRenderer("last_sample").train()(rr, w_ref)
    acc_ref = ns.AccumulationRenderer()(w_ref)
    sem_ref = ns.SemanticRenderer()(sr, w_ref)
    depth_r = ns.DepthRenderer("median")
    depth_ref = depth_r(w_ref, rs)
    g1, g2, g3 = torch.randn((R, 3), generator=g), torch.randn((R, 1), generator=g), torch.randn((R, 1), generator=g)
    (rgb_ref * g1).sum().add((acc_ref * g2).sum()).add((sem_ref * g3).sum()).backward()

    e_dev = edges.to(dev)
    dm = density.to(dev).requires_grad_(True)
    rm = rgb.to(dev).requires_grad_(True)
    sm = sem.to(dev).requires_grad_(True)
    w = ops.ray_weights(dm, e_dev[:, :-1, None], e_dev[:, 1:, None])
    # alpha = 1 - exp(-dd) cancels for small dd, so one ulp of expf (CUDA vs CPU libm) is a large *relative* error of a
    # tiny weight; weights live in [0,1] and every renderer sums them, so the bound that matters is absolute
    assert (w.detach().cpu() - w_ref.detach()).abs().max().item() < 5e-7, "weights"
    rgb_o, acc_o, sem_o = ops.render(w, rm, sm, L.BG_LAST_SAMPLE, None, False)
    assert_close(rgb_o, rgb_ref, 1e-5, "rgb", floor=1e-3)
    assert_close(acc_o, acc_ref, 1e-4, "acc", floor=1e-3)
    assert_close(sem_o, sem_ref, 1e-4, "sem", floor=1e-2)
    # median index: bit-exact given identical weights (cumsum in double like torch's CPU cumsum)
    depth, idx = ops.render_median_depth(w_ref.detach().to(dev), e_dev[:, :-1, None], e_dev[:, 1:, None], want_index=True)
    assert torch.equal(idx.cpu().to(torch.int64), depth_r.last_median_index[:, 0])
    assert torch.equal(depth.cpu(), depth_ref)
    (rgb_o * g1.to(dev)).sum().add((acc_o * g2.to(dev)).sum()).add((sem_o * g3.to(dev)).sum()).backward()
    assert_close_scaled(rm.grad, rr.grad, 1e-5, "d_rgb")  # = g * w: same absolute-error argument as the weights
    assert_close_scaled(sm.grad, sr.grad, 1e-5, "d_sem")
    scale = dr.grad.abs().max().item()
    assert ((dm.grad.cpu() - dr.grad).abs().max().item() / scale) < 1e-4, "d_density"


def test_render_eval_mode_and_constant_background(dev):
    R, S = 64, 48
    g = torch.Generator().manual_seed(0)
    w = torch.rand((R, S, 1), generator=g) / S
    rgb = torch.rand((R, S, 3), generator=g)
    rgb[0, 0, 0] = float("nan")
    rgb[1, 3, 1] = float("inf")
    ref = ns.RGBRenderer(background_color="last_sample").eval()
    with ns.background_color_override_context(torch.zeros(3)):
        out_ref = ref(rgb, w)
    out, _, _ = ops.render(w.to(dev), rgb.to(dev), None, L.BG_CONSTANT, (0.0, 0.0, 0.0), True)
    assert_close(out, out_ref, 1e-5, "eval rgb black bg")
    out_ref2 = ns.RGBRenderer(background_color="last_sample").eval()(rgb, w)
    out2, _, _ = ops.render(w.to(dev), rgb.to(dev), None, L.BG_LAST_SAMPLE, None, True)
    assert_close(out2, out_ref2, 1e-5, "eval rgb last_sample")


@pytest.mark.parametrize("Sc,Sp", [(48, 256), (48, 96), (64, 512), (5, 3)])
def test_interlevel_and_distortion(dev, Sc, Sp):
    R = 200
    g = torch.Generator().manual_seed(Sc * 1000 + Sp)
    c = torch.sort(torch.rand((R, Sc + 1), generator=g), -1).values
    cp = torch.sort(torch.rand((R, Sp + 1), generator=g), -1).values
    cp[:, 0] = 0.0
    cp[:, -1] = 1.0
    cp[3] = torch.linspace(0, 1, Sp + 1)
    c[3, : min(Sc + 1, Sp + 1)] = cp[3, : min(Sc + 1, Sp + 1)]  # coincident edges: searchsorted side matters
    c = torch.sort(c, -1).values
    w = torch.rand((R, Sc), generator=g) / Sc
    wp = (torch.rand((R, Sp), generator=g) / Sp).requires_grad_(True)
    loss_ref = torch.mean(ns.lossfun_outer(c, w, cp, wp))
    loss_ref.backward()
    wpm = wp.detach().to(dev).requires_grad_(True)
    loss = ops.interlevel_term(c.to(dev), w.to(dev), cp.to(dev), wpm)
    assert abs(loss.item() - loss_ref.item()) <= 1e-5 * abs(loss_ref.item()) + 1e-9
    (loss * 3.0).backward()
    scale = wp.grad.abs().max().item() * 3.0 + 1e-20
    assert ((wpm.grad.cpu() - 3.0 * wp.grad).abs().max().item() / scale) < 1e-4
    d_ref = torch.mean(ns.lossfun_distortion(c, w))
    d = ops.distortion(c.to(dev), w.to(dev))
    assert abs(d.item() - d_ref.item()) <= 1e-4 * abs(d_ref.item())


def test_pixel_losses(dev):
    R = 4096
    g = torch.Generator().manual_seed(4)
    rgb = torch.rand((R, 3), generator=g).requires_grad_(True)
    sem = (torch.randn((R, 1), generator=g) * 4).requires_grad_(True)
    img = torch.rand((R, 3), generator=g)
    mask = (torch.rand((R, 1), generator=g) < 0.1).float()
    l_ref = torch.nn.MSELoss()(img, rgb) + 0.7 * torch.nn.BCEWithLogitsLoss()(sem, mask)
    l_ref.backward()
    rm = rgb.detach().to(dev).requires_grad_(True)
    sm = sem.detach().to(dev).requires_grad_(True)
    mse, bce = ops.pixel_losses(rm, sm, img.to(dev), mask.to(dev), 0.7)
    assert abs((mse + bce).item() - l_ref.item()) < 1e-5 * abs(l_ref.item())
    (mse + bce).backward()
    assert_close(rm.grad, rgb.grad, 1e-5, "d_rgb", floor=1e-7)
    assert_close(sm.grad, sem.grad, 1e-5, "d_sem", floor=1e-7)


def test_adam_matches_torch(dev):
    g = torch.Generator().manual_seed(5)
    p0 = torch.randn((10007,), generator=g)
    ref = torch.nn.Parameter(p0.clone())
    opt = torch.optim.Adam([ref], lr=1e-2, eps=1e-15)
    p = p0.clone().to(dev)
    m = torch.zeros_like(p)
    v = torch.zeros_like(p)
    for step in range(1, 6):
        grad = torch.randn((10007,), generator=g) * 10.0 ** float(step - 3)
        ref.grad = grad.clone()
        opt.step()
        ops.adam_step(p, grad.to(dev), m, v, 1e-2, step)
        assert_close(p, ref.detach(), 1e-5, f"adam step {step}", floor=1e-3)


@pytest.mark.parametrize("kind", ["adam", "radam"])
def test_explicit_scalar_update_matches_torch_adam_and_radam(dev, kind):
    """The flat-group update kernel with host-computed scalars (engine.step_scalars) against torch.optim.Adam / torch.optim.RAdam (the
    optimizer of the _big / _huge presets, fruit_nerf_config.py:100-108): the unrectified first steps (rho_t <= 5, no adaptive denominator),
    the rectified ones, and the gradient clear.  (Flat groups are 16-byte aligned with n % 4 == 0: engine.FlatGroup pads every tensor to 256 bytes.)"""
    from cropnerf_b200.engine import OptimizerSpec, step_scalars

    g = torch.Generator().manual_seed(6)
    p0 = torch.randn((10008,), generator=g)
    ref = torch.nn.Parameter(p0.clone())
    opt = (torch.optim.RAdam if kind == "radam" else torch.optim.Adam)([ref], lr=1e-2, eps=1e-15)
    spec = OptimizerSpec(lr=1e-2, eps=1e-15, lr_final=None, kind=kind)
    p = p0.clone().to(dev)
    m, v = torch.zeros_like(p), torch.zeros_like(p)
    for step in range(1, 13):
        grad = torch.randn((10008,), generator=g) * 10.0 ** float((step % 5) - 2)
        ref.grad = grad.clone()
        opt.step()
        gd = grad.to(dev)
        ops.adam_step_scalars(p, gd, m, v, step_scalars(spec, 1e-2, step))
        assert float(gd.abs().max()) == 0.0, "the update clears the gradient"
        # unrectified RAdam steps are lr * m_hat, O(1) with these gradients: fp32 rounding of the update is ~1e-7 absolute, so the floor is 0.1
        assert_close(p, ref.detach(), 1e-5, f"{kind} step {step}", floor=1e-1)


@pytest.mark.parametrize("with_pixels", [False, True])
def test_generate_rays_and_aabb_clip(dev, with_pixels):
    """Device ray generation + AABB slab test ("next" row f1) against the restated nerfstudio generate_rays / intersect_aabb."""
    from cropnerf_b200 import synthetic
    from cropnerf_b200.export import generate_rays

    W, H = 97, 61
    c2w = synthetic.make_cameras(5, seed=3)[2]
    fx, fy, cx, cy = 80.0, 82.0, 48.5, 30.25
    aabb = torch.tensor([[-0.3, -0.2, -0.25], [0.2, 0.3, 0.15]])
    if with_pixels:
        g = torch.Generator().manual_seed(1)
        pix = torch.stack([torch.randint(0, H, (500,), generator=g), torch.randint(0, W, (500,), generator=g)], -1)
        coords = pix.float() + 0.5
    else:
        pix = None
        yy, xx = torch.meshgrid(torch.arange(H), torch.arange(W), indexing="ij")
        coords = torch.stack([yy, xx], -1).reshape(-1, 2).float() + 0.5
    o_ref, d_ref, a_ref = ns.generate_pinhole_rays(c2w, fx, fy, cx, cy, coords)
    tmin, tmax = ns.intersect_aabb(o_ref, d_ref, aabb.reshape(-1))
    rb, cnt = generate_rays(c2w, fx, fy, cx, cy, W, H, dev, aabb=aabb, pixel_yx=pix, count_valid=True)
    assert_close(rb.origins, o_ref, 1e-6, "origins")
    assert_close(rb.directions, d_ref, 2e-6, "directions", floor=1e-2)
    assert_close(rb.pixel_area, a_ref, 2e-3, "pixel_area", floor=1e-9)
    hit_ref = tmin < 1e10
    hit = rb.nears[:, 0].cpu() < 1e10
    assert (hit == hit_ref).float().mean().item() >= 0.999
    both = hit & hit_ref
    assert both.sum() > 50, "test scene should hit the box"
    assert_close(rb.nears[:, 0].cpu()[both], tmin[both], 1e-4, "nears", floor=1e-2)
    assert_close(rb.fars[:, 0].cpu()[both], tmax[both], 1e-4, "fars", floor=1e-2)
    assert int(cnt.item()) == int(hit.sum().item())
    assert torch.allclose(rb.directions.norm(dim=-1).cpu(), torch.ones(rb.directions.shape[0]), atol=1e-5)


def test_camera_opt_kernels_against_torch(dev):
    """cnb_camera_opt_apply / cnb_camera_opt_bwd (csrc/camera_opt.cu) against the torch expression of nerfstudio's CameraOptimizer(SO3xR3):
    exp_map_SO3xR3, apply_to_raybundle, the regulariser, and autograd for the pose gradient."""
    from cropnerf_b200.fruit_nerf import exp_map_SO3xR3

    g = torch.Generator().manual_seed(1)
    C_, R = 9, 2000
    pose = torch.randn((C_, 6), generator=g) * 0.05
    pose[0] = 0.0
    pose[1, 3:] = torch.tensor([1e-3, -2e-3, 5e-4])
    cam = torch.randint(0, C_, (R,), generator=g)
    o = torch.randn((R, 3), generator=g)
    d = torch.nn.functional.normalize(torch.randn((R, 3), generator=g), dim=-1)
    go, gd = torch.randn((R, 3), generator=g) * 1e-3, torch.randn((R, 3), generator=g) * 1e-3
    pr = pose.clone().requires_grad_(True)
    M = exp_map_SO3xR3(pr[cam])
    o2 = o + M[:, :3, 3]
    d2 = torch.bmm(M[:, :3, :3], d[..., None]).squeeze(-1)
    reg = pr[:, :3].norm(dim=-1).mean() * 1e-2 + pr[:, 3:].norm(dim=-1).mean() * 1e-3
    ((o2 * go).sum() + (d2 * gd).sum() + reg).backward()
    pd, camd, od, dd = pose.to(dev), cam.to(dev, torch.int32), o.to(dev), d.to(dev)
    o_out, d_out = torch.empty_like(od), torch.empty_like(dd)
    L.check(L.lib().cnb_camera_opt_apply(pd.data_ptr(), camd.data_ptr(), od.data_ptr(), dd.data_ptr(), R, C_, o_out.data_ptr(), d_out.data_ptr(), L.stream_ptr(dev)), "apply")
    assert (o_out.cpu() - o2.detach()).abs().max() < 1e-6 and (d_out.cpu() - d2.detach()).abs().max() < 1e-6
    scratch = torch.empty((12 * C_,), device=dev)
    dpose = torch.zeros((C_, 6), device=dev)
    regl = torch.zeros((1,), device=dev)
    god, gdd = go.to(dev), gd.to(dev)
    L.check(L.lib().cnb_camera_opt_bwd(pd.data_ptr(), camd.data_ptr(), dd.data_ptr(), god.data_ptr(), gdd.data_ptr(), R, C_, 1e-2, 1e-3, 1.0,
                                       scratch.data_ptr(), dpose.data_ptr(), regl.data_ptr(), L.stream_ptr(dev)), "bwd")
    assert abs(float(regl) - float(reg.detach())) <= 1e-6 * float(reg.detach())
    err = (dpose.cpu() - pr.grad).abs().max().item() / pr.grad.abs().max().item()
    assert err < 1e-4, err
