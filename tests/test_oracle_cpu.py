"""CPU tests of the oracle itself: golden-fixture regression, known properties of the restated nerfstudio
primitives, and autograd consistency.  The reference ships no tests or vectors for this path
("parity unpinned"), so these pin the restatement."""
import os

import numpy as np
import pytest
import torch

from helpers import GOLDEN

from oracle import cases
from oracle import nerfstudio_torch as ns


@pytest.mark.parametrize("name", list(cases.CASES))
def test_golden_regression(name):
    ref = dict(np.load(os.path.join(GOLDEN, name + ".npz")))
    out = cases.run_case(name)
    assert set(out) == set(ref)
    for k, v in ref.items():
        if v.dtype.kind in "iu":
            assert np.array_equal(out[k], v), k
        else:
            np.testing.assert_allclose(out[k], v, rtol=2e-5, atol=1e-7, err_msg=k)


def test_scalings_match_survey():
    # fp32 pow rounding makes the field's top level 2047
    assert ns.hash_scalings(16, 16, 2048).tolist() == [16, 22, 30, 42, 58, 80, 111, 153, 212, 294, 406, 561, 776, 1072, 1482, 2047]
    assert ns.hash_scalings(5, 16, 128).tolist() == [16, 26, 45, 76, 128]
    assert ns.hash_scalings(5, 16, 256).tolist() == [16, 32, 64, 128, 256]


def test_hash_int64_equals_uint32_wraparound():
    # App. A.1: for non-negative coords and power-of-two tables the int64 hash equals uint32 wrap-around arithmetic
    enc = ns.HashEncoding(num_levels=3, min_res=16, max_res=2048, log2_hashmap_size=19)
    g = torch.Generator().manual_seed(0)
    v = torch.randint(0, 2049, (50000, 3, 3), generator=g, dtype=torch.int32)
    ref = enc.hash_fn(v) - enc.hash_offset
    x = v.numpy().astype(np.uint32)
    mine = (x[..., 0] ^ (x[..., 1] * np.uint32(2654435761)) ^ (x[..., 2] * np.uint32(805459861))) & np.uint32(2**19 - 1)
    assert np.array_equal(ref.numpy().astype(np.uint32), mine)


def test_hash_encoding_is_trilinear_at_nodes():
    enc = ns.HashEncoding(num_levels=1, min_res=4, max_res=4, log2_hashmap_size=10, hash_init_scale=1.0)
    node = torch.tensor([[0.25, 0.5, 0.75]])  # exact grid node at resolution 4
    idx, off = enc.corner_indices(node)
    assert torch.all(off == 0)
    assert torch.all(idx[..., 0] == idx[..., 6])  # ceil == floor
    out = enc(node)
    assert torch.equal(out[0], enc.hash_table[idx[0, 0, 6]])


def test_contraction_and_spacing():
    c = ns.SceneContraction()
    x = torch.tensor([[0.5, -0.2, 0.1], [4.0, 0.0, 0.0], [-1.0, 8.0, 2.0]])
    y = c(x)
    assert torch.equal(y[0], x[0])
    assert torch.allclose(y[1], torch.tensor([1.75, 0.0, 0.0]))
    assert y.abs().max() < 2
    s = ns.UniformLinDispPiecewiseSampler()
    t = torch.tensor([0.0, 0.3, 1.0, 7.0, 1000.0])
    assert torch.allclose(s.spacing_fn_inv(s.spacing_fn(t)), t, rtol=1e-4)


def test_trunc_exp_gradient_is_clamped():
    x = torch.tensor([-20.0, 0.0, 20.0], requires_grad=True)
    ns.trunc_exp(x).sum().backward()
    assert torch.allclose(x.grad, torch.exp(torch.tensor([-15.0, 0.0, 15.0])))


def test_weights_sum_to_opacity_and_pdf_is_monotone():
    g = torch.Generator().manual_seed(1)
    R, S = 50, 64
    edges = torch.sort(torch.rand((R, S + 1), generator=g) * 4, -1).values
    fr = ns.Frustums(torch.zeros(R, S, 3), torch.ones(R, S, 3), edges[:, :-1, None], edges[:, 1:, None], torch.ones(R, S, 1))
    rs = ns.RaySamples(fr, deltas=edges[:, 1:, None] - edges[:, :-1, None], spacing_starts=edges[:, :-1, None] / 4, spacing_ends=edges[:, 1:, None] / 4,
                       spacing_to_euclidean_fn=lambda x: x * 4)
    density = torch.rand((R, S, 1), generator=g) * 3
    w = rs.get_weights(density)
    total = (rs.deltas * density).sum(-2)
    assert torch.allclose(w.sum(-2), 1 - torch.exp(-total), atol=1e-5)
    rb = ns.RayBundle(torch.zeros(R, 3), torch.ones(R, 3), torch.ones(R, 1), torch.zeros(R, 1, dtype=torch.long), torch.zeros(R, 1), torch.ones(R, 1) * 4)
    pdf = ns.PDFSampler(include_original=False).eval()
    out = pdf(rb, rs, w, num_samples=32)
    bins = torch.cat([out.spacing_starts[..., 0], out.spacing_ends[..., -1:, 0]], -1)
    assert torch.all(bins[:, 1:] >= bins[:, :-1])
    assert pdf.last_inds.min() >= 1 and pdf.last_inds.max() <= S


def test_interlevel_loss_zero_when_proposal_envelopes():
    # a proposal histogram that upper-bounds the fine one has zero loss (mip-NeRF 360 eq. 13)
    t = torch.linspace(0, 1, 9)[None]
    w = torch.full((1, 8), 0.1)
    assert ns.lossfun_outer(t, w, t, w * 1.5).abs().max() == 0
    assert ns.lossfun_outer(t, w, t, w * 0.2).min() > 0  # even the 3-bin outer envelope (coincident edges) stays below w


def test_oracle_field_gradcheck_fp64():
    cfg = cases.make_config(dict(log2_hashmap_size=8, num_levels=4, max_res=64))
    torch.manual_seed(0)
    from oracle import fruit_torch as ft

    field = ft.FruitField(torch.tensor([[-1.0] * 3, [1.0] * 3]), num_images=4, num_levels=4, max_res=64, log2_hashmap_size=8,
                          spatial_distortion=ns.SceneContraction()).double()
    field.train()
    R, S = 3, 4
    g = torch.Generator().manual_seed(0)
    edges = torch.sort(torch.rand((R, S + 1), generator=g, dtype=torch.float64) * 2, -1).values
    rb = ns.RayBundle(torch.rand((R, 3), generator=g, dtype=torch.float64) - 0.5, torch.rand((R, 3), generator=g, dtype=torch.float64),
                      torch.ones(R, 1, dtype=torch.float64), torch.randint(0, 4, (R, 1), generator=g))
    rs = rb.get_ray_samples(edges[:, :-1, None], edges[:, 1:, None])
    table = field.mlp_base_grid.hash_table

    def fn(tab):
        field.mlp_base_grid.hash_table.data = tab
        out = field(rs)
        return out["density"].sum() + out["rgb"].sum()

    t0 = table.detach().clone()
    out = field(rs)
    (out["density"].sum() + out["rgb"].sum()).backward()
    analytic = table.grad.clone()
    rows = analytic.abs().sum(-1).nonzero()[:3, 0]
    for r in rows.tolist():
        for c in range(2):
            tp, tm = t0.clone(), t0.clone()
            tp[r, c] += 1e-6
            tm[r, c] -= 1e-6
            with torch.no_grad():
                num = (fn(tp) - fn(tm)) / 2e-6
            assert abs(num.item() - analytic[r, c].item()) <= 1e-5 * max(1.0, abs(num.item()))


def test_sh_components_equal_scipy_real_spherical_harmonics():
    """Independent pin of the SH basis (App. A.4): the 16 hard-coded polynomials of components_from_spherical_harmonics are the
    orthonormal REAL spherical harmonics of degree <= 3 without the Condon-Shortley phase, ordered l^2 + l + m -- checked against
    scipy.special.sph_harm_y on random unit vectors (an external implementation, not a restatement)."""
    from scipy import special

    rng = np.random.default_rng(0)
    d = rng.normal(size=(2000, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    theta, phi = np.arccos(np.clip(d[:, 2], -1.0, 1.0)), np.arctan2(d[:, 1], d[:, 0])
    out = ns.components_from_spherical_harmonics(3, torch.tensor(d, dtype=torch.float64)).numpy()
    for l in range(4):
        for m in range(-l, l + 1):
            Y = special.sph_harm_y(l, abs(m), theta, phi)
            want = Y.real if m == 0 else np.sqrt(2.0) * (-1) ** m * (Y.real if m > 0 else Y.imag)
            assert np.abs(want - out[:, l * l + l + m]).max() < 1e-13, (l, m)
    # orthonormality on the sphere (Monte-Carlo: 2000 points -> a few percent)
    gram = 4.0 * np.pi * (out.T @ out) / out.shape[0]
    assert np.abs(gram - np.eye(16)).max() < 0.15


def _ray_samples_from_edges(edges: torch.Tensor, far: float):
    R, S = edges.shape[0], edges.shape[1] - 1
    fr = ns.Frustums(torch.zeros(R, S, 3, dtype=edges.dtype), torch.ones(R, S, 3, dtype=edges.dtype), edges[:, :-1, None], edges[:, 1:, None],
                     torch.ones(R, S, 1, dtype=edges.dtype))
    return ns.RaySamples(fr, deltas=edges[:, 1:, None] - edges[:, :-1, None], spacing_starts=edges[:, :-1, None] / far,
                         spacing_ends=edges[:, 1:, None] / far, spacing_to_euclidean_fn=lambda x: x * far)


def test_weights_equal_the_transmittance_integral_of_a_piecewise_constant_medium():
    """get_weights (App. A.5) against the volume-rendering integral it discretises: for sigma constant on each bin,
    w_i = T(t_i) - T(t_{i+1}) with T(t) = exp(-int_0^t sigma), evaluated independently in numpy float64."""
    rng = np.random.default_rng(3)
    R, S = 20, 48
    edges = np.sort(rng.random((R, S + 1)) * 4.0, -1)
    sigma = rng.random((R, S)) * 5.0
    rs = _ray_samples_from_edges(torch.tensor(edges), 4.0)
    w = rs.get_weights(torch.tensor(sigma)[..., None])[..., 0].numpy()
    tau = np.concatenate([np.zeros((R, 1)), np.cumsum(sigma * np.diff(edges, axis=-1), -1)], -1)  # optical depth at every edge
    T = np.exp(-tau)
    assert np.abs(w - (T[:, :-1] - T[:, 1:])).max() < 1e-12


def test_pdf_sampler_is_the_inverse_cdf_of_the_padded_histogram():
    """PDFSampler (App. A.6) in eval mode against numpy's own piecewise-linear interpolation: the new bin edges are
    F^-1(u) at the bin-centre quantiles u_k = (k + 1/2) / (n + 1), F = cdf of (weights + 0.01) over the existing edges."""
    rng = np.random.default_rng(4)
    R, S, n = 16, 64, 32
    edges = np.sort(rng.random((R, S + 1)), -1)
    edges[:, 0], edges[:, -1] = 0.0, 1.0
    wts = rng.random((R, S)) ** 4
    rs = _ray_samples_from_edges(torch.tensor(edges), 1.0)
    rb = ns.RayBundle(torch.zeros(R, 3, dtype=torch.float64), torch.ones(R, 3, dtype=torch.float64), torch.ones(R, 1, dtype=torch.float64),
                      torch.zeros(R, 1, dtype=torch.long), torch.zeros(R, 1, dtype=torch.float64), torch.ones(R, 1, dtype=torch.float64))
    out = ns.PDFSampler(include_original=False).eval()(rb, rs, torch.tensor(wts)[..., None], num_samples=n)
    got = torch.cat([out.spacing_starts[..., 0], out.spacing_ends[..., -1:, 0]], -1).numpy()
    u = (np.arange(n + 1) + 0.5) / (n + 1)
    for r in range(R):
        p = wts[r] + 0.01
        cdf = np.concatenate([[0.0], np.cumsum(p / p.sum())])
        assert np.abs(np.interp(u, cdf, edges[r]) - got[r]).max() < 1e-9
    # and the samples are distributed like the histogram: mass of every original bin ~ number of new edges inside it
    hist = np.stack([np.histogram(got[r], bins=edges[r])[0] for r in range(R)])
    p = (wts + 0.01) / (wts + 0.01).sum(-1, keepdims=True)
    assert np.abs(hist / (n + 1) - p).max() < 1.0 / (n + 1) + 1e-9


def test_distortion_loss_equals_the_double_integral():
    """lossfun_distortion (App. A.8, mip-NeRF 360 eq. 15) against the quantity it is the closed form of: the double integral of
    w(u) w(v) |u - v| for the piecewise-constant density w_i / delta_i, integrated numerically on a fine grid."""
    rng = np.random.default_rng(5)
    S = 6
    t = np.array([0, 3, 11, 12, 30, 41, 50]) / 50.0  # edges on the integration grid: the midpoint rule is then exact up to O(1/G^2)
    w = rng.random(S)
    closed = ns.lossfun_distortion(torch.tensor(t)[None], torch.tensor(w)[None]).item()
    G = 4000
    x = (np.arange(G) + 0.5) / G
    dens = (w / np.diff(t))[np.clip(np.searchsorted(t, x, side="right") - 1, 0, S - 1)]
    numeric = (dens[:, None] * dens[None, :] * np.abs(x[:, None] - x[None, :])).sum() / G**2
    assert abs(closed - numeric) < 1e-5 * closed


def test_interlevel_outer_measure_bounds_the_overlap_by_brute_force():
    """`outer` (App. A.8, mip-NeRF 360 eq. 13): for every fine interval, the summed weight of ALL proposal bins that touch it --
    checked against an explicit double loop, and it must upper-bound the exactly overlapping proposal mass."""
    rng = np.random.default_rng(6)
    for _ in range(5):
        tf = np.sort(rng.random(13)); tf[0], tf[-1] = 0.0, 1.0
        tp = np.sort(rng.random(9)); tp[0], tp[-1] = 0.0, 1.0
        wp = rng.random(8)
        got = ns.outer(torch.tensor(tf[:-1])[None], torch.tensor(tf[1:])[None], torch.tensor(tp[:-1])[None], torch.tensor(tp[1:])[None],
                       torch.tensor(wp)[None])[0].numpy()
        for i in range(12):
            a, b = tf[i], tf[i + 1]
            touching = sum(wp[j] for j in range(8) if tp[j + 1] > a and tp[j] < b)          # open-interval intersection
            overlap = sum(wp[j] * max(0.0, min(b, tp[j + 1]) - max(a, tp[j])) / (tp[j + 1] - tp[j]) for j in range(8))
            assert got[i] >= touching - 1e-12 and got[i] >= overlap - 1e-12
            closed_touching = sum(wp[j] for j in range(8) if tp[j + 1] >= a and tp[j] <= b)  # closed intervals: at most one bin more per side
            assert got[i] <= closed_touching + 1e-12
