"""Size-independent properties at BASELINE.json's full sizes (fruit_nerf preset: 2^19 x 16 field table, 2^17 x 5 proposal tables;
32 768-ray render chunks, 4096-ray training batches) -- sizes the CPU oracle does not finish in seconds.  These do not need an oracle:
they are invariants of the reference's algorithm (ray independence -> permutation / chunking / sharding invariance, monotone PDF
samples, transmittance bounds, linearity of compositing in the colours, determinism and range of the integer hash indices)."""
import numpy as np
import pytest
import torch

from helpers import product_bundle, product_model

from cropnerf_b200 import engine, ops, synthetic
from cropnerf_b200.export import shard_range
from cropnerf_b200.fruit_nerf import FruitModel, FruitNerfModelConfig
from cropnerf_b200.rays import RayBundle

pytestmark = pytest.mark.gpu

R_RENDER = 32768


def _full_model(dev, precision="fp32", seed=0):
    torch.manual_seed(seed)
    model = FruitModel(FruitNerfModelConfig(precision=precision), num_train_data=300)
    model.load_state_dict(synthetic.randomize_state(model.state_dict(), seed=seed, table_scale=0.5))
    return model.to(dev).eval()


def _bundle(rays, dev, idx=None):
    pick = (lambda t: t) if idx is None else (lambda t: t[idx])
    return RayBundle(pick(rays["origins"]).to(dev), pick(rays["directions"]).to(dev), pick(rays["pixel_area"]).to(dev), pick(rays["camera_indices"]).to(dev))


@pytest.mark.parametrize("precision", ["fp32", "mixed"])
def test_render_is_invariant_to_permutation_chunking_and_sharding(dev, precision):
    """Rays are independent: rendering a permuted / chunked / rank-sharded batch gives the same BITS per ray -- the property the
    multi-GPU export split (rays sharded by range, no communication) rests on."""
    model = _full_model(dev, precision)
    rays = synthetic.make_rays(R_RENDER, seed=11, num_cameras=300)
    keys = ("rgb", "depth", "accumulation", "semantics", "prop_depth_0", "prop_depth_1")
    with torch.no_grad():
        whole = {k: v.clone() for k, v in model(_bundle(rays, dev)).items() if k in keys}
        perm = torch.randperm(R_RENDER, generator=torch.Generator().manual_seed(3))
        permuted = model(_bundle(rays, dev, perm))
        for k in keys:
            assert torch.equal(permuted[k], whole[k][perm.to(dev)]), f"{precision} {k}: permutation changed per-ray results"
        for world in (2, 8):  # export / projection sharding: contiguous ray ranges per rank
            parts = []
            for r in range(world):
                lo, hi = shard_range(R_RENDER, r, world)
                parts.append({k: v.clone() for k, v in model(_bundle(rays, dev, slice(lo, hi))).items() if k in keys})
            for k in keys:
                assert torch.equal(torch.cat([p[k] for p in parts]), whole[k]), f"{precision} {k}: {world}-way sharding changed per-ray results"
        # the chunk loop of get_outputs_for_camera_ray_bundle (fruit_nerf.py:377-404) with a ragged last chunk
        model.config.eval_num_rays_per_chunk = 5000
        host = model.get_outputs_for_camera_jagged_ray_bundle(_bundle(rays, dev))
        for k in keys:
            assert torch.equal(host[k].to(dev), whole[k]), f"{precision} {k}: chunked render differs"
    # physical bounds of the compositing
    acc = whole["accumulation"]
    assert float(acc.min()) >= 0.0 and float(acc.max()) <= 1.0 + 1e-5
    assert float(whole["rgb"].min()) >= 0.0 and float(whole["rgb"].max()) <= 1.0 + 1e-6  # eval: clamped (RGBRenderer)
    assert torch.isfinite(whole["depth"]).all() and float(whole["depth"].min()) >= 0.0 and float(whole["depth"].max()) <= 1000.0 * (1 + 1e-5)


def test_hash_indices_full_table_range_and_determinism(dev):
    """Integer work: indices of the full 16 x 2^19 grid are in [l*2^19, (l+1)*2^19), identical across calls and across batch splits."""
    g = torch.Generator().manual_seed(5)
    n = 1 << 20
    pos = torch.rand((n, 3), generator=g).to(dev)
    model = _full_model(dev)
    enc = model.field.mlp_base_grid
    scal = enc.scalings.tolist()
    feat, idx = ops.hashgrid_indices(pos, enc.hash_table, 16, 19, scal)
    feat2, idx2 = ops.hashgrid_indices(pos, enc.hash_table, 16, 19, scal)
    assert torch.equal(feat, feat2)
    assert torch.equal(idx, idx2)
    idx = idx.view(n, 16, 8).long()
    lv = torch.arange(16, device=dev).view(1, 16, 1) * (1 << 19)
    assert bool(((idx >= lv) & (idx < lv + (1 << 19))).all())
    _, half = ops.hashgrid_indices(pos[: n // 2].contiguous(), enc.hash_table, 16, 19, scal)
    assert torch.equal(half.view(-1, 16, 8).long(), idx[: n // 2])
    # level 0 (17^3 lattice points hashed into 2^19 slots) can touch at most 4913 distinct rows
    assert int(torch.unique(idx[:, 0, :]).numel()) <= 17**3


def test_pdf_samples_sorted_inside_near_far_and_counts(dev):
    """ProposalNetworkSampler at 32 768 rays: bin edges of every level are non-decreasing, inside [near, far], with the preset's counts."""
    model = _full_model(dev)
    rays = synthetic.make_rays(R_RENDER, seed=12, num_cameras=300)
    rb = model.collider(_bundle(rays, dev))
    with torch.no_grad():
        ray_samples, weights_list, ray_samples_list = model.proposal_sampler(rb, density_fns=model.density_fns)
    assert [int(rs.frustums.starts.shape[-2]) for rs in ray_samples_list + [ray_samples]] == [256, 96, 48]
    for rs in ray_samples_list + [ray_samples]:
        st, en = rs.frustums.starts[..., 0], rs.frustums.ends[..., 0]
        assert bool((en >= st).all()) and bool((st[:, 1:] >= st[:, :-1]).all())
        assert torch.equal(st[:, 1:], en[:, :-1])  # contiguous bins
        assert float(st.min()) >= 0.0 and float(en.max()) <= 1000.0 * (1 + 1e-5)  # eval: NearFarCollider resets the near plane to 0
    for w in weights_list:
        s = w[..., 0].sum(-1)
        assert float(w.min()) >= 0.0 and float(s.max()) <= 1.0 + 1e-4  # weights are a sub-probability distribution along the ray


def test_compositing_is_linear_in_the_colours(dev):
    """RGBRenderer without background: composite(w, a*c1 + b*c2) = a*composite(w, c1) + b*composite(w, c2) (4096 x 48 samples)."""
    from cropnerf_b200 import _lib as L

    g = torch.Generator().manual_seed(9)
    R, S = 4096, 48
    w = torch.rand((R, S, 1), generator=g).to(dev) / S
    c1, c2 = torch.rand((R, S, 3), generator=g).to(dev), torch.rand((R, S, 3), generator=g).to(dev)
    r1, acc, _ = ops.render(w, c1, None, L.BG_NONE, None, False)
    r2, _, _ = ops.render(w, c2, None, L.BG_NONE, None, False)
    r12, _, _ = ops.render(w, 0.25 * c1 + 2.0 * c2, None, L.BG_NONE, None, False)
    assert float((r12 - (0.25 * r1 + 2.0 * r2)).abs().max()) < 2e-6
    assert torch.allclose(acc[:, 0], w[..., 0].sum(-1), rtol=1e-6, atol=1e-7)


def test_training_step_is_deterministic_in_outputs_and_finite_at_full_size(dev):
    """Two trainers fed the same 4096-ray batch and jitter produce the same losses up to the order of the fp32 atomic loss sums (the forward is deterministic per ray) and finite, non-zero gradients in every parameter group of the full-size preset."""
    R = 4096
    rays = synthetic.make_rays(R, seed=21, num_cameras=300)
    targets = {k: v.to(dev) for k, v in synthetic.make_targets(R, seed=4).items()}
    jit = synthetic.make_jitter(R, 3, seed=8)
    losses, norms = [], []
    for _ in range(2):
        model = _full_model(dev, "mixed").train()
        feed = synthetic.JitterFeed(jit)
        model.proposal_sampler.initial_sampler.rand_fn = feed
        model.proposal_sampler.pdf_sampler.rand_fn = feed
        tr = engine.Trainer(model, force_proposal_update=True)
        fp = tr.fused
        feed.reset()
        ls, _ = fp.train_step(_bundle(rays, dev), targets, update_proposals=True)
        torch.cuda.synchronize()
        losses.append(ls[:4].clone())
        norms.append({n: float(g.grad.norm()) for n, g in tr.groups.items()})
    # the per-ray loss terms are identical; their batch sums are accumulated with fp32 atomics (order noise ~1e-6)
    assert float(((losses[0] - losses[1]).abs() / losses[0].abs()).max()) <= 1e-5
    for n in norms[0]:
        assert np.isfinite(norms[0][n]) and norms[0][n] > 0
        assert abs(norms[0][n] - norms[1][n]) <= 1e-3 * norms[0][n]


def test_unreachable_table_rows_never_get_gradient_and_masked_adam_is_bit_identical(dev):
    """The optimiser's skip bitmap (cnb_hashgrid_mark_reachable): rows of a coarse level that no lattice corner hashes to receive exactly zero
    gradient from a full-size training batch, so an Adam pass that skips them equals the dense pass bit for bit."""
    R = 4096
    model = _full_model(dev, "mixed").train()
    tr = engine.Trainer(model, force_proposal_update=True)
    frac = {n: g.live_fraction() for n, g in tr.groups.items()}
    assert 0.60 < frac["fields"] < 0.85, frac        # ~30 % of the 16-byte units of the 16 x 2^19 field table are unreachable (levels 0-5)
    assert 0.45 < frac["proposal_networks"] < 0.80, frac
    rays = synthetic.make_rays(R, seed=31, num_cameras=300)
    targets = {k: v.to(dev) for k, v in synthetic.make_targets(R, seed=5).items()}
    tr.fused.train_step(_bundle(rays, dev), targets, update_proposals=True)
    torch.cuda.synchronize()
    for name, g in tr.groups.items():
        n4 = g.flat.numel() // 4
        bits = torch.stack([(g.live >> k) & 1 for k in range(32)], dim=1).reshape(-1)[:n4].bool()
        gr = g.grad.view(n4, 4)
        assert float(gr[~bits].abs().max()) == 0.0, f"{name}: gradient landed in a row marked unreachable"
        assert float(gr[bits].abs().max()) > 0.0
        # masked vs dense fused Adam on copies of the real buffers
        outs = []
        for live in (g.live, None):
            p, gg, m, v = g.flat.clone(), g.grad.clone(), g.exp_avg.clone(), g.exp_avg_sq.clone()
            for step in (1, 2):
                ops.adam_step(p, gg, m, v, 1e-2, step, zero_grad=True, live=live)
                gg.copy_(g.grad)
            outs.append((p, m, v))
        for a, b in zip(outs[0], outs[1]):
            assert torch.equal(a, b), f"{name}: masked Adam differs from the dense pass"


@pytest.mark.parametrize("R", [0, 1, 33])
def test_empty_single_and_ragged_ray_bundles(dev, R):
    """Edge sizes of the render loop: an empty bundle, ONE ray (the reference notes its chunk loop cannot handle a size-1 chunk,
    fruit_nerf.py:289-291 -- here it is just a batch) and a count that is not a multiple of anything; each ray's result equals the same ray
    rendered inside a larger batch."""
    model = _full_model(dev)
    rays = synthetic.make_rays(64, seed=17, num_cameras=300)
    with torch.no_grad():
        big = model(_bundle(rays, dev))
        out = model(_bundle(rays, dev, slice(0, R)))
        host = model.get_outputs_for_camera_jagged_ray_bundle(_bundle(rays, dev, slice(0, R))) if R > 0 else None
    for k in ("rgb", "depth", "accumulation", "semantics"):
        assert out[k].shape[0] == R
        assert torch.equal(out[k], big[k][:R]), k
        if host is not None:
            assert torch.equal(host[k].to(dev), big[k][:R]), k
    if R > 0:
        tr = engine.Trainer(_full_model(dev, "mixed").train(), force_proposal_update=True)
        targets = {k: v[:R].to(dev) for k, v in synthetic.make_targets(64, seed=4).items()}
        stats = tr.train_iteration(2000, _bundle(rays, dev, slice(0, R)), targets)
        assert np.isfinite(float(stats["loss"]))
