"""CPU tests: the C-ABI library loads and exports every symbol ``include/cropnerf_b200.h`` declares, ctypes mirrors of
the descriptor structs have the C compiler's sizes, host-side logic (ray layout, flat param groups, schedules, the
2-rank gradient all-reduce over gloo), and the product refuses to run without CUDA (no fallback)."""
import ctypes as C
import os
import re
import subprocess
import sys
import tempfile

import pytest
import torch

from helpers import ROOT

from cropnerf_b200 import _lib as L
from cropnerf_b200 import engine, ops, synthetic
from cropnerf_b200.fruit_nerf import FruitModel, FruitNerfModelConfig
from cropnerf_b200.rays import RayBundle, ray_layout

HEADER = os.path.join(ROOT, "include", "cropnerf_b200.h")


def _declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(cnb_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = L.lib()
    names = _declared_symbols()
    assert len(names) >= 24
    for n in names:
        assert hasattr(lib, n), f"{n} declared in the header but not exported"
    assert set(names) == set(L.SIGNATURES), set(names) ^ set(L.SIGNATURES)
    assert lib.cnb_version() == 100


def test_struct_sizes_match_c_compiler():
    prog = r"""
#include <stdio.h>
#include "cropnerf_b200.h"
int main(void){ printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu\n", sizeof(cnb_grid), sizeof(cnb_mlp), sizeof(cnb_warp), sizeof(cnb_samples), sizeof(cnb_density_field), sizeof(cnb_field), sizeof(cnb_camera), sizeof(cnb_train_cfg), sizeof(cnb_opt_group), sizeof(cnb_p2p_comm), sizeof(cnb_p2p_group), sizeof(cnb_ddp_group_step), sizeof(cnb_image_set)); return 0; }
"""
    with tempfile.TemporaryDirectory() as d:
        src = os.path.join(d, "s.c")
        open(src, "w").write(prog)
        exe = os.path.join(d, "s")
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), src, "-o", exe])
        sizes = [int(v) for v in subprocess.check_output([exe]).split()]
    mine = [C.sizeof(t) for t in (L.Grid, L.Mlp, L.Warp, L.Samples, L.DensityField, L.Field, L.Camera, L.TrainCfg, L.OptGroup, L.P2PComm, L.P2PGroup, L.DdpGroupStep, L.ImageSet)]
    assert sizes == mine


def test_no_cuda_means_loud_failure_not_fallback():
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    assert L.lib().cnb_device_count() < 0
    assert "cudaGetDeviceCount" in L.last_error()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.ray_weights(torch.rand(4, 8, 1), torch.rand(4, 8, 1), torch.rand(4, 8, 1))
    model = FruitModel(FruitNerfModelConfig(log2_hashmap_size=8), num_train_data=4)
    rays = synthetic.make_rays(8, num_cameras=4)
    rb = RayBundle(rays["origins"], rays["directions"], rays["pixel_area"], rays["camera_indices"])
    with pytest.raises(RuntimeError):
        model(rb)


def test_argument_validation_reports_through_last_error():
    lib = L.lib()
    g = L.Grid()
    g.num_levels = 99
    rc = lib.cnb_hashgrid_fwd(C.byref(g), None, 0, None, None, None)
    assert rc == -1 and "hashgrid" in L.last_error()
    m = L.Mlp()
    m.num_layers = 9
    assert lib.cnb_mlp_fwd(C.byref(m), None, 0, 0, None, None, None) == -1
    assert lib.cnb_sample_spaced(None, None, None, None, 0, 0, 5, 0, None, None, None) == -1


def test_ray_layout_uses_edge_views_without_copies():
    R, S = 7, 5
    rays = synthetic.make_rays(R, num_cameras=4)
    rb = RayBundle(rays["origins"], rays["directions"], rays["pixel_area"], rays["camera_indices"])
    edges = torch.arange(R * (S + 1), dtype=torch.float32).view(R, S + 1)
    rs = rb.get_ray_samples(edges[:, :-1, None], edges[:, 1:, None])
    o, d, starts, ends, cam, r, s, stride = ray_layout(rs)
    assert (r, s, stride) == (R, S, S + 1)
    assert starts.data_ptr() == edges.data_ptr() and ends.data_ptr() == edges.data_ptr() + 4
    assert cam.dtype == torch.int32 and cam.shape == (R,)
    # density_fn-style samples: arbitrary positions -> one ray per point
    from cropnerf_b200.rays import Frustums, RaySamples

    pos = torch.rand(3, 4, 3)
    rs2 = RaySamples(Frustums(pos, torch.ones_like(pos), torch.zeros_like(pos[..., :1]), torch.zeros_like(pos[..., :1]), None))
    o2, d2, st2, en2, cam2, r2, s2, stride2 = ray_layout(rs2)
    assert (r2, s2) == (12, 1) and torch.equal(o2, pos.reshape(12, 3))


def test_flat_groups_alias_parameters_and_lr_schedule():
    model = FruitModel(FruitNerfModelConfig(log2_hashmap_size=8, proposal_net_args_list=[
        {"hidden_dim": 16, "log2_hashmap_size": 8, "num_levels": 5, "max_res": 128, "use_linear": False}] * 2), num_train_data=4)
    groups = model.get_param_groups()
    assert set(groups) == {"proposal_networks", "fields"}  # camera_opt appears only when the optimizer is on (fruit_nerf.py:191-196)
    tr = engine.Trainer(model)
    fg = tr.groups["fields"]
    n = sum(p.numel() for p in fg.params)
    assert fg.flat.numel() >= n and fg.flat.numel() == fg.grad.numel()
    assert all((q.data_ptr() - fg.flat.data_ptr()) % 256 == 0 and (q.grad.data_ptr() - fg.grad.data_ptr()) % 256 == 0 for q in fg.params)  # vector loads / red.v2 need aligned tables
    p = model.field.mlp_head.layers[0].weight
    p.data.fill_(3.0)
    assert (fg.flat == 3.0).sum() == p.numel()
    p.grad.fill_(2.0)
    assert (fg.grad == 2.0).sum() == p.numel()
    spec = engine.OptimizerSpec()
    assert abs(engine.exponential_decay_lr(0, spec) - 1e-2) < 1e-12
    assert abs(engine.exponential_decay_lr(200000, spec) - 1e-4) < 1e-12
    assert abs(engine.exponential_decay_lr(100000, spec) - 1e-3) < 1e-9


def test_proposal_update_schedule_and_anneal():
    model = FruitModel(FruitNerfModelConfig(log2_hashmap_size=8), num_train_data=4)
    s = model.proposal_sampler
    assert s.update_sched(0) == 1 and s.update_sched(5000) == 5 and s.update_sched(2500) == 2.5
    cbs = model.get_training_callbacks()
    assert len(cbs) == 2
    cbs[0].func(0)
    assert s._anneal == 0.0
    cbs[0].func(1000)
    assert abs(s._anneal - 1.0) < 1e-12
    cbs[0].func(100)
    assert abs(s._anneal - (10 * 0.1) / (9 * 0.1 + 1)) < 1e-12


_WORKER = r"""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from cropnerf_b200 import engine
from cropnerf_b200.fruit_nerf import FruitModel, FruitNerfModelConfig
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", rank=rank, world_size=world)
torch.manual_seed(0)
prop = [{"hidden_dim": 16, "log2_hashmap_size": 8, "num_levels": 5, "max_res": 128, "use_linear": False}] * 2
model = FruitModel(FruitNerfModelConfig(log2_hashmap_size=8, proposal_net_args_list=prop), num_train_data=4)
tr = engine.Trainer(model, world_size=world)
for name, g in tr.groups.items():
    g.grad.copy_(torch.arange(g.grad.numel(), dtype=torch.float32) * (rank + 1))
tr.all_reduce_gradients()
ok = True
for name, g in tr.groups.items():
    want = torch.arange(g.grad.numel(), dtype=torch.float32) * sum(r + 1 for r in range(world))
    ok = ok and torch.equal(g.grad, want)
# ray-range sharding used by export / projection: disjoint cover
from cropnerf_b200.export import shard_range
lo, hi = shard_range(1000003, rank, world)
t = torch.tensor([lo, hi, hi - lo], dtype=torch.int64)
parts = [torch.zeros_like(t) for _ in range(world)]
dist.all_gather(parts, t)
ok = ok and parts[0][0].item() == 0 and parts[-1][1].item() == 1000003 and all(parts[i][1] == parts[i + 1][0] for i in range(world - 1))
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
"""


def test_two_rank_gradient_allreduce_gloo():
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "w.py")
        open(path, "w").write(_WORKER)
        import socket

        with socket.socket() as sock:   # a free port: a fixed one can collide with another rendezvous on the same machine
            sock.bind(("127.0.0.1", 0))
            port = sock.getsockname()[1]
        procs = []
        for rank in range(2):
            env = dict(os.environ, RANK=str(rank), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
            procs.append(subprocess.Popen([sys.executable, path, ROOT], env=env))
        codes = [p.wait(timeout=300) for p in procs]
    assert codes == [0, 0]


def test_camera_optimizer_so3xr3_host_math():
    """exp_map_SO3xR3 (nerfstudio lie_groups.py as used by CameraOptimizer, fruit_nerf.py:114-116): rotation block equals the
    matrix exponential of the skew matrix, translation passes through, and apply_to_raybundle is differentiable."""
    import torch
    from cropnerf_b200.fruit_nerf import CameraOptimizer, exp_map_SO3xR3
    from cropnerf_b200.rays import RayBundle

    g = torch.Generator().manual_seed(0)
    tv = torch.randn((7, 6), generator=g) * 0.3
    m = exp_map_SO3xR3(tv)
    w = tv[:, 3:]
    skew = torch.zeros((7, 3, 3))
    skew[:, 0, 1], skew[:, 0, 2], skew[:, 1, 0], skew[:, 1, 2], skew[:, 2, 0], skew[:, 2, 1] = -w[:, 2], w[:, 1], w[:, 2], -w[:, 0], -w[:, 1], w[:, 0]
    assert torch.allclose(m[:, :3, :3], torch.matrix_exp(skew), atol=1e-5)
    assert torch.equal(m[:, :3, 3], tv[:, :3])
    opt = CameraOptimizer(5, "SO3xR3")
    with torch.no_grad():
        opt.pose_adjustment.copy_(torch.randn((5, 6), generator=g) * 0.05)
    rb = RayBundle(torch.randn((9, 3), generator=g), torch.nn.functional.normalize(torch.randn((9, 3), generator=g), dim=-1), None,
                   torch.randint(0, 5, (9, 1), generator=g))
    d0 = rb.directions.clone()
    opt.apply_to_raybundle(rb)
    assert torch.allclose(rb.directions.norm(dim=-1), d0.norm(dim=-1), atol=1e-5)  # rotations preserve length
    (rb.origins.sum() + (rb.directions * d0).sum()).backward()
    assert opt.pose_adjustment.grad is not None and opt.pose_adjustment.grad.abs().sum() > 0
    reg = {}
    opt.get_loss_dict(reg)
    assert "camera_opt_regularizer" in reg
    assert CameraOptimizer(5, "off").apply_to_raybundle(rb) is None


def test_grad_scaler_grow_and_backoff_host_logic():
    """torch.amp.GradScaler's update rule on the host side of engine.GradScaler (the flag lives wherever the gradients do; a CPU tensor here)."""
    from cropnerf_b200 import engine

    sc = engine.GradScaler(init_scale=1024.0, growth_interval=3)
    flag = sc.flag("cpu")
    for _ in range(2):
        assert sc.update() is False
    assert sc.scale == 1024.0
    assert sc.update() is False and sc.scale == 2048.0          # third clean step in a row: grow
    flag.fill_(1)
    assert sc.update() is True and sc.scale == 1024.0 and sc.skipped_steps == 1 and int(flag.item()) == 0   # inf seen: back off, flag cleared
    assert sc.update() is False and sc.growth_tracker == 1
    sd = sc.state_dict()
    sc2 = engine.GradScaler()
    sc2.load_state_dict(sd)
    assert sc2.scale == 1024.0 and sc2.growth_tracker == 1
    off = engine.GradScaler(enabled=False)
    assert off.scale == 1.0 and off.update() is False


def test_live_bitmap_bit_packing_of_nonzero_moments():
    """FlatGroup.include_nonzero_moments: units (float4s) whose Adam moments are non-zero are OR-ed into the skip bitmap, bit i of word w
    = unit 32 w + i, including bit 31 (the sign bit of the int32 word)."""
    from cropnerf_b200 import engine

    p = torch.nn.Parameter(torch.zeros(64 * 5))  # 320 floats = 80 units = 2.5 words
    g = engine.FlatGroup([p])
    n4 = g.flat.numel() // 4
    g.live = torch.zeros(((n4 + 31) // 32,), dtype=torch.int32)
    for unit in (0, 31, 32, 63, 79):
        g.exp_avg[4 * unit + 2] = 1.0
    g.exp_avg_sq[4 * 5] = 2.0
    g.include_nonzero_moments()
    bits = torch.stack([(g.live >> k) & 1 for k in range(32)], dim=1).reshape(-1)[:n4]
    assert sorted(torch.nonzero(bits)[:, 0].tolist()) == [0, 5, 31, 32, 63, 79]
    assert abs(g.live_fraction() - 6 / 80) < 1e-9
    g.live[0] |= 2  # bits already set stay set
    g.include_nonzero_moments()
    assert int((g.live[0] >> 1) & 1) == 1


def test_step_scalars_reproduce_torch_adam_and_radam():
    """engine.step_scalars: the update kernels compute p -= (s0 / s4) * m / (sqrt(v) / s5 + s3); with the scalars chosen per optimizer kind the
    same formula is torch.optim.Adam and torch.optim.RAdam (rectified and unrectified steps).  Emulated in float64 on the host."""
    import torch

    from cropnerf_b200.engine import OptimizerSpec, step_scalars

    for kind, cls in (("adam", torch.optim.Adam), ("radam", torch.optim.RAdam)):
        g = torch.Generator().manual_seed(3)
        p = torch.randn(64, generator=g, dtype=torch.float64)
        ref = torch.nn.Parameter(p.clone())
        opt = cls([ref], lr=1e-2, eps=1e-15)
        spec = OptimizerSpec(lr=1e-2, eps=1e-15, lr_final=None, kind=kind)
        m, v = torch.zeros_like(p), torch.zeros_like(p)
        for t in range(1, 15):
            gr = torch.randn(64, generator=g, dtype=torch.float64)
            s = step_scalars(spec, 1e-2, t)
            m = m + (1 - s[1]) * (gr - m)
            v = s[2] * v + (1 - s[2]) * gr * gr
            p = p - (s[0] / s[4]) * (m / (v.sqrt() / s[5] + s[3]))
            ref.grad = gr.clone()
            opt.step()
            # the scalars carry the betas as the kernels hold them (fp32-rounded): 0.999 -> 0.99900001287, hence not 1e-12
            assert float((p - ref.detach()).abs().max()) < 2e-6, (kind, t)
