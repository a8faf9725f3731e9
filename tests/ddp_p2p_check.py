"""Multi-GPU check of the peer-memory optimiser step (csrc/ddp_p2p.cu), run under torchrun on >= 2 GPUs:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tests/ddp_p2p_check.py

1. kernel exactness: cnb_ddp_adam_update (peer loads, and NVLS multimem when the fabric has it) against all_reduce + the torch Adam
   formula, bit-identical replicas, moments only inside the owned slice;
2. engine.Trainer(ddp="p2p") against Trainer(ddp="nccl") on the same per-rank rays;
3. device times of the fused step vs NCCL all-reduce + fused Adam at the fruit_nerf preset's sizes.
Rank 0 prints one line starting with DDP_P2P_CHECK and a JSON summary; exit code 0 = all checks passed."""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from cropnerf_b200 import ddp, engine, ops, synthetic  # noqa: E402


def adam_reference(p, g, m, v, lr, step, b1=0.9, b2=0.999, eps=1e-15):
    m = m + (1 - b1) * (g - m)
    v = b2 * v + (1 - b2) * g * g
    bc1, bc2 = 1 - b1**step, 1 - b2**step
    p = p - (lr / bc1) * (m / (v.sqrt() / bc2**0.5 + eps))
    return p, m, v


def check_kernel(comm, dev, rank, world, multimem, n=1_000_000 + 64 * 3):
    g_t, g_ptrs, mc_g = comm.alloc_floats(n)
    p_t, p_ptrs, mc_p = comm.alloc_floats(n)
    grp = ddp.PeerGroup(comm, g_ptrs, p_ptrs, mc_g, mc_p)
    gen = torch.Generator(device="cpu").manual_seed(1234)
    p0 = torch.randn((n,), generator=gen).to(dev)
    p_t.copy_(p0)
    m = torch.zeros((n,), device=dev); v = torch.zeros((n,), device=dev)
    p_ref, m_ref, v_ref = p0.clone(), torch.zeros_like(p0), torch.zeros_like(p0)
    lo, hi = ddp.owned_range(n, rank, world)
    worst = 0.0
    for step in range(1, 4):
        gr = torch.Generator(device="cpu").manual_seed(100 * step + rank)
        g_local = (torch.randn((n,), generator=gr) * 0.1).to(dev)
        g_t.copy_(g_local)
        g_sum = g_local.clone()
        dist.all_reduce(g_sum)
        p_ref, m_ref, v_ref = adam_reference(p_ref, g_sum / world, m_ref, v_ref, 1e-2, step)
        comm.barrier()
        ddp.ddp_adam_update(comm, grp, m, v, n, 1e-2, step, inv_grad_scale=1.0 / world, multimem=multimem)
        comm.barrier()
        torch.cuda.synchronize()
        err = ((p_t - p_ref).abs() / (p_ref.abs() + 1e-3)).max().item()
        worst = max(worst, err)
        assert err < 2e-5, f"params differ from all_reduce + Adam: {err}"
        assert torch.allclose(m[lo:hi], m_ref[lo:hi], rtol=1e-5, atol=1e-7) and torch.allclose(v[lo:hi], v_ref[lo:hi], rtol=1e-5, atol=1e-9)
        outside = torch.cat([m[:lo], m[hi:]])
        assert float(outside.abs().max()) == 0.0, "moments must only be maintained inside the owned slice"
    # replicas are bit-identical
    gathered = [torch.empty_like(p_t) for _ in range(world)]
    dist.all_gather(gathered, p_t.clone())
    for k in range(1, world):
        assert torch.equal(gathered[0], gathered[k]), f"replica {k} differs from replica 0"
    # a group flagged "gradient zero everywhere" still takes the momentum-only step
    g_t.fill_(123.0)  # must be ignored
    p_ref, m_ref, v_ref = adam_reference(p_ref, torch.zeros_like(p_ref), m_ref, v_ref, 1e-2, 4)
    comm.barrier()
    ddp.ddp_adam_update(comm, grp, m, v, n, 1e-2, 4, inv_grad_scale=1.0 / world, grads_zero=True, multimem=multimem)
    comm.barrier()
    torch.cuda.synchronize()
    assert ((p_t - p_ref).abs() / (p_ref.abs() + 1e-3)).max().item() < 2e-5
    return worst


def check_trainer(dev, rank, world, mode):
    from helpers import product_bundle, product_model
    from oracle import cases

    R = 256
    cfg = cases.make_config({}, small=False)
    _, state = cases.build_oracle(cfg, 20, 0, 0.5)
    rays = synthetic.make_rays(R, seed=6 + rank, num_cameras=20)  # every rank renders its own rays
    targets = {k: v.to(dev) for k, v in synthetic.make_targets(R, seed=3 + rank).items()}
    jit = synthetic.make_jitter(R, 3, seed=2 + rank)
    out = {}
    for which in ("nccl", mode):
        model = product_model(cfg, state, 20, dev, True, precision="mixed")
        feed = synthetic.JitterFeed(jit)
        model.proposal_sampler.initial_sampler.rand_fn = feed
        model.proposal_sampler.pdf_sampler.rand_fn = feed
        tr = engine.Trainer(model, world_size=world, ddp=which, cuda_graph=True)
        assert tr.ddp == which, (tr.ddp, which)
        losses = []
        for step in (3000, 3001, 3002, 3003, 3004, 3005, 3006):  # mixes proposal-update and frozen steps
            feed.reset()
            st = tr.train_iteration(step, product_bundle(rays, dev), targets)
            losses.append(float(st["loss"]))
        torch.cuda.synchronize()
        if tr.comm is not None:
            assert not tr.comm.timed_out()
            tr.gather_optimizer_state()
        out[which] = (losses, {n: g.flat.clone() for n, g in tr.groups.items()}, {n: g.exp_avg.clone() for n, g in tr.groups.items()})
    for a, b in zip(out["nccl"][0], out[mode][0]):
        assert abs(a - b) <= 2e-3 * abs(a) + 1e-6, (out["nccl"][0], out[mode][0])
    worst = 0.0
    for n in out["nccl"][1]:
        d = (out["nccl"][1][n] - out[mode][1][n]).abs().max().item()
        worst = max(worst, d)
        assert d < 8e-2, (n, d)  # 7 Adam steps of lr 1e-2 on gradients that differ by summation order
        # replicas identical
        mine = out[mode][1][n]
        ref0 = mine.clone()
        dist.broadcast(ref0, src=0)
        assert torch.equal(mine, ref0), f"{n}: replicas differ"
        me = out[mode][2][n]
        m0 = me.clone()
        dist.broadcast(m0, src=0)
        assert torch.equal(me, m0), f"{n}: gathered moments differ"
    return worst


def time_step(dev, world, comm, multimem_ok):
    """fruit_nerf preset flat-group sizes: fields 16.8 M + MLPs, proposal_networks 2.6 M."""
    sizes = {"fields": 16 * (1 << 19) * 2 + 17 * 1024 + 300 * 32, "proposal_networks": 2 * 5 * (1 << 17) * 2 + 512}
    res = {}
    bufs = {}
    for name, n in sizes.items():
        n = (n + 63) // 64 * 64
        g_t, g_ptrs, mc_g = comm.alloc_floats(n)
        p_t, p_ptrs, mc_p = comm.alloc_floats(n)
        bufs[name] = (n, g_t, p_t, ddp.PeerGroup(comm, g_ptrs, p_ptrs, mc_g, mc_p), torch.zeros((n,), device=dev), torch.zeros((n,), device=dev))
    flush = torch.empty((192 << 20,), device=dev, dtype=torch.uint8)

    def run_p2p(mm):
        comm.barrier()
        for n, g_t, p_t, grp, m, v in bufs.values():
            ddp.ddp_adam_update(comm, grp, m, v, n, 1e-2, 5, inv_grad_scale=1.0 / world, multimem=mm)
        comm.barrier()
        for n, g_t, *_ in bufs.values():
            g_t.zero_()

    def run_nccl():
        works = [dist.all_reduce(g_t, async_op=True) for n, g_t, *_ in bufs.values()]
        for w, (n, g_t, p_t, grp, m, v) in zip(works, bufs.values()):
            w.wait()
            ops.adam_step(p_t, g_t, m, v, 1e-2, 5, inv_grad_scale=1.0 / world, zero_grad=True)

    def only_barriers():
        comm.barrier(); comm.barrier()

    def only_update(mm=False):
        for n, g_t, p_t, grp, m, v in bufs.values():
            ddp.ddp_adam_update(comm, grp, m, v, n, 1e-2, 5, inv_grad_scale=1.0 / world, multimem=mm)

    def dbg(flags):
        import ctypes as C
        from cropnerf_b200 import _lib as L
        def f():
            for n, g_t, p_t, grp, m, v in bufs.values():
                L.check(L.lib().cnb_ddp_adam_update(C.byref(comm.struct), C.byref(grp.struct), m.data_ptr(), v.data_ptr(), n, 1e-2, 0.9, 0.999, 1e-15, 5, 0.5, flags,
                                                    L.stream_ptr(dev)), "dbg")
        return f

    reg = {name: [torch.zeros((b[0],), device=dev) for _ in range(4)] for name, b in bufs.items()}

    def adam_symm():
        for n, g_t, p_t, grp, m, v in bufs.values():
            ops.adam_step(p_t, g_t, m, v, 1e-2, 5, inv_grad_scale=1.0 / world, zero_grad=True)

    def adam_regular():
        for p_r, g_r, m_r, v_r in reg.values():
            ops.adam_step(p_r, g_r, m_r, v_r, 1e-2, 5, inv_grad_scale=1.0 / world, zero_grad=True)

    def only_zero():
        for n, g_t, *_ in bufs.values():
            g_t.zero_()

    def only_allreduce():
        for w in [dist.all_reduce(g_t, async_op=True) for n, g_t, *_ in bufs.values()]:
            w.wait()

    variants = {"p2p": lambda: run_p2p(False), "nccl_allreduce_plus_adam": run_nccl, "part_two_barriers": only_barriers, "part_update_kernels": only_update,
                "part_grad_clear": only_zero, "adam_zero4_symmetric_bufs": adam_symm, "adam_zero4_regular_bufs": adam_regular, "dbg_local_loads": dbg(16), "dbg_local_stores": dbg(32), "dbg_all_local": dbg(48), "mm_u1": dbg(2), "mm_u2": dbg(2 | 0x100), "mm_u4": dbg(2 | 0x200), "mm_u1_grads_zero": dbg(2 | 1), "p2p_grads_zero": dbg(1), "part_nccl_allreduce": only_allreduce}
    if multimem_ok:
        variants["p2p_multimem"] = lambda: run_p2p(True)
    for label, fn in variants.items():
        for _ in range(3):
            fn()
        torch.cuda.synchronize(); dist.barrier()
        tot = 0.0
        iters = 10
        for _ in range(iters):
            flush.fill_(1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record()
            torch.cuda.synchronize()
            tot += e0.elapsed_time(e1)
        t = torch.tensor([tot / iters], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        res[label + "_ms"] = round(float(t.item()), 4)
    return res


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    summary = {"world": world}
    comm = ddp.PeerComm.create(dev)
    summary["multicast"] = comm.multicast
    summary["kernel_max_rel_err"] = check_kernel(comm, dev, rank, world, multimem=False)
    mm_ok = False
    if comm.multicast and os.environ.get("CNB_SKIP_MULTIMEM", "0") != "1":
        try:
            summary["kernel_multimem_max_rel_err"] = check_kernel(comm, dev, rank, world, multimem=True)
            mm_ok = True
        except AssertionError as e:
            summary["kernel_multimem_error"] = str(e)[:200]
    fast = os.environ.get("CNB_CHECK_FAST", "0") == "1"
    if not fast:
        summary["trainer_p2p_vs_nccl_max_param_diff"] = check_trainer(dev, rank, world, "p2p")
    if mm_ok and not fast:
        summary["trainer_multimem_vs_nccl_max_param_diff"] = check_trainer(dev, rank, world, "p2p_multimem")
    summary.update(time_step(dev, world, comm, mm_ok))
    assert not comm.timed_out()
    dist.barrier()
    if rank == 0:
        print("DDP_P2P_CHECK " + json.dumps(summary), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
