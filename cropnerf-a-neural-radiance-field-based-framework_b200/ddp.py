"""Peer-memory data parallelism: the plumbing around ``csrc/ddp_p2p.cu``.

The reference's multi-GPU strategy is DDP (``fruit_pipeline.py:119-121``): gradient all-reduce + the same Adam step on every
replica.  Here the flat gradient / parameter buffers of every param group are SYMMETRIC allocations
(``torch.distributed._symmetric_memory``: one cuMem allocation per rank, mapped into every peer's address space over
NVLink / NVSwitch, plus an NVLS multicast mapping where the fabric has one), and one kernel per group does
reduce-scatter + Adam + all-gather over those mappings (``cnb_ddp_adam_update``).  torch is used for what it is here for:
allocating / exchanging the mappings and the process group; the data path is the library's own kernels.

Everything in this module needs CUDA devices with peer access; ``PeerComm.create`` raises when the rendezvous is not possible
and the caller (``engine.Trainer``) then keeps the NCCL all-reduce path.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional, Tuple

import torch
import torch.distributed as dist
from torch import Tensor

from . import _lib as L


def owned_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """[lo, hi) of the flat group's elements whose Adam state rank ``rank`` owns: whole float4s, ``ceil(n/4 / world)`` per rank
    (same arithmetic as ``cnb_p2p_owned_range``; pure Python so the host logic is testable without the library)."""
    n4 = n // 4
    q = (n4 + world - 1) // world
    a, b = min(rank * q, n4), min((rank + 1) * q, n4)
    return 4 * a, 4 * b


class PeerComm:
    """Flag blocks + barrier state of one process group, and the allocator of symmetric buffers."""

    FLAG_WORDS = 64

    def __init__(self, device: torch.device, group=None, timeout_ms: int = 10000):
        import torch.distributed._symmetric_memory as symm_mem

        self._symm = symm_mem
        self.group = group if group is not None else dist.group.WORLD
        self.world, self.rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        if self.world > L.MAX_PEERS:
            raise RuntimeError(f"PeerComm supports up to {L.MAX_PEERS} ranks")
        self.device = torch.device(device)
        self._keep: List[object] = []
        flags, hdl = self._alloc((self.FLAG_WORDS,), torch.int32)
        self.flags = flags
        self.state = torch.zeros((8,), device=self.device, dtype=torch.int32)
        c = L.P2PComm()
        c.world, c.rank, c.timeout_ms = self.world, self.rank, int(timeout_ms)
        for k in range(self.world):
            c.flags[k] = int(hdl.buffer_ptrs[k])
        c.state = self.state.data_ptr()
        c.channel = 0
        self.struct = c
        self.multicast = bool(int(hdl.multicast_ptr))

    @staticmethod
    def create(device, group=None, timeout_ms: int = 10000) -> "PeerComm":
        return PeerComm(device, group, timeout_ms)

    def _alloc(self, shape, dtype):
        """zero-filled symmetric tensor + its rendezvous handle; collective (every rank calls it in the same order)."""
        t = self._symm.empty(tuple(shape), dtype=dtype, device=self.device)
        t.zero_()
        hdl = self._symm.rendezvous(t, self.group)
        torch.cuda.current_stream(self.device).synchronize()
        hdl.barrier(channel=0)  # every rank's zero fill has landed before anyone touches a peer's buffer
        self._keep.append((t, hdl))
        return t, hdl

    def alloc_floats(self, n: int):
        t, hdl = self._alloc((n,), torch.float32)
        off = t.data_ptr() - int(hdl.buffer_ptrs[self.rank])
        ptrs = [int(p) + off for p in hdl.buffer_ptrs]
        mc = int(hdl.multicast_ptr)
        return t, ptrs, (mc + off if mc else 0)

    def barrier(self, channel: int = 0) -> None:
        """Cross-GPU barrier on the current stream.  Barriers on different channels are independent (different streams may run them concurrently)."""
        c = self.struct
        if channel != 0:
            c = L.P2PComm.from_buffer_copy(self.struct)
            c.channel = int(channel)
        L.check(L.lib().cnb_p2p_barrier(C.byref(c), L.stream_ptr(self.device)), "p2p_barrier")

    def timed_out(self) -> bool:
        """True when a barrier gave up waiting for a peer (synchronises)."""
        return bool(self.state[1::2].any().item())


class PeerGroup:
    """Peer mappings of one flat param group's gradient and parameter buffers."""

    def __init__(self, comm: PeerComm, grad_ptrs: List[int], param_ptrs: List[int], mc_grad: int, mc_param: int):
        g = L.P2PGroup()
        for k in range(comm.world):
            g.grad[k] = grad_ptrs[k]
            g.param[k] = param_ptrs[k]
        g.mc_grad = mc_grad or None
        g.mc_param = mc_param or None
        self.struct = g
        self.has_multicast = bool(mc_grad and mc_param)


def ddp_adam_update(comm: PeerComm, group: PeerGroup, exp_avg: Tensor, exp_avg_sq: Tensor, n: int, lr: float, step: int, beta1: float = 0.9,
                    beta2: float = 0.999, eps: float = 1e-15, inv_grad_scale: float = 1.0, grads_zero: bool = False, multimem: bool = False) -> None:
    flags = (L.P2P_GRADS_ZERO if grads_zero else 0) | (L.P2P_MULTIMEM if multimem else 0)
    L.check(L.lib().cnb_ddp_adam_update(C.byref(comm.struct), C.byref(group.struct), exp_avg.data_ptr(), exp_avg_sq.data_ptr(), int(n), lr, beta1, beta2, eps,
                                        int(step), inv_grad_scale, flags, L.stream_ptr(comm.device)), "ddp_adam_update")
