"""Probe: does torch symmetric memory (CUDA P2P over NVLink) work on the GPU box?  torchrun --nproc-per-node 2 tests/micro/symm_probe.py"""
import os
import sys
import time

import torch
import torch.distributed as dist
import torch.distributed._symmetric_memory as symm_mem

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ.get("LOCAL_RANK", rank))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
print(rank, "can_access_peer", [torch.cuda.can_device_access_peer(local, j) for j in range(world) if j != local], flush=True)
t = symm_mem.empty((1 << 20,), dtype=torch.float32, device=dev)
t.fill_(float(rank + 1))
hdl = symm_mem.rendezvous(t, dist.group.WORLD)
print(rank, "buffer_ptrs", [hex(p) for p in hdl.buffer_ptrs], "signal", [hex(p) for p in hdl.signal_pad_ptrs], "multicast", hex(hdl.multicast_ptr), "own data_ptr", hex(t.data_ptr()), flush=True)
hdl.barrier(channel=0)
peer = (rank + 1) % world
pb = hdl.get_buffer(peer, (1 << 20,), torch.float32)
val = float(pb[:16].sum().item()) / 16
print(rank, "peer value", val, "expected", peer + 1, flush=True)
hdl.barrier(channel=0)
# bandwidth of a plain peer read
torch.cuda.synchronize()
big = symm_mem.empty((64 << 20,), dtype=torch.float32, device=dev)
h2 = symm_mem.rendezvous(big, dist.group.WORLD)
pb2 = h2.get_buffer(peer, (64 << 20,), torch.float32)
dst = torch.empty_like(big)
h2.barrier(channel=0)
for _ in range(3):
    dst.copy_(pb2)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    dst.copy_(pb2)
e1.record(); torch.cuda.synchronize()
print(rank, "peer copy GB/s", 10 * big.numel() * 4 / (e0.elapsed_time(e1) * 1e-3) / 1e9, flush=True)
h2.barrier(channel=0)
dist.barrier()
dist.destroy_process_group()
