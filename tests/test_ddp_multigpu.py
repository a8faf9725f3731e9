"""Peer-memory data parallelism (csrc/ddp_p2p.cu): host-side slicing on the CPU; the kernels on >= 2 GPUs (skipped on a 1-GPU box)
by running tests/ddp_p2p_check.py under torchrun."""
import json
import os
import subprocess
import sys

import pytest
import torch

import helpers  # noqa: F401
from cropnerf_b200 import ddp


def test_owned_ranges_partition_the_group():
    import ctypes as C

    from cropnerf_b200 import _lib as L

    for n in (0, 4, 64, 1_000_000 + 192, 16 * (1 << 19) * 2 + 17408 + 9600):
        for world in (1, 2, 3, 8, 16):
            spans = [ddp.owned_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n // 4 * 4
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            assert all(lo % 4 == 0 and hi % 4 == 0 for lo, hi in spans)  # whole float4s
            for r, (lo, hi) in enumerate(spans):  # the C entry computes the same split
                a, b = C.c_int64(), C.c_int64()
                L.lib().cnb_p2p_owned_range(n, r, world, C.byref(a), C.byref(b))
                assert (a.value, b.value) == (lo, hi)


def test_p2p_entry_points_validate_arguments():
    import ctypes as C

    from cropnerf_b200 import _lib as L

    comm = L.P2PComm()
    comm.world, comm.rank = 2, 5
    assert L.lib().cnb_p2p_barrier(C.byref(comm), None) == -1 and "rank" in L.last_error()
    comm.rank = 0
    assert L.lib().cnb_p2p_barrier(C.byref(comm), None) == -1 and "null" in L.last_error()


@pytest.mark.gpu
def test_peer_memory_optimizer_step_on_two_gpus():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs (the driver's multi-GPU tier / gpurun --gpus 2 runs it)")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1", "--master-port", "29577",
           os.path.join(root, "tests", "ddp_p2p_check.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env={**os.environ, "CNB_CHECK_FAST": "0"})
    line = [ln for ln in out.stdout.splitlines() if ln.startswith("DDP_P2P_CHECK ")]
    assert out.returncode == 0 and line, out.stdout[-2000:] + out.stderr[-4000:]
    summary = json.loads(line[-1][len("DDP_P2P_CHECK "):])
    assert summary["world"] == 2 and summary["kernel_max_rel_err"] < 2e-5
