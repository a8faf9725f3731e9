"""CPU tests for the device-side training batch (pixel sampler + ray generator,
data/fruit_datamanager.py:188-197): the oracle restatement against hand cases and against the single-camera ray oracle, the
camera record layout the kernel reads, and loud failure without CUDA."""
import ctypes as C

import numpy as np
import pytest
import torch

from cropnerf_b200 import _lib as L
from cropnerf_b200.datamanager import DeviceTrainBatches
from cropnerf_b200.export import PinholeCamera
from oracle import nerfstudio_torch as ns


def _scene(n=3, h=6, w=8, seed=0):
    g = torch.Generator().manual_seed(seed)
    images = torch.randint(0, 256, (n, h, w, 3), generator=g, dtype=torch.uint8)
    masks = (torch.rand((n, h, w), generator=g) < 0.3).to(torch.uint8)
    c2w = torch.zeros(n, 3, 4)
    for i in range(n):
        q, _ = torch.linalg.qr(torch.randn(3, 3, generator=g))
        c2w[i, :, :3] = q
        c2w[i, :, 3] = torch.rand(3, generator=g) - 0.5
    fx = torch.full((n,), 10.0) + torch.arange(n)
    fy = torch.full((n,), 11.0) - torch.arange(n) * 0.5
    cx = torch.full((n,), w / 2.0)
    cy = torch.full((n,), h / 2.0)
    return images, masks, c2w, fx, fy, cx, cy


def test_pixel_sampler_indices_hand_cases():
    r = torch.tensor([[0.0, 0.0, 0.0], [0.999, 0.999, 0.999], [0.5, 0.5, 0.5], [0.3, 0.17, 0.126]])
    idx = ns.pixel_sampler_indices(r, 3, 6, 8)
    assert idx.tolist() == [[0, 0, 0], [2, 5, 7], [1, 3, 4], [0, 1, 1]]   # (rand * [N,H,W]) truncated toward zero
    assert idx.dtype == torch.int64


def test_next_train_batch_equals_per_camera_ray_oracle_and_pixel_lookup():
    images, masks, c2w, fx, fy, cx, cy = _scene()
    fimg = images.float() / 255.0
    rand3 = torch.rand((64, 3), generator=torch.Generator().manual_seed(1))
    idx, o, d, area, img, m = ns.next_train_batch(rand3, fimg, masks.float(), c2w, fx, fy, cx, cy)
    assert idx[:, 0].max() < 3 and idx[:, 1].max() < 6 and idx[:, 2].max() < 8
    for i in range(rand3.shape[0]):
        c, y, x = idx[i].tolist()
        o1, d1, a1 = ns.generate_pinhole_rays(c2w[c], float(fx[c]), float(fy[c]), float(cx[c]), float(cy[c]), torch.tensor([[y + 0.5, x + 0.5]]))
        assert torch.equal(o1[0], o[i]) and torch.equal(d1[0], d[i]) and torch.equal(a1[0], area[i])
        assert torch.equal(img[i], fimg[c, y, x]) and float(m[i, 0]) == float(masks[c, y, x])
    assert torch.allclose(d.norm(dim=-1), torch.ones(64), atol=1e-6)
    # uint8 storage loses nothing: value / 255 in float32 is the loader's own arithmetic
    assert torch.equal(fimg, torch.from_numpy(images.numpy().astype(np.float32) / np.float32(255.0)))


def test_camera_records_are_what_the_kernel_reads():
    """k_sample_train_batch reads every cnb_camera as 16 leading floats (c2w row-major [3,4], fx, fy, cx, cy) of a 72-byte record."""
    _, _, c2w, fx, fy, cx, cy = _scene()
    assert C.sizeof(L.Camera) == 72 and L.Camera.fx.offset == 48 and L.Camera.cy.offset == 60
    cams = (L.Camera * 3)()
    for i in range(3):
        for k, v in enumerate(c2w[i].reshape(-1).tolist()):
            cams[i].c2w[k] = v
        cams[i].fx, cams[i].fy, cams[i].cx, cams[i].cy = float(fx[i]), float(fy[i]), float(cx[i]), float(cy[i])
    raw = np.frombuffer(bytes(cams), dtype=np.float32).reshape(3, 18)
    for i in range(3):
        assert np.array_equal(raw[i, :12], c2w[i].reshape(-1).numpy())
        assert raw[i, 12:16].tolist() == [float(fx[i]), float(fy[i]), float(cx[i]), float(cy[i])]


def test_device_batches_refuse_cpu_and_bad_shapes():
    images, masks, c2w, fx, fy, cx, cy = _scene()
    cams = [PinholeCamera(c2w[i], float(fx[i]), float(fy[i]), float(cx[i]), float(cy[i]), 8, 6) for i in range(3)]
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        DeviceTrainBatches(images, masks, cams, device="cpu")
    if not torch.cuda.is_available():
        assert L.lib().cnb_sample_train_batch(None, None, 16, None, None, None, None, None, None, None, None) != 0
        assert b"sample_train_batch" in L.lib().cnb_last_error()
