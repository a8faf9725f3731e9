"""GPU parity of the device-side training batch (csrc/train_batch.cu, row f1: FruitDataManager.next_train,
data/fruit_datamanager.py:188-197) against the oracle: indices and gathered pixels bit-exact, ray directions to 2e-6 and pixel areas to
2e-3 relative (fused multiply-adds in the norm / pixel-area expressions; the bounds of test_generate_rays_and_aabb_clip).

First run on a B200 in round 2 (gpurun_out/r2_trainbatch.log: 2 passed); the round-1 gate is gone."""
import pytest
import torch

from oracle import nerfstudio_torch as ns

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("storage", ["uint8", "float32"])
def test_device_train_batch_matches_oracle(storage):
    from cropnerf_b200.datamanager import DeviceTrainBatches
    from cropnerf_b200.export import PinholeCamera

    g = torch.Generator().manual_seed(0)
    n, h, w, R = 7, 54, 96, 4096
    images = torch.randint(0, 256, (n, h, w, 3), generator=g, dtype=torch.uint8)
    masks = (torch.rand((n, h, w), generator=g) < 0.1).to(torch.uint8)
    c2w = torch.zeros(n, 3, 4)
    for i in range(n):
        q, _ = torch.linalg.qr(torch.randn(3, 3, generator=g))
        c2w[i, :, :3] = q
        c2w[i, :, 3] = torch.rand(3, generator=g) - 0.5
    fx = torch.full((n,), 72.0) + torch.arange(n)
    fy = torch.full((n,), 71.0) - torch.arange(n) * 0.5
    cx, cy = torch.full((n,), w / 2.0), torch.full((n,), h / 2.0)
    rand3 = torch.rand((R, 3), generator=g)
    rand3[0] = torch.tensor([0.0, 0.0, 0.0])
    rand3[1] = torch.tensor([1.0 - 2.0**-24] * 3)          # largest float32 below 1: must stay in range
    cams = [PinholeCamera(c2w[i], float(fx[i]), float(fy[i]), float(cx[i]), float(cy[i]), w, h) for i in range(n)]
    fimg = images.float() / 255.0
    dm = DeviceTrainBatches(images if storage == "uint8" else fimg, masks, cams, num_rays_per_batch=R, device="cuda:0",
                            rand_fn=lambda shape, dev: rand3.to(dev))
    bundle, batch = dm.next_train(0)
    torch.cuda.synchronize()
    idx, o, d, area, img, m = ns.next_train_batch(rand3, fimg, masks.float(), c2w, fx, fy, cx, cy)
    idx = torch.minimum(idx, torch.tensor([n - 1, h - 1, w - 1]))   # the kernel clamps where torch would raise
    assert torch.equal(batch["indices"].cpu().long(), idx)
    assert torch.equal(bundle.camera_indices.cpu().long()[:, 0], idx[:, 0])
    assert torch.equal(batch["image"].cpu(), fimg[idx[:, 0], idx[:, 1], idx[:, 2]])
    assert torch.equal(batch["fruit_mask"].cpu()[:, 0], masks.float()[idx[:, 0], idx[:, 1], idx[:, 2]])
    assert torch.equal(bundle.origins.cpu(), c2w[idx[:, 0]][:, :3, 3])
    if not torch.equal(idx, ns.pixel_sampler_indices(rand3, n, h, w)):
        keep = (idx == ns.pixel_sampler_indices(rand3, n, h, w)).all(-1)
        d, area = d[keep], area[keep]
        got_d, got_a = bundle.directions.cpu()[keep], bundle.pixel_area.cpu()[keep]
    else:
        got_d, got_a = bundle.directions.cpu(), bundle.pixel_area.cpu()
    assert (got_d - d).abs().max() <= 2e-6                                  # tolerances of test_generate_rays_and_aabb_clip
    assert ((got_a - area).abs() <= 2e-3 * area.abs() + 1e-12).all()       # product of two differences of nearly equal unit vectors
