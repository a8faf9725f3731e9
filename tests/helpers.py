"""Shared helpers for the parity tests: build the product model from the oracle's state, error metrics."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from cropnerf_b200 import synthetic  # noqa: E402
from cropnerf_b200.fruit_nerf import FruitModel, FruitNerfModelConfig  # noqa: E402
from cropnerf_b200.rays import RayBundle  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")


def rel_err(a, b, floor=1e-3):
    """elementwise |a-b| / (|b| + floor) as a numpy array (floor keeps near-zero references from blowing up)."""
    a = np.asarray(a.detach().cpu() if isinstance(a, torch.Tensor) else a, dtype=np.float64)
    b = np.asarray(b.detach().cpu() if isinstance(b, torch.Tensor) else b, dtype=np.float64)
    return np.abs(a - b) / (np.abs(b) + floor)


def assert_close(a, b, rtol, name, floor=1e-3, frac=1.0):
    """all (or at least `frac`) of the elements within rtol relative error."""
    e = rel_err(a, b, floor)
    ok = (e <= rtol).mean()
    assert ok >= frac, f"{name}: only {ok*100:.3f}% within rtol={rtol} (max {e.max():.3e}, mean {e.mean():.3e})"
    return e


def assert_close_scaled(a, b, tol, name):
    """max |a-b| <= tol * max |b|  (for quantities with cancellation, e.g. dot products / gradients)."""
    a = np.asarray(a.detach().cpu() if isinstance(a, torch.Tensor) else a, dtype=np.float64)
    b = np.asarray(b.detach().cpu() if isinstance(b, torch.Tensor) else b, dtype=np.float64)
    scale = np.abs(b).max() + 1e-30
    err = np.abs(a - b).max() / scale
    assert err <= tol, f"{name}: max abs err / max |ref| = {err:.3e} > {tol}"
    return err


def product_model(oracle_cfg, state, num_images, dev, training, precision="fp32", test_mode="val"):
    kw = {k: getattr(oracle_cfg, k) for k in oracle_cfg.__dataclass_fields__ if k in FruitNerfModelConfig.__dataclass_fields__}
    cfg = FruitNerfModelConfig(**kw)
    cfg.precision = precision
    model = FruitModel(cfg, num_train_data=num_images, test_mode=test_mode)
    missing, unexpected = model.load_state_dict(state, strict=False)
    assert not [m for m in missing if "hash_table" in m or "weight" in m or "bias" in m], missing
    model = model.to(dev)
    model.train(training)
    return model


def product_bundle(rays, dev, near_far=None):
    rb = RayBundle(
        origins=rays["origins"].to(dev),
        directions=rays["directions"].to(dev),
        pixel_area=rays["pixel_area"].to(dev),
        camera_indices=rays["camera_indices"].to(dev),
    )
    if near_far is not None:
        rb.nears = torch.full_like(rb.pixel_area, near_far[0])
        rb.fars = torch.full_like(rb.pixel_area, near_far[1])
    return rb
