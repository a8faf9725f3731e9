"""Row (b) of SURVEY.md section 8, the drop-in boundary: the reference's own model file runs on the product classes.

``tests/ref_dropin.py`` imports the UNMODIFIED ``/root/reference/crop_nerf/fruit_nerf/fruit_nerf.py`` with the nerfstudio names bound to
cropnerf_b200 (and ``fruit_nerf.fruit_field`` to the product's FruitField) and lets the reference's ``populate_modules`` /
``get_param_groups`` / ``get_training_callbacks`` / ``setup_inference`` / ``forward`` drive them.  Needs the reference tree (this
container); skipped where it is absent.  The numerical side of the same wiring is covered on the GPU by the product FruitModel against the
reference-executed fixtures (tests/golden/ref_*.npz, test_model_gpu.py) -- the GPU boxes have no /root/reference."""
import json
import os
import subprocess
import sys

import pytest

from oracle import ref_shim

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(not ref_shim.reference_available(), reason="/root/reference not present")
def test_reference_model_file_drives_the_product_classes():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "ref_dropin.py")], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-3000:]
    rep = json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][-1])
    assert rep["ok"] and rep["model_class"] == "fruit_nerf.fruit_nerf.FruitModel" and rep["model_file"].startswith("/root/reference/")
    assert rep["parameters"] >= 25 and ("cpu_forward" in rep or "gpu_forward" in rep)


def test_one_backend_per_process():
    """the oracle-bound and the product-bound shims cannot be mixed in one interpreter (sys.modules is global)"""
    code = ("import sys; sys.path.insert(0, %r)\nfrom oracle import ref_shim\nref_shim.install_shims('oracle')\n"
            "try:\n    ref_shim.install_shims('product')\nexcept RuntimeError as e:\n    print('refused')\n" % ROOT)
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert "refused" in out.stdout, out.stderr[-2000:]


def test_method_plugin_module_is_import_guarded():
    """`cropnerf_b200.method_config` (the fruit_nerf_config.py-shaped MethodSpecification) never raises on import: without nerfstudio it
    says why it is unavailable."""
    from cropnerf_b200 import method_config

    assert (method_config.fruit_nerf_b200_method is None) == (method_config.UNAVAILABLE is not None)
    for preset in ("fruit_nerf_b200_method_big", "fruit_nerf_b200_method_huge"):   # the two larger presets (exact fp32 mode)
        assert (getattr(method_config, preset) is None) == (method_config.UNAVAILABLE is not None)
    if method_config.fruit_nerf_b200_method is None:
        assert "nerfstudio" in method_config.UNAVAILABLE or "fruit_nerf" in method_config.UNAVAILABLE


def test_big_preset_modules_have_the_reference_shapes():
    """The FruitField dimensions of `fruit_nerf_big` / `_huge` (fruit_nerf_config.py:86-98,141-150) build the same modules in the product as the
    reference's own fruit_field.py does through the oracle shims: a strict state-dict load both ways (host side only; the kernels' side is
    tests/test_model_gpu.py::test_big_preset_model_vs_oracle)."""
    from cropnerf_b200.fruit_nerf import FruitModel, FruitNerfModelConfig
    from oracle import cases

    cfg = cases.make_config(dict(log2_hashmap_size=12, geo_feat_dim=30, hidden_dim_semantics=128, num_layers_semantic=3, max_res=4096,
                                 num_nerf_samples_per_ray=128, num_proposal_samples_per_ray=(512, 256)))
    oracle, state = cases.build_oracle(cfg, 7, seed=0, table_scale=0.5)
    kw = {k: getattr(cfg, k) for k in cfg.__dataclass_fields__ if k in FruitNerfModelConfig.__dataclass_fields__}
    model = FruitModel(FruitNerfModelConfig(**kw), num_train_data=7)
    model.load_state_dict(state, strict=True)
    shapes = {k: tuple(v.shape) for k, v in model.state_dict().items()}
    assert shapes["field.mlp_semantics.layers.0.weight"] == (128, 30) and shapes["field.mlp_semantics.layers.2.weight"] == (64, 128)
    assert shapes["field.mlp_head.layers.0.weight"] == (64, 78) and shapes["field.mlp_base_mlp.layers.1.weight"] == (31, 64)
    oracle.load_state_dict(model.state_dict(), strict=True)
