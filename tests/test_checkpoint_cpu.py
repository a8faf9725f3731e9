"""nerfstudio checkpoint compatibility ("next" row f4): the reference's module names load name for name, DDP / pipeline
prefixes are stripped, tcnn-trained files are refused, and save -> resume restores weights and Adam moments.  No GPU."""
import os

import pytest
import torch

import helpers  # noqa: F401  (sys.path)
from cropnerf_b200 import checkpoint, engine
from cropnerf_b200.fruit_nerf import FruitModel, FruitNerfModelConfig
from oracle import cases


def _model(num_images=6):
    cfg = cases.make_config(dict(log2_hashmap_size=8))
    kw = {k: getattr(cfg, k) for k in cfg.__dataclass_fields__ if k in FruitNerfModelConfig.__dataclass_fields__}
    return cfg, FruitModel(FruitNerfModelConfig(**kw), num_train_data=num_images)


@pytest.mark.parametrize("prefix", ["_model.", "module._model.", "_model.module."])
def test_reference_named_checkpoint_loads(tmp_path, prefix):
    cfg, model = _model()
    oracle, state = cases.build_oracle(cfg, 6, seed=3, table_scale=0.5)  # state-dict names of the reference's modules (pinned by oracle/ref_shim.py)
    ckpt = {"step": 4000, "pipeline": {prefix + k: v for k, v in state.items()}, "optimizers": {}, "schedulers": {}, "scalers": {}}
    ckpt["pipeline"]["datamanager.train_camera_optimizer.pose_adjustment"] = torch.zeros(6, 6)  # not a model entry: ignored
    d = tmp_path / "nerfstudio_models"
    d.mkdir()
    torch.save(ckpt, str(d / checkpoint.checkpoint_name(4000)))
    torch.save({"step": 2000, "pipeline": {}}, str(d / checkpoint.checkpoint_name(2000)))
    assert checkpoint.latest_checkpoint(str(d)).endswith("step-000004000.ckpt")
    assert checkpoint.load_nerfstudio_checkpoint(model, str(d)) == 4000
    own = model.state_dict()
    for k, v in state.items():
        if k in own:
            assert torch.equal(own[k], v), k
    assert torch.equal(model.field.mlp_base_grid.hash_table, state["field.mlp_base_grid.hash_table"])
    assert torch.equal(model.proposal_networks[1].mlp_base[1].layers[0].weight, state["proposal_networks.1.mlp_base.1.layers.0.weight"])


def test_shape_mismatch_missing_and_tcnn_are_loud():
    cfg, model = _model()
    _, state = cases.build_oracle(cfg, 6, seed=3, table_scale=0.5)
    bad = dict(state)
    bad["field.embedding_appearance.embedding.weight"] = torch.zeros(7, 32)
    with pytest.raises(ValueError, match="shape"):
        checkpoint.load_nerfstudio_checkpoint(model, {"pipeline": {"_model." + k: v for k, v in bad.items()}})
    lacking = {k: v for k, v in state.items() if "mlp_head" not in k}
    with pytest.raises(KeyError, match="lacks"):
        checkpoint.load_nerfstudio_checkpoint(model, {"pipeline": {"_model." + k: v for k, v in lacking.items()}})
    tcnn = {"_model.field.mlp_base.tcnn_encoding.params": torch.zeros(100, dtype=torch.float16), "_model.field.mlp_head.tcnn_encoding.params": torch.zeros(10)}
    with pytest.raises(ValueError, match="tiny-cuda-nn"):
        checkpoint.load_nerfstudio_checkpoint(model, {"step": 1, "pipeline": tcnn})


def test_save_and_resume_restores_weights_and_adam_moments(tmp_path):
    _, model = _model()
    trainer = engine.Trainer(model, fused=False)
    g = torch.Generator().manual_seed(0)
    for grp in trainer.groups.values():
        grp.flat.copy_(torch.randn(grp.flat.shape, generator=g))
        grp.exp_avg.copy_(torch.randn(grp.flat.shape, generator=g))
        grp.exp_avg_sq.copy_(torch.rand(grp.flat.shape, generator=g))
    trainer.opt_step = 37
    path = checkpoint.save_nerfstudio_checkpoint(str(tmp_path), model, step=36, trainer=trainer)
    assert os.path.basename(path) == "step-000000036.ckpt"
    raw = torch.load(path, map_location="cpu", weights_only=False)
    assert set(raw) >= {"step", "pipeline", "optimizers", "schedulers", "scalers"}
    assert set(raw["optimizers"]) == set(trainer.groups) and all(k.startswith("_model.") for k in raw["pipeline"])
    # torch.optim.Adam accepts the stored optimizer state for the same parameter list
    opt = torch.optim.Adam(trainer.groups["fields"].params, lr=1e-2, eps=1e-15)
    opt.load_state_dict(raw["optimizers"]["fields"])
    p0 = trainer.groups["fields"].params[0]
    assert torch.equal(opt.state[p0]["exp_avg"], raw["optimizers"]["fields"]["state"][0]["exp_avg"])

    _, model2 = _model()
    trainer2 = engine.Trainer(model2, fused=False)
    step, opt_step = checkpoint.resume_trainer(trainer2, str(tmp_path))
    assert (step, opt_step) == (36, 37) and trainer2.opt_step == 37
    for name, grp in trainer.groups.items():
        grp2 = trainer2.groups[name]
        for p, p2 in zip(grp.params, grp2.params):
            assert torch.equal(p.data, p2.data)
            off = (p2.data_ptr() - grp2.flat.data_ptr()) // 4
            off1 = (p.data_ptr() - grp.flat.data_ptr()) // 4
            assert torch.equal(grp2.exp_avg[off : off + p2.numel()], grp.exp_avg[off1 : off1 + p.numel()])
            assert torch.equal(grp2.exp_avg_sq[off : off + p2.numel()], grp.exp_avg_sq[off1 : off1 + p.numel()])
        # parameters are still views of the flat buffer after loading (load_state_dict copies in place)
        assert all(p2.data_ptr() >= grp2.flat.data_ptr() and p2.data_ptr() < grp2.flat.data_ptr() + 4 * grp2.flat.numel() for p2 in grp2.params)


def test_metric_module_entries_of_a_reference_checkpoint_do_not_break_strict_loading(tmp_path):
    """The reference model owns metric modules (fruit_nerf.py:181-183: psnr, ssim, LearnedPerceptualImagePatchSimilarity); torchmetrics'
    LPIPS keeps its backbone in the state dict, so a checkpoint written by the reference trainer carries `_model.lpips.net.*` entries the
    ray-render path has no module for.  They are dropped before the strict check; unknown keys elsewhere still fail."""
    cfg, model = _model()
    _, state = cases.build_oracle(cfg, 6, seed=3, table_scale=0.5)
    pipe = {"_model." + k: v for k, v in state.items()}
    pipe["_model.lpips.net.net.slice1.0.weight"] = torch.zeros(64, 3, 11, 11)
    pipe["_model.lpips.net.lin0.model.1.weight"] = torch.zeros(1, 64, 1, 1)
    pipe["_model.psnr.dummy"] = torch.zeros(1)
    path = str(tmp_path / checkpoint.checkpoint_name(7))
    torch.save({"step": 7, "pipeline": pipe, "optimizers": {}, "schedulers": {}, "scalers": {}}, path)
    assert checkpoint.load_nerfstudio_checkpoint(model, path, strict=True) == 7        # loaded with weights_only=True
    assert torch.equal(model.field.mlp_base_grid.hash_table, state["field.mlp_base_grid.hash_table"])
    pipe["_model.field.some_new_head.weight"] = torch.zeros(3)
    with pytest.raises(KeyError, match="does not know"):
        checkpoint.load_nerfstudio_checkpoint(model, {"step": 7, "pipeline": pipe})


class _Evil:
    def __reduce__(self):
        return (print, ("pickled code ran",))


def test_checkpoint_files_are_loaded_without_unpickling_code(tmp_path):
    path = str(tmp_path / checkpoint.checkpoint_name(1))
    torch.save({"step": 1, "pipeline": {}, "evil": _Evil()}, path)
    _, model = _model()
    with pytest.raises(RuntimeError, match="weights_only"):
        checkpoint.load_nerfstudio_checkpoint(model, path, strict=False)
