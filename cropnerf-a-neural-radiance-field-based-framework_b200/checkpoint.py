"""nerfstudio checkpoint compatibility ("next" row f4 of SURVEY.md section 8; SURVEY.md section 5 "Checkpoint / resume").

The reference trains and evaluates through the nerfstudio Trainer: ``steps_per_save=2000`` (``fruit_nerf_config.py:33``)
writes ``nerfstudio_models/step-{step:09d}.ckpt`` and ``eval_setup(load_config)`` (``export/exporter_nerfacto.py:105``,
``scripts/semantic_projection.py:139-143``) loads the latest one.  Such a file is a ``torch.save`` of::

    {"step": int,
     "pipeline":   pipeline.state_dict()      # model entries are prefixed "_model." ("module._model." / "_model.module." under DDP)
     "optimizers": {group: torch.optim.Adam.state_dict()},
     "schedulers": {group: scheduler.state_dict()},
     "scalers":    GradScaler.state_dict()}

The B200 modules keep the reference's module / parameter names (``field.mlp_base_grid.hash_table``,
``field.mlp_base_mlp.layers.*``, ``proposal_networks.{0,1}.*`` ..., ``fruit_field.py:99-167``), so a checkpoint trained with
``implementation="torch"`` loads name for name.  A checkpoint trained with tiny-cuda-nn (``implementation="tcnn"``, the
default where tcnn is installed) stores one opaque fp16-trained ``tcnn_encoding.params`` blob per module and encodes a
DIFFERENT function (dense coarse levels, +0.5 cell offset, bias-free padded MLPs): it is detected and
refused with an explanation instead of being loaded into kernels that would render something else.
"""
from __future__ import annotations

import os
import re
from typing import Dict, Optional, Tuple

import torch
from torch import Tensor

_MODEL_PREFIXES = ("module._model.module.", "module._model.", "_model.module.", "_model.")


def checkpoint_name(step: int) -> str:
    """nerfstudio Trainer.save_checkpoint file name."""
    return f"step-{step:09d}.ckpt"


def latest_checkpoint(load_dir: str) -> str:
    """``eval_setup`` picks the highest step in ``nerfstudio_models/`` (eval_utils.eval_load_checkpoint)."""
    steps = []
    for name in os.listdir(load_dir):
        m = re.fullmatch(r"step-(\d+)\.ckpt", name)
        if m:
            steps.append(int(m.group(1)))
    if not steps:
        raise FileNotFoundError(f"no step-*.ckpt under {load_dir}")
    return os.path.join(load_dir, checkpoint_name(max(steps)))


def model_state_from_pipeline_state(pipeline_state: Dict[str, Tensor]) -> Dict[str, Tensor]:
    """Strip the pipeline / DDP prefixes; entries that do not belong to the model (datamanager state) are dropped."""
    out: Dict[str, Tensor] = {}
    for key, value in pipeline_state.items():
        for prefix in _MODEL_PREFIXES:
            if key.startswith(prefix):
                out[key[len(prefix):]] = value
                break
    return out


def _refuse_tcnn(state: Dict[str, Tensor]) -> None:
    bad = [k for k in state if k.endswith("tcnn_encoding.params") or k.endswith(".params") and ("mlp_base" in k or "mlp_head" in k or "encoding" in k)]
    if bad:
        raise ValueError(
            "this checkpoint was trained with tiny-cuda-nn modules (" + ", ".join(sorted(bad)[:3]) + ", ...): tcnn's hash grid "
            "(dense coarse levels, +0.5 cell offset, fp16 tables) and bias-free FullyFusedMLP are a different function from "
            "nerfstudio's torch implementation that this library reproduces; retrain / export with implementation='torch'")


# state-dict prefixes of modules the reference model owns that are NOT part of the ray-render path (fruit_nerf.py:181-183 builds
# psnr / ssim / LearnedPerceptualImagePatchSimilarity metric modules; torchmetrics' LPIPS carries its VGG/Alex weights in the state dict)
_NON_PRODUCT_PREFIXES = ("lpips.", "psnr.", "ssim.", "rgb_loss.", "binary_cross_entropy_loss.", "cross_entropy_loss.")


def _load_file(path: str, allow_pickle: bool):
    """nerfstudio checkpoints hold dicts, tensors and scalars only: load with ``weights_only=True``; full unpickling (arbitrary code
    from the file) only on the caller's explicit ``allow_pickle=True``."""
    try:
        return torch.load(path, map_location="cpu", weights_only=True)
    except Exception as e:  # pickle.UnpicklingError and friends
        if not allow_pickle:
            raise RuntimeError(f"{path}: not loadable with weights_only=True ({type(e).__name__}: {e}); pass allow_pickle=True only for files you trust") from e
        return torch.load(path, map_location="cpu", weights_only=False)


def load_nerfstudio_checkpoint(model, path_or_state, strict: bool = True, allow_pickle: bool = False) -> int:
    """Load a nerfstudio ``step-*.ckpt`` (path, directory holding them, or the already loaded dict) into a B200 ``FruitModel``.
    Returns the training step stored in the file.  Shapes must match exactly (``num_train_data`` / ``log2_hashmap_size`` of
    the model must be the checkpoint's); with ``strict`` every learnable tensor of the model must be present."""
    if isinstance(path_or_state, (str, os.PathLike)):
        path = str(path_or_state)
        if os.path.isdir(path):
            path = latest_checkpoint(path)
        loaded = _load_file(path, allow_pickle)
    else:
        loaded = path_or_state
    pipeline_state = loaded["pipeline"] if "pipeline" in loaded else loaded
    state = model_state_from_pipeline_state(pipeline_state) if any(k.startswith(_MODEL_PREFIXES) for k in pipeline_state) else dict(pipeline_state)
    _refuse_tcnn(state)
    state = {k: v for k, v in state.items() if not k.startswith(_NON_PRODUCT_PREFIXES)}  # metric modules of the reference model
    own = model.state_dict()
    for key, value in state.items():
        if key in own and tuple(own[key].shape) != tuple(value.shape):
            raise ValueError(f"checkpoint tensor {key} has shape {tuple(value.shape)}, the model expects {tuple(own[key].shape)}")
    # the modules register nerfstudio's aliases of one tensor (mlp_base.0 == mlp_base_grid, encoding == mlp_base.0): fill in
    # whichever alias a checkpoint lacks so either spelling loads
    for key in own:
        if key not in state:
            for a, b in ((".mlp_base.0.", ".mlp_base_grid."), (".mlp_base.1.", ".mlp_base_mlp."), (".mlp_base.0.", ".encoding.")):
                for src, dst in ((a, b), (b, a)):
                    if dst in key and key.replace(dst, src) in state:
                        state[key] = state[key.replace(dst, src)]
    missing, unexpected = model.load_state_dict(state, strict=False)
    learnable = {n for n, _ in model.named_parameters()}
    really_missing = [k for k in missing if k in learnable]
    if strict and really_missing:
        raise KeyError(f"checkpoint lacks model parameters: {really_missing[:8]}")
    if strict and unexpected:
        raise KeyError(f"checkpoint has tensors the model does not know: {list(unexpected)[:8]}")
    return int(loaded.get("step", 0)) if isinstance(loaded, dict) else 0


def _adam_state_dict(group, spec, opt_step: int) -> dict:
    """torch.optim.Adam.state_dict() layout for one flat group (per-parameter views of the flat moment buffers)."""
    state = {}
    base = group.flat.data_ptr()
    for i, p in enumerate(group.params):
        off = (p.data_ptr() - base) // 4
        n = p.numel()
        state[i] = {"step": torch.tensor(float(opt_step)), "exp_avg": group.exp_avg[off : off + n].view(p.shape).detach().cpu().clone(),
                    "exp_avg_sq": group.exp_avg_sq[off : off + n].view(p.shape).detach().cpu().clone()}
    pg = {"lr": spec.lr, "betas": tuple(spec.betas), "eps": spec.eps, "weight_decay": 0, "amsgrad": False, "maximize": False, "foreach": None,
          "capturable": False, "differentiable": False, "fused": None, "params": list(range(len(group.params)))}
    return {"state": state, "param_groups": [pg]}


def save_nerfstudio_checkpoint(directory: str, model, step: int, trainer=None) -> str:
    """Write ``step-{step:09d}.ckpt`` in the nerfstudio layout (same nesting and key names as the reference trainer's files; read back by
    :func:`load_nerfstudio_checkpoint`.  The reference model's metric modules -- ``lpips.net.*`` etc. -- are not written, so a reference
    install has to load it with ``strict=False``).  With a :class:`engine.Trainer` the Adam moments of every
    flat group are stored in ``torch.optim.Adam.state_dict()`` form so training can resume."""
    os.makedirs(directory, exist_ok=True)
    pipeline_state = {"_model." + k: v.detach().cpu().clone() for k, v in model.state_dict().items()}
    ckpt = {"step": int(step), "pipeline": pipeline_state, "optimizers": {}, "schedulers": {}, "scalers": {}}
    if trainer is not None:
        if hasattr(trainer, "gather_optimizer_state"):
            trainer.gather_optimizer_state()  # peer-memory data parallelism shards the moments across ranks (collective call)
        for name, group in trainer.groups.items():
            ckpt["optimizers"][name] = _adam_state_dict(group, trainer.optimizers[name], trainer.opt_step)
            ckpt["schedulers"][name] = {"last_epoch": int(step), "_step_count": int(step) + 1}
        if getattr(trainer, "grad_scaler", None) is not None:
            ckpt["scalers"] = trainer.grad_scaler.state_dict()
    path = os.path.join(directory, checkpoint_name(step))
    torch.save(ckpt, path)
    return path


def resume_trainer(trainer, path_or_state, allow_pickle: bool = False) -> Tuple[int, Optional[int]]:
    """Load model weights AND Adam moments back into a :class:`engine.Trainer` (flat groups).  Returns (step, opt_step)."""
    if isinstance(path_or_state, (str, os.PathLike)):
        path = str(path_or_state)
        if os.path.isdir(path):
            path = latest_checkpoint(path)
        loaded = _load_file(path, allow_pickle)
    else:
        loaded = path_or_state
    step = load_nerfstudio_checkpoint(trainer.model, loaded, strict=True)  # param.data are views of the flat buffers: copied in place
    opt_step = None
    for name, sd in loaded.get("optimizers", {}).items():
        group = trainer.groups.get(name)
        if group is None:
            continue
        if len(sd["state"]) not in (0, len(group.params)):
            raise ValueError(f"optimizer group {name}: checkpoint has {len(sd['state'])} parameter states, the model has {len(group.params)}")
        base = group.flat.data_ptr()
        for i, p in enumerate(group.params):
            st = sd["state"].get(i)
            if st is None:
                continue
            off = (p.data_ptr() - base) // 4
            n = p.numel()
            group.exp_avg[off : off + n].copy_(st["exp_avg"].reshape(-1))
            group.exp_avg_sq[off : off + n].copy_(st["exp_avg_sq"].reshape(-1))
            opt_step = int(float(st["step"]))
    for group in trainer.groups.values():  # moments that arrived non-zero keep their units live in the optimiser's skip bitmap
        group.include_nonzero_moments()
    if opt_step is not None:
        trainer.opt_step = opt_step
    if getattr(trainer, "grad_scaler", None) is not None:
        trainer.grad_scaler.load_state_dict(loaded.get("scalers", {}))
    return step, opt_step
