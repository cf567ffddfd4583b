"""Checkpoints in the reference's on-disk format (open_flamingo/train/train_utils.py:299-375, resume at
open_flamingo/train/train.py:297-308), so runs can move between the reference trainer and this one:

    {"epoch": int, "model_state_dict": {...}, "optimizer_state_dict": {...}, "lr_scheduler_state_dict": {...}}

`model_state_dict` holds what the reference keeps: every trainable tensor under the name `model.named_parameters()`
reports it (a gated block therefore appears under `lang_encoder.<decoder path>.{i}.gated_cross_attn_layer.*`),
frozen tensors dropped unless their name contains "embed", and the duplicate views under
`lang_encoder.old_decoder_blocks`, `lang_encoder.gated_cross_attn_layers` and everything under `vision_encoder`
removed.  It loads with `load_state_dict(strict=False)` on either side (README.md:125-126).
The optimizer state of `train.FlatTrainer` is flat (one exp_avg / exp_avg_sq buffer in bucket order) and is NOT
interchangeable with torch.optim.AdamW's per-parameter state; it is stored under the same key with a format tag.
"""
import os

import torch

_DUPLICATE_PATHS = ("lang_encoder.old_decoder_blocks", "lang_encoder.gated_cross_attn_layers", "vision_encoder")


def trainable_state_dict(model):
    """`filter_state_dict_to_trainable(model, model.state_dict())` of the reference (train_utils.py:299-334)."""
    frozen = set()
    for name, p in model.named_parameters():          # de-duplicated: an aliased tensor is reported once
        if "fsdp" in name or "embed" in name or isinstance(p, torch.nn.Embedding):
            continue
        if not p.requires_grad:
            frozen.add(name.replace("._checkpoint_wrapped_module", ""))
    return {k: v for k, v in model.state_dict().items()
            if k not in frozen and not any(d in k for d in _DUPLICATE_PATHS)}


def save_checkpoint(path, model, trainer=None, epoch=0, lr_scheduler_state=None):
    """Rank-0 style save (train_utils.py:337-375).  Tensors are moved to the CPU."""
    ckpt = {"epoch": int(epoch),
            "model_state_dict": {k: v.detach().cpu().clone() for k, v in trainable_state_dict(model).items()},
            "optimizer_state_dict": None if trainer is None else flat_optimizer_state(trainer),
            "lr_scheduler_state_dict": lr_scheduler_state or {}}
    d = os.path.dirname(os.path.abspath(path))
    os.makedirs(d, exist_ok=True)
    torch.save(ckpt, path)
    return ckpt


def load_checkpoint(path_or_dict, model, trainer=None):
    """Resume as train.py:297-308 does: strip a DDP `module.` prefix, `load_state_dict(strict=False)`; returns the
    epoch to resume from (saved epoch + 1).  With `trainer`, also restores the flat optimizer state and refreshes
    the bf16 operand copies of the parameters."""
    ckpt = torch.load(path_or_dict, map_location="cpu", weights_only=False) if isinstance(path_or_dict, (str, os.PathLike)) \
        else path_or_dict
    msd = {k.replace("module.", ""): v for k, v in ckpt["model_state_dict"].items()}
    missing, unexpected = model.load_state_dict(msd, strict=False)
    if unexpected:
        raise KeyError(f"checkpoint has keys this model does not know: {sorted(unexpected)[:5]} ...")
    if trainer is not None:
        if ckpt.get("optimizer_state_dict") is not None:
            load_flat_optimizer_state(trainer, ckpt["optimizer_state_dict"])
        trainer.refresh_w16()
    return int(ckpt["epoch"]) + 1


def flat_optimizer_state(trainer):
    return {"format": "ofk-flat-adamw-v1", "step": int(trainer.step_count),
            "names": [name for name, _, _, _ in trainer.bucket.entries],
            "exp_avg": trainer.exp_avg.detach().cpu().clone(), "exp_avg_sq": trainer.exp_avg_sq.detach().cpu().clone(),
            "extra": None if trainer.extra_opt is None else trainer.extra_opt.state_dict()}


def load_flat_optimizer_state(trainer, state):
    if state.get("format") != "ofk-flat-adamw-v1":
        raise ValueError("optimizer state was not written by open_flamingo_b200.train.FlatTrainer")
    if state["names"] != [name for name, _, _, _ in trainer.bucket.entries]:
        raise ValueError("optimizer state belongs to a different parameter layout")
    trainer.exp_avg.copy_(state["exp_avg"])
    trainer.exp_avg_sq.copy_(state["exp_avg_sq"])
    trainer.step_count = int(state["step"])
    trainer.step_dev.fill_(float(trainer.step_count))
    if trainer.extra_opt is not None and state.get("extra") is not None:
        trainer.extra_opt.load_state_dict(state["extra"])
