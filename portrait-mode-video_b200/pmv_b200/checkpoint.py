"""Loading reference (PySlowFast ``.pyth``) checkpoints into the pmv_b200 modules — SURVEY.md section 8 row f4.

Restates the PyTorch branch of ``slowfast.utils.checkpoint.load_checkpoint`` (MViT/slowfast/utils/checkpoint.py:191-563)
for the MViTv2 path: a checkpoint is ``{"model_state": state_dict, "epoch": ..., "optimizer_state": ...}``; entries whose
shape matches the model are taken as they are, ``attn.rel_pos_*`` tables of a different length are resized by 1-D
linear interpolation (:476-490, the resolution-change case of ``get_rel_pos``), everything else is reported and skipped
(``load_state_dict(strict=False)``, :540-545).  Parameter names of the pmv_b200 modules equal the reference's, so no
renaming is involved.  Host-side only; no kernels run here.
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Iterable, Optional, Union

import torch


def _clear_names(state: dict, patterns: Iterable[str]) -> dict:
    """checkpoint.py:312-328: remove the first occurrence of every pattern from the entry names."""
    for item in patterns:
        renamed = OrderedDict()
        for k, v in state.items():
            renamed[k.replace(item, "", 1) if item in k else k] = v
        state = renamed
    return state


def adapt_state_dict(pre_train: dict, model_dict: dict):
    """Returns (matched entries, names in the checkpoint that were not used) — checkpoint.py:468-520."""
    matched, not_used = {}, []
    for k, v in pre_train.items():
        if k not in model_dict:
            not_used.append(k)
            continue
        if tuple(v.shape) == tuple(model_dict[k].shape):
            matched[k] = v
        elif "attn.rel_pos" in k:
            t = v.float().t().unsqueeze(0)                                                   # :478
            t = torch.nn.functional.interpolate(t, size=model_dict[k].shape[0], mode="linear")  # :479-483
            matched[k] = t[0].t().to(v.dtype)                                                # :484
        else:
            not_used.append(k)
    return matched, not_used


def load_checkpoint(path_or_state: Union[str, dict], model: torch.nn.Module, optimizer=None, epoch_reset: bool = False,
                    clear_name_pattern: Iterable[str] = ()) -> int:
    """Loads a reference checkpoint (file path or the already loaded dict).  Returns the checkpoint's epoch, or -1 when
    it carries none / ``epoch_reset`` (checkpoint.py:547-563)."""
    ckpt = torch.load(path_or_state, map_location="cpu", weights_only=False) if isinstance(path_or_state, str) else path_or_state
    module = model.module if hasattr(model, "module") and isinstance(model.module, torch.nn.Module) else model  # :225 (DDP wrapper)
    state = ckpt["model_state"] if "model_state" in ckpt else ckpt
    state = _clear_names(OrderedDict((k[7:] if k.startswith("module.") else k, v) for k, v in state.items()), clear_name_pattern)
    model_dict = module.state_dict()
    matched, not_used = adapt_state_dict(state, model_dict)
    missing, unexpected = module.load_state_dict(matched, strict=False)
    load_checkpoint.last_report = dict(not_loaded=[k for k in model_dict if k not in matched], not_used=not_used,
                                       missing=list(missing), unexpected=list(unexpected))
    epoch = -1
    if "epoch" in ckpt and not epoch_reset:
        if optimizer is not None and "optimizer_state" in ckpt:
            optimizer.load_state_dict(ckpt["optimizer_state"])
        epoch = int(ckpt["epoch"])
    if optimizer is not None and hasattr(optimizer, "sync_low_precision"):
        optimizer.sync_low_precision()  # the weights changed under the optimizer's bf16 operand copies (fine-tuning too)
    return epoch


load_checkpoint.last_report = None
