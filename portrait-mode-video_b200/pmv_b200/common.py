"""Mirror of slowfast/models/common.py (Mlp, DropPath) on the pmv_b200 kernels."""
from __future__ import annotations

import torch
import torch.nn as nn

from . import functional as Fn


def compute_dtype_of(module) -> torch.dtype:
    return getattr(module, "compute_dtype", torch.bfloat16)


class Mlp(nn.Module):
    """fc1 -> exact-erf GELU -> fc2 (common.py:7-34).  Same constructor and state_dict keys
    (fc1.weight/bias, fc2.weight/bias).  Dropout (drop_rate > 0) is not on the MViTv2 path and is refused."""

    compute_dtype = torch.bfloat16

    def __init__(self, in_features, hidden_features=None, out_features=None, act_layer=nn.GELU, drop_rate=0.0):
        super().__init__()
        if drop_rate > 0.0:
            raise NotImplementedError("pmv_b200.Mlp: drop_rate > 0 is not supported (MViTv2 uses 0.0)")
        if act_layer is not nn.GELU:
            raise NotImplementedError("pmv_b200.Mlp: only the exact-erf nn.GELU activation is implemented")
        self.drop_rate = drop_rate
        out_features = out_features or in_features
        hidden_features = hidden_features or in_features
        self.fc1 = nn.Linear(in_features, hidden_features)
        self.act = act_layer()
        self.fc2 = nn.Linear(hidden_features, out_features)

    def forward(self, x, residual=None, row_scale=None, rows_per_scale=1):
        T = compute_dtype_of(self)
        if x.dtype != T:
            x = x.to(T)
        return Fn.mlp(x, self.fc1.weight, self.fc1.bias, self.fc2.weight, self.fc2.bias, residual, row_scale, rows_per_scale)


def drop_path_scale(batch: int, drop_prob: float, training: bool, device, dtype=torch.float32):
    """Per-sample DropPath factor mask / keep_prob (common.py:46-59), or None when it is the identity.
    Consumes the torch RNG exactly like the reference (one uniform per sample)."""
    if drop_prob == 0.0 or not training:
        return None
    keep = 1 - drop_prob
    mask = keep + torch.rand((batch,), dtype=dtype, device=device)
    mask.floor_()
    return mask / keep


class DropPath(nn.Module):
    """Stochastic depth per sample (common.py:62-70).  Inside MultiScaleBlock the factor is folded into the
    GEMM epilogue; this module form is kept for API compatibility."""

    def __init__(self, drop_prob=None):
        super().__init__()
        self.drop_prob = drop_prob

    def forward(self, x):
        s = drop_path_scale(x.shape[0], self.drop_prob or 0.0, self.training, x.device, x.dtype)
        if s is None:
            return x
        return x * s.view((x.shape[0],) + (1,) * (x.ndim - 1))
