"""pmv_b200 — B200-native MViTv2 pooling-attention block (host side).

Mirrors ``slowfast.models.attention`` / ``slowfast.models.common`` of bytedance/Portrait-Mode-Video
(MViT/ fork): ``MultiScaleBlock``, ``MultiScaleAttention``, ``Mlp``, ``DropPath`` with the reference
constructor arguments, ``forward`` contract and ``state_dict`` keys; all compute runs in the
hand-written sm_100a kernels of libpmv_b200.so through the C ABI in include/pmv_b200.h.
"""
from . import _lib  # noqa: F401

__all__ = ["_lib"]
