"""Mirror of stem_helper.PatchEmbed (stem_helper.py:293-325) on the pmv_b200 kernels."""
from __future__ import annotations

import torch
import torch.nn as nn

from . import functional as Fn
from .common import compute_dtype_of


class PatchEmbed(nn.Module):
    """Conv3d patchify stem.  Same constructor / state_dict (proj.weight, proj.bias) as the reference.
    ``forward_tokens`` additionally fuses the cls-token concat (video_model_builder.py:2115-2121)."""

    compute_dtype = torch.bfloat16

    def __init__(self, dim_in=3, dim_out=768, kernel=(1, 16, 16), stride=(1, 4, 4), padding=(1, 7, 7), conv_2d=False):
        super().__init__()
        if conv_2d:
            raise NotImplementedError("pmv_b200.PatchEmbed: conv_2d stems are not on the MViTv2 video path")
        self.kernel, self.stride, self.padding = tuple(kernel), tuple(stride), tuple(padding)
        self.proj = nn.Conv3d(dim_in, dim_out, kernel_size=kernel, stride=stride, padding=padding)

    def forward_tokens(self, x, cls_token):
        """clip [B, C, T, H, W] fp32 -> (tokens [B, 1+L, dim_out] fp32 with the cls token in row 0, [T', H', W'])."""
        return Fn.patch_embed(x.float(), self.proj.weight, self.proj.bias, cls_token, self.kernel, self.stride,
                              self.padding, compute_dtype_of(self))

    def forward(self, x, keep_spatial=False):
        if keep_spatial:
            raise NotImplementedError("pmv_b200.PatchEmbed: keep_spatial is not on the MViTv2 path")
        zero_cls = torch.zeros(1, 1, self.proj.out_channels, device=x.device)
        tok, thw = self.forward_tokens(x, zero_cls)
        B = x.shape[0]
        return tok[:, 1:], torch.Size([B, self.proj.out_channels, *thw])
