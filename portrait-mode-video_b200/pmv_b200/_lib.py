"""ctypes binding of libpmv_b200.so (the C ABI declared in include/pmv_b200.h).

The product path has no CPU or eager-PyTorch fallback: if the shared library is missing or a call
fails, an exception is raised.
"""
from __future__ import annotations

import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PMV_B200_LIB", os.path.join(_HERE, "libpmv_b200.so"))  # override: debug builds only

F32, BF16 = 0, 1
ACT_NONE, ACT_GELU, ACT_GELU_BWD = 0, 1, 2
GEMM_TN, GEMM_NN, GEMM_NT_REDUCE_M = 0, 1, 2

_p, _i, _i64, _f = C.c_void_p, C.c_int, C.c_int64, C.c_float
_u32, _u64 = C.c_uint32, C.c_uint64


class Epilogue(C.Structure):
    _fields_ = [("bias", _p), ("act", _i), ("aux_in", _p), ("aux_out", _p), ("ld_aux", _i64), ("row_scale", _p),
                ("rows_per_scale", _i64), ("residual", _p), ("ld_residual", _i64), ("accumulate", _i),
                ("out_group", _i64), ("out_skip", _i64)]


class PoolJob(C.Structure):
    _fields_ = [("w", _p), ("gamma", _p), ("beta", _p), ("out", _p), ("out_ld", _i64), ("dout", _p), ("dout_ld", _i64),
                ("grads", _p), ("stride_hw", _i), ("which", _i), ("xhat", _p), ("rstd", _p), ("dout_f32", _i), ("onehot", _i)]


class AdamWTensor(C.Structure):
    _fields_ = [("param", _p), ("grad", _p), ("exp_avg", _p), ("exp_avg_sq", _p), ("shadow", _p), ("numel", _i64),
                ("weight_decay", _f), ("lr_scale", _f)]


_SIGS = {
    "pmv_version": (_i, []),
    "pmv_set_pdl": (None, [_i]),
    "pmv_head_loss_fwd": (_i, [_p, _i64, _p, _p, _p, _p, _p, _f, _p, _p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _f, _p]),
    "pmv_head_loss_bwd_workspace_bytes": (_i64, [_i, _i, _i]),
    "pmv_head_loss_bwd": (_i, [_p, _p, _p, _p, _p, _p, _p, _f, _p, _p, _p, _p, _i64, _i, _p, _p, _p, _p, _p, _i, _i, _i, _p]),
    "pmv_adamw_workspace_bytes": (_i64, [_p, _i]),
    "pmv_adamw_step": (_i, [_p, _i, _p, _f, _f, _f, _f, _p, _p, _p, _p]),
    "pmv_has_tcgen05": (_i, []),
    "pmv_layernorm_fwd": (_i, [_p, _p, _p, _p, _i, _p, _p, _i64, _i, _f, _p]),
    "pmv_layernorm_bwd_workspace_bytes": (_i64, [_i64, _i]),
    "pmv_layernorm_bwd": (_i, [_p, _i, _p, _p, _p, _p, _p, _p, _p, _p, _i64, _i, _p]),
    "pmv_gemm": (_i, [_i, _p, _i64, _p, _i64, _p, _i64, _i64, _i64, _i64, _i, _i, C.POINTER(Epilogue), _i, _i, _p]),
    "pmv_colsum_workspace_bytes": (_i64, [_i64, _i64]),
    "pmv_colsum_cast": (_i, [_p, _i, _i64, _i64, _i64, _p, _i64, _p, _p, _p, _i, _i64, _p]),
    "pmv_pool_ln_fwd": (_i, [_p, _i64, _i64, _i64, _p, _p, _p, _p, _i64, _i, _i, _i, _i, _i, _i, _f, _i, _p]),
    "pmv_pool_ln_bwd_workspace_bytes": (_i64, [_i, _i, _i, _i, _i, _i]),
    "pmv_pool_ln_bwd": (_i, [_p, _i64, _i64, _i64, _p, _p, _p, _i64, _p, _p, _p, _i, _i, _i, _i, _i, _i, _f, _i, _p]),
    "pmv_pool_ln_qkv_fwd": (_i, [_p, _i64, _i64, _i64, _i64, C.POINTER(PoolJob), _i, _i, _i, _i, _i, _i, _f, _i, _p]),
    "pmv_pool_ln_qkv_bwd_workspace_bytes": (_i64, [_i, _i, _i, _i, _i, C.POINTER(_i), _i]),
    "pmv_pool_ln_qkv_bwd": (_i, [_p, _i64, _i64, _i64, _i64, C.POINTER(PoolJob), _i, _p, _p, _i, _i, _i, _i, _i, _f, _i, _p]),
    "pmv_maxpool_skip_fwd": (_i, [_p, _p, _p, _i, _i, _i, _i, _i, _p]),
    "pmv_maxpool_skip_bwd": (_i, [_p, _p, _p, _i, _i, _i, _i, _i, _p]),
    "pmv_relpos_fwd_workspace_bytes": (_i64, [_i, _i, _i, _i, _i, _i, _i]),
    "pmv_relpos_augment_q": (_i, [_p, _i64, _p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _f, _i, _i, _p]),
    "pmv_relpos_augment_k": (_i, [_p, _i64, _i, _i, _i, _i, _i, _p]),
    "pmv_relpos_bwd_workspace_bytes": (_i64, [_i, _i, _i, _i, _i, _i, _i]),
    "pmv_relpos_augment_q_bwd": (_i, [_p, _p, _i64, _p, _p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _f, _i, _i, _p]),
    "pmv_attention_fwd": (_i, [_p, _p, _i64, _i, _p, _i64, _p, _p, _p, _i, _i, _i, _i, _f, _i, _i, _i, _p]),
    "pmv_attention_bwd_workspace_bytes": (_i64, [_i, _i, _i, _i]),
    "pmv_attention_bwd": (_i, [_p, _p, _i64, _i, _p, _i64, _p, _p, _p, _p, _p, _i64, _p, _i64, _p, _i, _i, _i, _i, _f, _i, _i, _i, _p]),
    "pmv_patch_im2col": (_i, [_p, _p, _i64, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _p]),
    "pmv_probe_umma": (_i, [_p, _i, _u64, _u64, _u32, _u32, _u32, _i, _u32, _u32, _i, _p, _i, _p, _i, _p]),
    "pmv_probe_tma": (_i, [_p, _i, _u64, _u64, _u64, _u32, _u32, _i, _i, _i, _p, _i, _p]),
}

_lib = None


def exported_symbols():
    return list(_SIGS)


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python portrait-mode-video_b200/build.py` "
                "(there is no CPU / eager fallback on the product path)")
        handle = C.CDLL(LIB_PATH)
        handle.pmv_last_error.restype = C.c_char_p
        handle.pmv_last_error.argtypes = []
        for name, (res, args) in _SIGS.items():
            fn = getattr(handle, name)  # AttributeError here = stale build: rebuild the library
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def check(rc: int, what: str):
    if rc != 0:
        raise RuntimeError(f"{what} failed (code {rc}): {lib().pmv_last_error().decode()}")


def dt(t_or_dtype) -> int:
    d = t_or_dtype.dtype if isinstance(t_or_dtype, torch.Tensor) else t_or_dtype
    if d == torch.float32:
        return F32
    if d == torch.bfloat16:
        return BF16
    raise TypeError(f"pmv_b200 supports float32 / bfloat16 activations, got {d}")


def ptr(t):
    if t is None:
        return None
    assert t.is_cuda, "pmv_b200 kernels take CUDA tensors (no CPU fallback)"
    return t.data_ptr()


def stream():
    return torch.cuda.current_stream().cuda_stream
