"""Row f3: the optimizer step of the reference recipe as one multi-tensor CUDA pass.

Reference: ``slowfast/models/optimizer.py:14-131`` (``construct_optimizer``: parameter groups with zero weight decay for
BatchNorm, 1-D parameters / biases (``SOLVER.ZERO_WD_1D_PARAM``) and the names in ``model.no_weight_decay()``;
``torch.optim.AdamW(eps=1e-8)``) and ``tools/train_net.py:190-199`` (``clip_grad_norm_(CLIP_GRAD_L2NORM)`` before the
step).  ``MVITv2_S_16x4.yaml:62-75``: AdamW, WEIGHT_DECAY 0.05, ZERO_WD_1D_PARAM True, CLIP_GRAD_L2NORM 1.0.

``FusedAdamW.step()`` = [global gradient norm ->] clip coefficient + bias corrections -> AdamW update of every tensor,
which also rewrites the bf16 operand copy of each matrix weight (``weight_lp``: what the GEMMs read), so the next
forward launches no cast kernels.  Learning rate and step count live on the device: a captured CUDA graph follows the
LR schedule (``set_lr``) without re-capture.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Iterable, List, Optional

import torch

from . import _lib as L
from . import ops


def param_groups(model: torch.nn.Module, weight_decay: float, zero_wd_1d: bool = True, bn_weight_decay: float = 0.0) -> List[dict]:
    """The three groups of optimizer.py:29-83 (LAYER_DECAY == 1.0 branch), in the reference's order."""
    skip = model.no_weight_decay() if hasattr(model, "no_weight_decay") else {}
    bn, non_bn, zero = [], [], []
    for name_m, m in model.named_modules():
        is_bn = isinstance(m, torch.nn.modules.batchnorm._NormBase)
        for name_p, p in m.named_parameters(recurse=False):
            name = "{}.{}".format(name_m, name_p).strip(".")
            if not p.requires_grad:
                continue
            if is_bn:
                bn.append(p)
            elif any(k in name for k in skip):
                zero.append(p)
            elif zero_wd_1d and (p.dim() == 1 or name.endswith(".bias")):
                zero.append(p)
            else:
                non_bn.append(p)
    groups = [dict(params=bn, weight_decay=bn_weight_decay), dict(params=non_bn, weight_decay=weight_decay),
              dict(params=zero, weight_decay=0.0)]
    return [g for g in groups if len(g["params"])]


def low_precision_weight(w: torch.Tensor, dtype: torch.dtype) -> Optional[torch.Tensor]:
    """The operand copy maintained by FusedAdamW for `w`, if it is current (the parameter has not been modified by
    anything else since: torch bumps ``_version`` on every in-place write)."""
    sh = getattr(w, "_pmv_lp", None)
    if sh is not None and sh.dtype == dtype and getattr(w, "_pmv_lp_version", -1) == w._version:
        return sh
    return None


class FusedAdamW:
    """torch.optim.AdamW semantics (amsgrad=False, maximize=False) over parameter groups ``[{params, weight_decay[, lr_scale]}]``.

    max_grad_norm: clip the global L2 norm of all gradients first (None: no clipping); ``grad_norm`` then holds the
    pre-clip norm (device scalar), like the value ``clip_grad_norm_`` returns.
    lp_dtype: keep an operand copy in this dtype for every parameter with >= 2 dimensions (None: no copies).
    """

    def __init__(self, params, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 1e-2,
                 max_grad_norm: Optional[float] = None, lp_dtype: Optional[torch.dtype] = torch.bfloat16):
        params = list(params)
        if params and not isinstance(params[0], dict):
            params = [dict(params=params)]
        self.param_groups = []
        for g in params:
            g = dict(g)
            g["params"] = list(g["params"])
            g.setdefault("weight_decay", weight_decay)
            g.setdefault("lr_scale", 1.0)
            self.param_groups.append(g)
        self.betas, self.eps = (float(betas[0]), float(betas[1])), float(eps)
        self.max_grad_norm = float(max_grad_norm) if max_grad_norm else 0.0
        self._flat = [(p, g) for g in self.param_groups for p in g["params"] if p.requires_grad]
        assert self._flat, "no parameters"
        dev = self._flat[0][0].device
        assert dev.type == "cuda", "FusedAdamW runs on the GPU only (no CPU fallback)"
        self.device = dev
        self.lr = torch.full((1,), float(lr), dtype=torch.float32, device=dev)
        self.step_count = torch.zeros(1, dtype=torch.int32, device=dev)
        self.grad_norm = torch.zeros(1, dtype=torch.float32, device=dev)
        self.state: Dict[torch.Tensor, dict] = {}
        for p, _ in self._flat:
            assert p.dtype == torch.float32 and p.is_contiguous(), "fp32 contiguous master parameters expected"
            st = dict(exp_avg=torch.zeros_like(p), exp_avg_sq=torch.zeros_like(p))
            if lp_dtype is not None and p.dim() >= 2:
                assert lp_dtype == torch.bfloat16
                st["lp"] = p.detach().to(lp_dtype)
                p._pmv_lp = st["lp"]
                p._pmv_lp_version = p._version
            self.state[p] = st
        self._arr = (L.AdamWTensor * len(self._flat))()
        for i, (p, g) in enumerate(self._flat):
            st = self.state[p]
            lp = st.get("lp")
            self._arr[i] = L.AdamWTensor(p.data_ptr(), None, st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr(),
                                         lp.data_ptr() if lp is not None else None, p.numel(), float(g["weight_decay"]),
                                         float(g["lr_scale"]))
        nbytes = L.lib().pmv_adamw_workspace_bytes(self._arr, len(self._flat))
        self._ws = torch.empty(nbytes // 4, dtype=torch.float32, device=dev)
        self._numel = sum(p.numel() for p, _ in self._flat)
        self.arena = ops.attach_arena([p for p, _ in self._flat])  # split-K weight-gradient memory of THIS model

    # ------------------------------------------------------------------ schedule
    def set_lr(self, lr: float):
        """Device-side learning rate (lr_policy.py: per-iteration cosine / warm-up values go through here)."""
        self.lr.fill_(float(lr))

    def zero_grad(self, set_to_none: bool = True):
        if set_to_none:
            self.arena.reset(self.device)  # last step's gradients are dead: one fill for all split-K weight gradients
        else:
            self.arena.exhaust()  # the gradients stay alive (some inside the arena): hand nothing out until they are dropped
        for p, _ in self._flat:
            if set_to_none:
                p.grad = None
            elif p.grad is not None:
                p.grad.zero_()

    def sync_low_precision(self):
        """Re-derive the operand copies after the parameters were changed by anything but step() (checkpoint load)."""
        for p, _ in self._flat:
            lp = self.state[p].get("lp")
            if lp is not None:
                lp.copy_(p.detach())
                p._pmv_lp_version = p._version

    # ------------------------------------------------------------------ step
    @torch.no_grad()
    def step(self):
        for i, (p, _) in enumerate(self._flat):
            g = p.grad
            assert g is not None, "a parameter has no gradient (find_unused_parameters=False semantics)"
            assert g.dtype == torch.float32 and g.is_contiguous() and g.device == self.device
            self._arr[i].grad = g.data_ptr()
        n = len(self._flat)
        nlaunch = 1 + (-(-n // 320)) * 2  # prep + (norm, update) per 320 tensors (+ a 4-byte copy node for grad_norm)
        ops._run("pmv_adamw_step", nlaunch, dict(bytes=self._numel * 34), self._arr, n, L.ptr(self.lr), self.betas[0], self.betas[1],
                 self.eps, self.max_grad_norm, L.ptr(self.step_count), L.ptr(self.grad_norm), L.ptr(self._ws), L.stream())
        for p, _ in self._flat:  # the kernel regenerated every operand copy from p: they are current whatever happened before
            if "lp" in self.state[p]:
                p._pmv_lp_version = p._version

    # ------------------------------------------------------------------ checkpointing: torch.optim.AdamW's layout
    def state_dict(self):
        """``torch.optim.AdamW.state_dict()`` layout — ``{"state": {idx: {step, exp_avg, exp_avg_sq}}, "param_groups":
        [{lr, betas, eps, weight_decay, ..., params: [idx, ...]}]}`` with idx running over the groups in order — so that a
        checkpoint written here resumes in the reference (utils/checkpoint.py:130-160 saves ``optimizer.state_dict()``) and
        the reference's ``optimizer_state`` resumes here (:547-553)."""
        step = float(self.step_count.item())
        lr = float(self.lr.item())
        state, groups, idx = {}, [], 0
        for g in self.param_groups:
            ids = []
            for p in g["params"]:
                if not p.requires_grad:
                    continue
                st = self.state[p]
                state[idx] = dict(step=torch.tensor(step), exp_avg=st["exp_avg"].clone(), exp_avg_sq=st["exp_avg_sq"].clone())
                ids.append(idx)
                idx += 1
            groups.append(dict(lr=lr * float(g["lr_scale"]), betas=self.betas, eps=self.eps, weight_decay=float(g["weight_decay"]),
                               amsgrad=False, maximize=False, foreach=None, capturable=False, differentiable=False, fused=None,
                               lr_scale=float(g["lr_scale"]), params=ids))
        return dict(state=state, param_groups=groups)

    def load_state_dict(self, sd):
        """Accepts the torch.optim layout (above; what a reference ``.pyth`` checkpoint carries) and the private layout
        of round 1 (``{"step", "lr", "state": [...]}``).  Moments are matched by position: the reference numbers the
        parameters group by group in ``construct_optimizer`` order, which ``param_groups()`` reproduces."""
        if "param_groups" in sd:  # torch.optim layout
            groups = sd["param_groups"]
            ids = [i for g in groups for i in g["params"]]
            if len(ids) != len(self._flat):
                raise ValueError(f"optimizer state has {len(ids)} parameters, this optimizer {len(self._flat)}")
            steps = set()
            for (p, _), i in zip(self._flat, ids):
                st = sd["state"].get(i, sd["state"].get(str(i)))
                if st is None:  # torch keeps no state for a parameter that was never stepped
                    self.state[p]["exp_avg"].zero_()
                    self.state[p]["exp_avg_sq"].zero_()
                    continue
                if tuple(st["exp_avg"].shape) != tuple(p.shape):
                    raise ValueError(f"optimizer state {i}: shape {tuple(st['exp_avg'].shape)} vs parameter {tuple(p.shape)}")
                self.state[p]["exp_avg"].copy_(st["exp_avg"])
                self.state[p]["exp_avg_sq"].copy_(st["exp_avg_sq"])
                steps.add(int(float(st["step"])))
            if len(steps) > 1:
                raise ValueError(f"per-parameter step counts differ ({sorted(steps)}): one shared step count is kept")
            self.step_count.fill_(steps.pop() if steps else 0)
            # one learning rate + a scale per group: take the group with scale 1 (or the first) as the base
            base = next((float(g["lr"]) / float(g.get("lr_scale", 1.0)) for g in groups if float(g.get("lr_scale", 1.0)) != 0.0), None)
            if base is not None:
                self.lr.fill_(base)
        else:
            self.step_count.fill_(int(sd["step"]))
            self.lr.fill_(float(sd["lr"]))
            for (p, _), s_ in zip(self._flat, sd["state"]):
                self.state[p]["exp_avg"].copy_(s_["exp_avg"])
                self.state[p]["exp_avg_sq"].copy_(s_["exp_avg_sq"])
        self.sync_low_precision()
