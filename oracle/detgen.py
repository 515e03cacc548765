"""Deterministic tensor generator shared by the golden-fixture script and the tests.

TEST INFRASTRUCTURE ONLY (see oracle/mvit_oracle.py header).

Fixtures under tests/golden/ store only configuration, seeds and the *outputs* the
reference produced; inputs and weights are regenerated from the seeds with numpy's
PCG64 stream (bit-stable), so the fixtures stay small.
"""
from __future__ import annotations

import zlib
from typing import Dict, Tuple

import numpy as np
import torch


def _rng(seed: int, name: str) -> np.random.Generator:
    return np.random.Generator(np.random.PCG64([seed, zlib.crc32(name.encode())]))


def det_normal(shape, seed: int, name: str, std: float = 1.0, mean: float = 0.0) -> torch.Tensor:
    a = _rng(seed, name).standard_normal(tuple(shape)).astype(np.float32)
    return torch.from_numpy(a * np.float32(std) + np.float32(mean))


def det_params(shapes: Dict[str, Tuple[int, ...]], seed: int, perturb_1d: float = 0.1) -> Dict[str, torch.Tensor]:
    """Weights in the style of the reference init (trunc-normal std 0.02 for matrices /
    conv kernels / rel-pos tables / cls token; LN weight 1, biases 0 —
    video_model_builder.py:2018-2025, attention.py:300-310), with every 1-D parameter
    perturbed by ``perturb_1d``·N(0,1) so LN affine terms and biases are exercised
    (SURVEY.md §8d config 2)."""
    out = {}
    for name, shape in shapes.items():
        if len(shape) == 1:
            base = 1.0 if (name.endswith("weight") and ("norm" in name)) else 0.0
            out[name] = det_normal(shape, seed, name, perturb_1d, base)
        else:
            w = det_normal(shape, seed, name, 1.0).clamp_(-2.0, 2.0) * 0.02
            out[name] = w
    return out
