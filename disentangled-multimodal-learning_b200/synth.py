"""Deterministic synthetic bags and weights (numpy PCG64, stable across torch versions).

The reference's data loader (data/dataset.py) is missing from its tree;
its bags are ResNet-50 patch features [N,1024] (post-ReLU, normalised) plus omic vectors
of 59 / 361 genes (config/config_mine_diag2021.yaml:27-30).  These helpers generate
stand-ins of that shape for the tests, the golden fixtures and bench.py.
"""
from __future__ import annotations

import math
import zlib
from typing import Dict, Iterable, Tuple

import numpy as np
import torch


def _rng(seed: int, name: str) -> np.random.Generator:
    return np.random.Generator(np.random.PCG64([seed & 0xFFFFFFFF, zlib.crc32(name.encode())]))


def normal(shape: Iterable[int], seed: int, name: str, scale: float = 1.0) -> torch.Tensor:
    a = _rng(seed, name).standard_normal(tuple(shape), dtype=np.float32) * np.float32(scale)
    return torch.from_numpy(a)


def uniform(shape: Iterable[int], seed: int, name: str, bound: float = 1.0) -> torch.Tensor:
    a = _rng(seed, name).random(tuple(shape), dtype=np.float32) * np.float32(2 * bound) - np.float32(bound)
    return torch.from_numpy(a)


def fill_like(shapes: Dict[str, Tuple[int, ...]], seed: int, gain: float = 1.0) -> Dict[str, torch.Tensor]:
    """Weights for a module given {state_dict key: shape}: U(-1/sqrt(fan_in), 1/sqrt(fan_in))
    for matrices/conv kernels (torch's default family), LayerNorm-style vectors near 1/0."""
    out = {}
    for k in sorted(shapes):
        shp = tuple(shapes[k])
        if k.endswith("output_range"):
            out[k] = torch.full(shp, 6.0)
        elif k.endswith("output_shift"):
            out[k] = torch.full(shp, -3.0)
        elif "norm" in k.split(".")[-2:][0] and k.endswith("weight") and len(shp) == 1:
            out[k] = 1.0 + 0.1 * uniform(shp, seed, k)
        elif len(shp) <= 1:
            # biases: small but non-zero so that every bias path is exercised
            out[k] = uniform(shp, seed, k, 0.1 * gain)
        elif k.endswith("cls_token"):
            out[k] = normal(shp, seed, k)
        else:
            fan_in = int(np.prod(shp[1:]))
            out[k] = uniform(shp, seed, k, gain / math.sqrt(max(fan_in, 1)))
    return out


def synthetic_bag(N: int, seed: int = 42, B: int = 1, feat: int = 1024):
    """x_path = per-feature standardised relu(randn) [B,N,feat]; omic vectors randn
    (SURVEY.md section 8(d) 'Synthetic inputs')."""
    x = normal((B, N, feat), seed, "x_path").clamp_(min=0)
    mu = x.mean(dim=1, keepdim=True)
    sd = x.std(dim=1, keepdim=True).clamp_(min=1e-6)
    x = (x - mu) / sd
    return dict(
        x_path=x,
        x_omic_tumor=normal((B, 59), seed, "x_omic_tumor"),
        x_omic_immune=normal((B, 361), seed, "x_omic_immune"),
        label_diag=torch.from_numpy(_rng(seed, "label_diag").integers(0, 4, size=(B,))).long(),
        label_grade=torch.from_numpy(_rng(seed, "label_grade").integers(0, 3, size=(B,))).long(),
        label_surv=torch.from_numpy(_rng(seed, "label_surv").integers(0, 4, size=(B,))).long(),
        censor=torch.from_numpy(_rng(seed, "censor").integers(0, 2, size=(B,))).long(),
    )
