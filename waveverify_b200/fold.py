"""Fold re-parametrised conv weights of a reference-style state_dict into plain tensors.

weight-norm (torch parametrizations.weight_norm applied at modules/conv.py:74, dim=0):
    W[o] = g[o] * v[o] / ||v[o]||          for Conv1d AND ConvTranspose1d (per leading index)
weight-standardisation (modules/weight_standardization.py:108-147):
    W = g * scale * (v - mean) / sqrt(max(var * fan_in, 1e-7))
Checkpoints written by scripts/train.py:1624-1629 already hold plain `weight` keys.
Do NOT use v.norm(dim=0) as scripts/train.py:1524-1587 does - wrong axis.
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Dict

import torch

_WN_G = ".parametrizations.weight.original0"
_WN_V = ".parametrizations.weight.original1"


def fold_state_dict(sd: Dict[str, torch.Tensor]) -> "OrderedDict[str, torch.Tensor]":
    out: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    for key, val in sd.items():
        if key.endswith(_WN_V):
            base = key[: -len(_WN_V)]
            g = sd[base + _WN_G].detach().double().cpu()
            v = val.detach().double().cpu()
            norm = torch.linalg.vector_norm(v.reshape(v.shape[0], -1), dim=1)
            w = v * (g.reshape(-1) / norm).reshape(-1, *([1] * (v.dim() - 1)))
            out[base + ".weight"] = w.float().contiguous()
        elif key.endswith(_WN_G):
            continue
        elif key.endswith(".weight_v"):
            base = key[: -len(".weight_v")]
            v = val.detach().double().cpu()
            g = sd[base + ".weight_g"].detach().double().cpu()
            sc = sd[base + ".weight_scale"].detach().double().cpu() if base + ".weight_scale" in sd else 1.0
            flat = v.reshape(v.shape[0], -1)
            shape = (-1, *([1] * (v.dim() - 1)))
            mean = flat.mean(dim=1).reshape(shape)
            var = flat.var(dim=1, unbiased=False).reshape(shape)
            w = g * sc * (v - mean) / torch.sqrt(torch.clamp(var * flat.shape[1], min=1e-7))
            out[base + ".weight"] = w.float().contiguous()
        elif key.endswith(".weight_g") or key.endswith(".weight_scale"):
            continue
        else:
            out[key] = val.detach().float().cpu().contiguous()
    return out


def unfold_plain_weight(w: torch.Tensor):
    """Plain weight -> (g, v) with g*v/||v|| == w, for loading parametrization-free checkpoints."""
    flat = w.reshape(w.shape[0], -1)
    g = torch.linalg.vector_norm(flat, dim=1).reshape(-1, *([1] * (w.dim() - 1)))
    v = w.clone()
    zero = g.reshape(-1) == 0                 # all-zero rows (zero-initialised / pruned channels): g = 0 with a unit-norm v,
    if bool(zero.any()):                      # so that g * v / ||v|| stays 0 instead of 0 * 0 / 0 = NaN
        v.reshape(v.shape[0], -1)[zero] = flat.shape[1] ** -0.5
    return g, v
