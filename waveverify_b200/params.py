"""Parameter inventory of the three networks, keyed exactly like the reference's state_dict.

The reference builds its parameters as a side effect of constructing nn.Module trees
(modules/seanet.py:602-881, 1018-1210; model/detector.py:204-213; model/locator.py).  Here
the inventory is a flat table  name -> (shape, role)  derived from the topology, so the
host shim can (a) expose `state_dict()` / `load_state_dict()` with the reference's key
names and (b) hand folded weights to the CUDA library by name.
"""
from __future__ import annotations

import math
from collections import OrderedDict
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch

WAV_STD = 0.1122080159
SPEC_MEANS = [-4.554, -4.315, -4.021, -3.726, -3.477]
SPEC_STDS = [2.830, 2.837, 2.817, 2.796, 2.871]


@dataclass
class NetConfig:
    """Topology of one network (the supported subset of the reference constructor kwargs)."""
    kind: str                                  # "generator" | "detector" | "locator"
    sample_rate: int = 16000
    channels_audio: int = 1
    dimension: int = 128
    msg_dimension: int = 16
    channels_enc: int = 64
    channels_dec: int = 96
    n_fft_base: int = 64
    n_residual_enc: int = 2
    n_residual_dec: int = 3
    res_scale: float = 0.5773502691896258
    res_scale_dec: float = 0.5773502691896258
    strides: List[int] = field(default_factory=lambda: [8, 5, 4, 2])
    kernel_size: int = 5
    last_kernel_size: int = 5
    residual_kernel_size: int = 5
    norm: str = "weight_norm"
    bias: bool = True
    zero_init: bool = True
    nbits: int = 16
    output_dim: int = 32
    embedding_dim: int = 64
    embedding_layers: int = 2
    freq_bands: int = 4

    @property
    def hop_length(self) -> int:
        return int(np.prod(self.strides))

    @property
    def enc_ratios(self) -> List[int]:
        return list(reversed(self.strides))        # modules/seanet.py:646


_UNSUPPORTED = {
    # option -> the only value this B200 path implements (the shipped conf/base.yml value)
    "activation": "ELU", "dilation_base": 1, "skip": "identity", "act_all": False,
    "expansion": 1, "groups": -1, "encoder_l2norm": True, "spec": "stft",
    "spec_compression": "log", "pad_mode": "constant", "causal": True, "inout_norm": True,
    "final_activation": "Tanh",
}


def config_from_kwargs(kind: str, kw: dict) -> NetConfig:
    """Map reference constructor kwargs (model/generator.py:63-106, model/detector.py:82-114,
    model/locator.py:84-115) to a NetConfig; reject options the CUDA path does not implement."""
    kw = dict(kw)
    if kind == "locator" and "nbits" in kw:      # model/locator.py:84-115 has no such kwarg
        raise TypeError("Locator.__init__() got an unexpected keyword argument 'nbits'")
    for k, want in _UNSUPPORTED.items():
        if k in kw and kw[k] != want:
            raise NotImplementedError(
                f"{kind}: {k}={kw[k]!r} is not implemented by the B200 path (only {want!r})")
        kw.pop(k, None)
    ak = kw.pop("activation_kwargs", {"alpha": 1.0})
    if ak not in ({"alpha": 1.0}, {}, None):
        raise NotImplementedError(f"{kind}: activation_kwargs={ak!r} (only ELU alpha=1.0)")
    nk = kw.pop("norm_kwargs", {})
    if nk:
        raise NotImplementedError(f"{kind}: norm_kwargs={nk!r} not supported")
    for k in ("spec_layer", "spec_learnable"):      # accepted and ignored by the reference too
        kw.pop(k, None)
    if "res_scale_enc" in kw:
        kw["res_scale"] = kw.pop("res_scale_enc")
    if kw.get("norm", "weight_norm") not in ("weight_norm", "none", "weight_standardization"):
        raise NotImplementedError(f"{kind}: norm={kw['norm']!r} not supported")
    if kw.get("channels_audio", 1) != 1:
        raise NotImplementedError("only mono audio is supported")
    for k in ("kernel_size", "last_kernel_size", "residual_kernel_size"):
        if kw.get(k, 5) != 5:
            raise NotImplementedError(f"{kind}: {k}={kw[k]} (only 5)")
    cfg = NetConfig(kind=kind, **kw)
    if cfg.sample_rate <= 0:
        raise ValueError(f"Sample rate must be positive, got {cfg.sample_rate}")
    if cfg.dimension <= 0:
        raise ValueError(f"Dimension must be positive, got {cfg.dimension}")
    if not cfg.strides or any(s <= 0 for s in cfg.strides):
        raise ValueError(f"Invalid strides: {cfg.strides}. All values must be positive.")
    if kind == "locator":
        cfg.nbits = 1
    return cfg


# role strings drive the fixture initialiser and the fold
def _conv(spec, p, cout, cin_g, k, bias, norm, transposed=False):
    inner = "convtr.convtr" if transposed else "conv.conv"
    if norm == "weight_norm":
        # parametrised modules list `bias` first, then the parametrization's originals
        if bias:
            spec[f"{p}.{inner}.bias"] = ((cout,), "bias")
        spec[f"{p}.{inner}.parametrizations.weight.original0"] = ((cout, 1, 1), "g")
        spec[f"{p}.{inner}.parametrizations.weight.original1"] = ((cout, cin_g, k), "v")
        return
    if norm == "weight_standardization":
        spec[f"{p}.{inner}.weight_g"] = ((cout, 1, 1), "g")
        spec[f"{p}.{inner}.weight_v"] = ((cout, cin_g, k), "v")
        spec[f"{p}.{inner}.weight_scale"] = ((1,), "one")
    else:
        spec[f"{p}.{inner}.weight"] = ((cout, cin_g, k), "w")
    if bias:
        spec[f"{p}.{inner}.bias"] = ((cout,), "bias")


def _resblock(spec, p, C, cfg):
    if cfg.zero_init:                      # own parameters precede children in state_dict order
        spec[f"{p}.res_scale_param"] = ((1,), "scale_param")
    for a, d in ((1, 2), (4, 5)):
        _conv(spec, f"{p}.block.{a}", C, C, 1, False, cfg.norm)
        _conv(spec, f"{p}.block.{d}", C, 1, 5, cfg.bias, cfg.norm)


def _encoder_spec(spec: OrderedDict, cfg: NetConfig, p="encoder"):
    """Key order follows module registration order in SEANetEncoder.__init__ (seanet.py:657-846)."""
    C = cfg.channels_enc
    nfft = cfg.n_fft_base
    _conv(spec, f"{p}.conv_pre.1", C, 1, 5, cfg.bias, cfg.norm)
    blocks, specs, downs = OrderedDict(), OrderedDict(), OrderedDict()
    for s, r in enumerate(cfg.enc_ratios):
        for j in range(cfg.n_residual_enc):
            _resblock(blocks, f"{p}.blocks.{s}.{j}", C, cfg)
        if cfg.zero_init:
            specs[f"{p}.spec_blocks.{s}.scale_param"] = ((1,), "scale_param")
        specs[f"{p}.spec_blocks.{s}.spec.weight"] = ((nfft + 2, 1, nfft), "dft")
        _conv(specs, f"{p}.spec_blocks.{s}.layer", C, nfft // 2 + 1, 1, False, cfg.norm)
        _conv(downs, f"{p}.downsample.{s}.2", 2 * C, C, 1, False, cfg.norm)
        _conv(downs, f"{p}.downsample.{s}.3", 2 * C, 1, 2 * r, cfg.bias, cfg.norm)
        C *= 2
        nfft *= 2
    spec.update(blocks); spec.update(specs); spec.update(downs)
    if cfg.zero_init:
        spec[f"{p}.spec_post.scale_param"] = ((1,), "scale_param")
    spec[f"{p}.spec_post.spec.weight"] = ((nfft + 2, 1, nfft), "dft")
    _conv(spec, f"{p}.spec_post.layer", C, nfft // 2 + 1, 1, False, cfg.norm)
    _conv(spec, f"{p}.conv_post.1", C, 1, 5, False, cfg.norm)
    _conv(spec, f"{p}.conv_post.2", cfg.dimension, C, 1, cfg.bias, cfg.norm)
    E = cfg.embedding_dim
    spec[f"{p}.msg_embedding.0.weight"] = ((E, cfg.msg_dimension), "linear")
    spec[f"{p}.msg_embedding.0.bias"] = ((E,), "bias")
    for i in range(cfg.embedding_layers):
        j = 1 + 2 * i
        spec[f"{p}.msg_embedding.{j}.weight"] = ((E, E), "linear")
        spec[f"{p}.msg_embedding.{j}.bias"] = ((E,), "bias")
    for s in range(len(cfg.strides)):
        for b in range(cfg.freq_bands):
            q = f"{p}.film_layers.{s}.{b}"
            spec[f"{q}.gamma_layer.weight"] = ((1, E), "film_w")
            spec[f"{q}.gamma_layer.bias"] = ((1,), "film_gamma_b")
            spec[f"{q}.beta_layer.weight"] = ((1, E), "film_w")
            spec[f"{q}.beta_layer.bias"] = ((1,), "bias")


def _decoder_spec(spec: OrderedDict, cfg: NetConfig, p="decoder.model"):
    """SEANetDecoder.__init__ nn.Sequential indices (seanet.py:1067-1204)."""
    mult = 2 ** len(cfg.strides)
    C = mult * cfg.channels_dec
    i = 0
    _conv(spec, f"{p}.{i}", C, cfg.dimension, 1, False, cfg.norm); i += 1
    _conv(spec, f"{p}.{i}", C, 1, 5, cfg.bias, cfg.norm); i += 1
    dcfg = NetConfig(**{**cfg.__dict__})
    for r in cfg.strides:
        i += 2
        _conv(spec, f"{p}.{i}", C, 1, 2 * r, False, cfg.norm, transposed=True); i += 1
        _conv(spec, f"{p}.{i}", C // 2, C, 1, cfg.bias, cfg.norm); i += 1
        for _ in range(cfg.n_residual_dec):
            _resblock(spec, f"{p}.{i}", C // 2, dcfg); i += 1
        C //= 2
    i += 2
    _conv(spec, f"{p}.{i}", 1, C, 5, cfg.bias, cfg.norm)


def param_spec(cfg: NetConfig) -> "OrderedDict[str, Tuple[Tuple[int, ...], str]]":
    spec: OrderedDict = OrderedDict()
    _encoder_spec(spec, cfg)
    if cfg.kind == "generator":
        _decoder_spec(spec, cfg)
    else:
        hop = cfg.hop_length
        spec["reverse_convolution.weight"] = ((cfg.dimension, cfg.output_dim, hop), "head_w")
        spec["reverse_convolution.bias"] = ((cfg.output_dim,), "bias")
        spec["last_layer.weight"] = ((cfg.nbits, cfg.output_dim, 1), "last_w")
        spec["last_layer.bias"] = ((cfg.nbits,), "bias")
    return spec


def dft_weight(n_fft: int) -> torch.Tensor:
    """The fixed conv-as-DFT buffer, restating the fp32 op sequence of CausalSTFT.__init__
    (modules/conv.py:995-1012) so the fp32 angle rounding matches the reference buffer."""
    window = torch.hann_window(n_fft)
    t = torch.arange(n_fft, dtype=torch.float32).view(1, 1, n_fft)
    f = torch.arange(n_fft // 2 + 1, dtype=torch.float32).view(-1, 1, 1)
    ang = -2 * math.pi / n_fft * f * t
    return torch.cat([torch.cos(ang), torch.sin(ang)], dim=0) * window


def fixture_state_dict(cfg: NetConfig, seed: int = 0) -> "OrderedDict[str, torch.Tensor]":
    """Deterministic NON-DEGENERATE weights: at the reference's default init FiLM
    is a no-op, the locator mask is all-zero and bits are clip independent, so parity fixtures
    use these instead.  numpy MT19937 seeded per key => identical on every machine."""
    import zlib
    sd: OrderedDict = OrderedDict()
    for name, (shape, role) in param_spec(cfg).items():
        rng = np.random.RandomState((zlib.crc32(name.encode()) + 7919 * seed) % (2 ** 31))
        if role == "dft":
            sd[name] = dft_weight(shape[-1]); continue
        z = rng.standard_normal(shape).astype(np.float32)
        if role == "g":
            a = np.clip(1.0 + 0.25 * z, 0.3, None)
            if name.startswith("decoder.model.") and shape[0] == 1:
                a = 0.6 * a                       # last conv -> watermark amplitude
            if ".spec_" in name:
                a = 0.7 * a
            v = a
        elif role in ("v", "w"):
            v = z / math.sqrt(max(1, shape[1] * shape[2]))
        elif role == "one":
            v = np.ones(shape, np.float32)
        elif role == "bias":
            v = 0.1 * z
        elif role == "scale_param":
            v = 1.0 + 0.3 * z
        elif role == "linear":
            v = 1.5 * z / math.sqrt(shape[1])
        elif role == "film_w":
            v = 0.08 * z
        elif role == "film_gamma_b":
            v = 1.0 + 0.1 * z
        elif role == "head_w":
            v = z / math.sqrt(shape[0])
        elif role == "last_w":
            v = (3.0 if cfg.kind == "detector" else 1.2) * z / math.sqrt(shape[1])
        else:
            raise KeyError(role)
        if name == "last_layer.bias" and cfg.kind == "locator":
            v = v + 0.4
        sd[name] = torch.from_numpy(np.ascontiguousarray(v, dtype=np.float32))
    return sd


def default_init_state_dict(cfg: NetConfig, seed: int = 0) -> "OrderedDict[str, torch.Tensor]":
    """Random init in the spirit of the reference (trunc-normal 0.02 convs / linears, zero biases,
    N(0,1) bias before the L2 norm, zero scale params; seanet.py:825-857).  Not bit-identical to
    the reference's RNG stream - real use loads a checkpoint."""
    g = torch.Generator().manual_seed(seed)
    sd: OrderedDict = OrderedDict()
    for name, (shape, role) in param_spec(cfg).items():
        if role == "dft":
            sd[name] = dft_weight(shape[-1])
        elif role in ("v", "w", "linear", "film_w", "last_w", "head_w"):
            t = torch.empty(shape)
            torch.nn.init.trunc_normal_(t, std=0.02, generator=g)
            sd[name] = t
        elif role == "g":
            sd[name] = torch.zeros(shape)          # filled below from ||v||
        elif role == "one":
            sd[name] = torch.ones(shape)
        elif role == "scale_param":
            sd[name] = torch.zeros(shape)
        else:
            sd[name] = torch.zeros(shape)
    for name in list(sd):
        if name.endswith("original0"):
            v = sd[name[:-1] + "1"]
            sd[name] = v.flatten(1).norm(dim=1).view(-1, 1, 1)
        elif name.endswith("weight_g"):
            sd[name] = torch.ones_like(sd[name])
    k = "encoder.conv_post.2.conv.conv.bias"
    if k in sd:
        sd[k] = torch.randn(sd[k].shape, generator=g)
    return sd
