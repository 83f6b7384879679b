"""Import the upstream reference modules from /root/reference in THIS container only.

TEST INFRASTRUCTURE - not part of the product path. Used by oracle/make_golden.py and by
the in-container oracle-vs-reference check. The reference needs `audiotools` (not
installed), so a minimal stand-in for `AudioSignal` / `ml.BaseModel` is injected and
model/{generator,detector,locator}.py are loaded by path (importing the `model`
package would pull in matplotlib/julius/pesq/pystoi which are absent).
"""
import importlib.util
import os
import sys
import types

import torch
from torch import nn

REF_ROOT = os.environ.get("WV_REFERENCE_ROOT", "/root/reference")


class _AudioSignal:
    def __init__(self, audio_data, sample_rate=16000):
        self.audio_data = audio_data
        self.sample_rate = sample_rate

    @property
    def device(self):
        return self.audio_data.device

    @property
    def batch_size(self):
        return self.audio_data.shape[0]

    def to(self, device):
        self.audio_data = self.audio_data.to(device)
        return self

    def __add__(self, other):
        o = other.audio_data if isinstance(other, _AudioSignal) else other
        return _AudioSignal(self.audio_data + o, self.sample_rate)


def available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "modules"))


def load():
    """Returns (Generator, Detector, Locator, AudioSignal, cfg) from the reference tree."""
    import yaml
    if "audiotools" not in sys.modules:
        at = types.ModuleType("audiotools")
        at.AudioSignal = _AudioSignal
        ml = types.ModuleType("audiotools.ml")
        ml.BaseModel = nn.Module
        at.ml = ml
        sys.modules["audiotools"] = at
        sys.modules["audiotools.ml"] = ml
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    import logging
    logging.disable(logging.CRITICAL)
    out = {}
    for name in ("generator", "detector", "locator"):
        spec = importlib.util.spec_from_file_location(
            f"_wvref_{name}", os.path.join(REF_ROOT, "model", f"{name}.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        out[name] = mod
    with open(os.path.join(REF_ROOT, "conf", "base.yml")) as f:
        cfg = yaml.safe_load(f)
    return (out["generator"].Generator, out["detector"].Detector, out["locator"].Locator,
            sys.modules["audiotools"].AudioSignal, cfg)


def build(zero_init=False):
    """Build G, D, L from conf/base.yml with the SURVEY F5 fix-ups (bias=True, no Locator.nbits)."""
    G, D, L, AS, cfg = load()
    g = G(**{**cfg["Generator"], "bias": True, "zero_init": zero_init}).eval()
    d = D(**{**cfg["Detector"], "bias": True, "zero_init": zero_init}).eval()
    l = L(**{**{k: v for k, v in cfg["Locator"].items() if k != "nbits"}, "bias": True,
             "zero_init": zero_init}).eval()
    return g, d, l, AS
