"""`AudioSignal` container.  The reference takes `audiotools.AudioSignal` (third party, absent
here); only `.audio_data`, `.sample_rate`, `.device`, `.to()`, `.batch_size`, `+` and the
constructor are touched on this path (model/generator.py:383-416, model/watermarking.py:440),
so a stand-in with exactly that surface is used when audiotools is not importable."""
from __future__ import annotations

import torch

try:  # pragma: no cover - audiotools is not installed in the build image
    from audiotools import AudioSignal as _ATSignal  # type: ignore
except Exception:  # noqa: BLE001
    _ATSignal = None


class AudioSignal:
    def __init__(self, audio_data, sample_rate: int = 16000):
        if not torch.is_tensor(audio_data):
            audio_data = torch.as_tensor(audio_data, dtype=torch.float32)
        if audio_data.dim() == 1:
            audio_data = audio_data[None, None]
        elif audio_data.dim() == 2:
            audio_data = audio_data[None]
        self.audio_data = audio_data
        self.sample_rate = sample_rate

    @property
    def device(self):
        return self.audio_data.device

    @property
    def batch_size(self):
        return self.audio_data.shape[0]

    @property
    def signal_length(self):
        return self.audio_data.shape[-1]

    def to(self, device):
        self.audio_data = self.audio_data.to(device)
        return self

    def clone(self):
        return AudioSignal(self.audio_data.clone(), self.sample_rate)

    def __add__(self, other):
        o = other.audio_data if is_signal(other) else other
        return AudioSignal(self.audio_data + o, self.sample_rate)

    __radd__ = __add__


def is_signal(x) -> bool:
    return isinstance(x, AudioSignal) or (_ATSignal is not None and isinstance(x, _ATSignal)) or (
        hasattr(x, "audio_data") and hasattr(x, "sample_rate"))


def make_like(sig, audio_data):
    cls = type(sig) if is_signal(sig) else AudioSignal
    try:
        return cls(audio_data, sample_rate=sig.sample_rate)
    except TypeError:
        return AudioSignal(audio_data, sample_rate=sig.sample_rate)
