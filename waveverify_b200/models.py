"""Host-side mirror of the reference model classes for the embed / detect / locate path.

`Generator`, `Detector`, `Locator` and `AudioWatermarking` keep the reference's constructor
kwargs, forward signatures, attributes, error behaviour and state_dict key names
(model/generator.py, model/detector.py, model/locator.py, model/watermarking.py:423-441) but
own no PyTorch compute: every forward folds the parameters once, hands them to the sm_100a
library (include/wv_b200.h) and launches hand-written CUDA kernels on the caller's stream.
There is no CPU / eager fallback - inputs must live on a CUDA device.
"""
from __future__ import annotations

import ctypes as C
import logging
import math
from typing import Dict, Optional, Tuple

import numpy as np
import torch
from torch import nn

from . import _lib
from .audio import AudioSignal, is_signal, make_like
from .fold import _WN_G, _WN_V, fold_state_dict, unfold_plain_weight
from .params import NetConfig, config_from_kwargs, default_init_state_dict, param_spec

logger = logging.getLogger(__name__)


class _Net:
    """Owns one `wv_net*` (weights + plans + workspace on one device)."""

    def __init__(self, cfg: NetConfig, folded: Dict[str, torch.Tensor], device_index: int, precise: bool = False):
        L = _lib.lib()
        c = _lib.NetConfigC()
        c.kind = _lib.KIND[cfg.kind]
        c.sample_rate = cfg.sample_rate
        c.dimension = cfg.dimension
        c.channels_enc = cfg.channels_enc
        c.channels_dec = cfg.channels_dec
        c.n_fft_base = cfg.n_fft_base
        c.n_residual_enc = cfg.n_residual_enc
        c.n_residual_dec = cfg.n_residual_dec
        c.n_strides = len(cfg.strides)
        for i, s in enumerate(cfg.strides):
            c.strides[i] = int(s)
        c.res_scale_enc = cfg.res_scale
        c.res_scale_dec = cfg.res_scale_dec
        c.nbits = cfg.nbits
        c.output_dim = cfg.output_dim
        c.msg_dimension = cfg.msg_dimension
        c.embedding_dim = cfg.embedding_dim
        c.embedding_layers = cfg.embedding_layers
        c.freq_bands = cfg.freq_bands
        c.precise = int(bool(precise))
        arr = (_lib.TensorC * len(folded))()
        keep = []
        for i, (name, t) in enumerate(folded.items()):
            t = t.detach().to(torch.float32).cpu().contiguous()
            keep.append(t)
            arr[i].name = name.encode()
            arr[i].data = t.data_ptr()
            arr[i].ndim = t.dim()
            for j, s in enumerate(t.shape):
                arr[i].shape[j] = s
        h = C.c_void_p()
        _lib.check(L.wv_net_create(C.byref(c), arr, len(folded), device_index, C.byref(h)),
                   f"wv_net_create({cfg.kind})")
        self.handle = h
        self.device_index = device_index
        self._L = L

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                self._L.wv_net_destroy(self.handle)
                self.handle = None
        except Exception:  # noqa: BLE001 - interpreter teardown
            pass


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream(device) -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


class _B200Module(nn.Module):
    KIND = ""

    def __init__(self, **kwargs):
        super().__init__()
        self.cfg = config_from_kwargs(self.KIND, kwargs)
        self._spec = param_spec(self.cfg)
        init = default_init_state_dict(self.cfg)
        for name, (shape, role) in self._spec.items():
            self._register(name, init[name], buffer=(role == "dft"))
        self._nets: Dict[bool, _Net] = {}      # precise flag -> native net (fp16 fast path / fp32-accurate path)
        self._net_sig = None
        self._chunk_samples = 0
        self.eval()

    # ---- parameter tree with the reference's key names ----------------------------------
    def _register(self, name: str, tensor: torch.Tensor, buffer: bool):
        parts = name.split(".")
        m: nn.Module = self
        for p in parts[:-1]:
            if p not in m._modules:
                m.add_module(p, nn.Module())
            m = m._modules[p]
        if buffer:
            m.register_buffer(parts[-1], tensor.clone())
        else:
            m.register_parameter(parts[-1], nn.Parameter(tensor.clone(), requires_grad=False))

    def load_state_dict(self, state_dict, strict: bool = True, assign: bool = False):
        """Accepts the reference's parametrised keys (`...parametrizations.weight.original0/1`)
        and parametrization-free checkpoints with plain `...weight` (waveverify/core.py:370-408)."""
        own = set(self._spec.keys())
        sd = {}
        for k, v in state_dict.items():
            if k.endswith(".weight") and k not in own and (k[: -len(".weight")] + _WN_V) in own:
                g, vv = unfold_plain_weight(v.detach().float())
                sd[k[: -len(".weight")] + _WN_G] = g
                sd[k[: -len(".weight")] + _WN_V] = vv
            else:
                sd[k] = v
        self._net_sig = None
        return super().load_state_dict(sd, strict=strict)

    # ---- device / native handle -----------------------------------------------------------
    @property
    def device(self) -> torch.device:
        return next(self.parameters()).device

    def _signature(self):
        return tuple((p.data_ptr(), p._version) for p in list(self.parameters()) + list(self.buffers()))

    PRECISE_DEFAULT = False     # Locator: True (its thresholded mask must equal the fp32 reference's)

    def _native(self, precise: Optional[bool] = None) -> _Net:
        """The native net (built lazily, rebuilt when a parameter changes).  precise=True selects the
        fp32-accurate twin (split-fp16 tensor-core operands, fp32 epilogues; include/wv_b200.h)."""
        if precise is None:
            precise = self.PRECISE_DEFAULT
        precise = bool(precise)
        dev = self.device
        if dev.type != "cuda":
            raise RuntimeError(
                f"waveverify_b200 {self.KIND} must live on a CUDA device (got {dev}); there is no "
                "CPU fallback - call .cuda() first")
        sig = (dev.index, self._signature())
        if self._net_sig != sig:
            self._nets = {}
            self._net_sig = sig
        if precise not in self._nets:
            folded = fold_state_dict(self.state_dict())
            idx = dev.index if dev.index is not None else torch.cuda.current_device()
            net = _Net(self.cfg, folded, idx, precise=precise)
            if self._chunk_samples:
                _lib.check(_lib.lib().wv_net_set_chunk(net.handle, int(self._chunk_samples)), "set_chunk")
            self._nets[precise] = net
        return self._nets[precise]

    def set_chunk_samples(self, n: int):
        """Process at most ~n samples (clips x T) per internal sub-batch; 0 = whole batch."""
        self._chunk_samples = int(n)
        for net in self._nets.values():
            _lib.check(_lib.lib().wv_net_set_chunk(net.handle, int(n)), "set_chunk")

    def set_range_check(self, enable: bool):
        """Scan every launch's fp16 outputs for saturated (|v| = 65504) / non-finite values (slow; diagnostics)."""
        _lib.check(_lib.lib().wv_net_set_range_check(self._native(False).handle, int(bool(enable))), "set_range_check")

    def range_read(self) -> dict:
        """Counts since the last read: dict(saturated, nonfinite, max_abs).  Synchronises and resets."""
        sat, bad, mx = C.c_ulonglong(0), C.c_ulonglong(0), C.c_float(0.0)
        _lib.check(_lib.lib().wv_net_range_read(self._native(False).handle, C.byref(sat), C.byref(bad), C.byref(mx)), "range_read")
        return dict(saturated=int(sat.value), nonfinite=int(bad.value), max_abs=float(mx.value))

    def set_profile(self, enable: bool):
        _lib.check(_lib.lib().wv_net_set_profile(self._native().handle, int(bool(enable))), "set_profile")

    def profile_read(self):
        """Per-launch records of the last profiled forward: dict(tag, cls, ms, flops, bytes)."""
        L = _lib.lib()
        n = 4096
        ms = (C.c_float * n)(); fl = (C.c_double * n)(); by = (C.c_double * n)(); cl = (C.c_int * n)()
        k = L.wv_net_profile_read(self._native().handle, n, ms, fl, by, cl)
        if k < 0:
            _lib.check(k, "wv_net_profile_read")
        h = self._native().handle
        return [dict(tag=L.wv_net_profile_tag(h, i).decode(), cls=int(cl[i]), ms=float(ms[i]),
                     flops=float(fl[i]), bytes=float(by[i])) for i in range(k)]

    @torch.no_grad()
    def debug_tap(self, audio: torch.Tensor, msg, tag: str, which: int, shape, dtype=torch.float16):
        """Run up to the launch tagged `tag` and return its output buffer (tests only)."""
        x = self._check_audio(audio)
        B, _, T = x.shape
        m = None
        if msg is not None:
            m = msg.to(x.device).float().contiguous()
        out = torch.empty(shape, dtype=dtype, device=x.device)
        wr = C.c_size_t(0)
        _lib.check(_lib.lib().wv_debug_tap(self._native().handle, _ptr(x), _ptr(m), B, T, tag.encode(), which,
                                           _ptr(out), out.numel() * out.element_size(), C.byref(wr)), "wv_debug_tap")
        if wr.value != out.numel() * out.element_size():
            raise RuntimeError(f"tap '{tag}' holds {wr.value} bytes, expected {out.numel() * out.element_size()}")
        return out

    def launches(self, B: int, T: int) -> int:
        return int(_lib.lib().wv_net_launches(self._native().handle, B, T))

    def workspace_bytes(self) -> int:
        return int(_lib.lib().wv_net_workspace_bytes(self._native().handle))

    def _check_audio(self, audio: torch.Tensor) -> torch.Tensor:
        if audio.dim() == 2:
            audio = audio.unsqueeze(1)
        if audio.dim() != 3 or audio.shape[1] != 1:
            raise ValueError(f"Expected audio of shape [B, 1, T], got {tuple(audio.shape)}")
        if audio.device != self.device:
            raise RuntimeError(
                f"audio is on {audio.device} but the model is on {self.device}; there is no CPU fallback")
        if audio.shape[0] < 1 or audio.shape[-1] < 1:
            raise ValueError(f"Empty audio tensor {tuple(audio.shape)}")
        return audio.detach().to(torch.float32).contiguous()


class Generator(_B200Module):
    """Drop-in for model/generator.py:Generator (encoder + decoder, returns the RESIDUAL)."""
    KIND = "generator"

    def __init__(self, **kwargs):
        super().__init__(**kwargs)
        self.nbits = self.cfg.nbits
        self.ratios = self.cfg.strides
        self.dimension = self.cfg.dimension
        self.sample_rate = self.cfg.sample_rate
        self.hop_length = self.cfg.hop_length

    def preprocess(self, audio_data: torch.Tensor, sample_rate: Optional[int] = None) -> torch.Tensor:
        """model/generator.py:245-288 (unused by forward, kept for API parity)."""
        if sample_rate is None:
            sample_rate = self.sample_rate
        assert sample_rate == self.sample_rate, \
            f"Sample rate mismatch: expected {self.sample_rate}, got {sample_rate}"
        n = audio_data.shape[-1]
        pad = math.ceil(n / self.hop_length) * self.hop_length - n
        return nn.functional.pad(audio_data, (0, pad)) if pad > 0 else audio_data

    def _msg(self, msg: torch.Tensor, B: int) -> torch.Tensor:
        msg = msg.to(self.device).float()                      # modules/seanet.py:909
        if msg.dim() != 2 or msg.shape[1] != self.cfg.msg_dimension:
            raise ValueError(f"msg must be [B, {self.cfg.msg_dimension}], got {tuple(msg.shape)}")
        if msg.shape[0] > B:                                   # modules/seanet.py:953-961
            msg = msg[:B]
        elif msg.shape[0] < B:
            msg = msg.repeat(int(np.ceil(B / msg.shape[0])), 1)[:B]
        return msg.contiguous()

    @torch.no_grad()
    def embed_batch(self, audio: torch.Tensor, msg: torch.Tensor, want_wm: bool = True,
                    want_y: bool = True, want_latent: bool = False):
        """audio [B,1,T] (or [B,T]) cuda fp32, msg [B,16] -> (wm [B,1,T], y = audio + wm, latent)."""
        x = self._check_audio(audio)
        B, _, T = x.shape
        m = self._msg(msg, B)
        net = self._native()
        wm = torch.empty_like(x) if want_wm else None
        y = torch.empty_like(x) if want_y else None
        F = math.ceil(T / self.hop_length)
        lat = torch.empty(B, self.dimension, F, device=x.device, dtype=torch.float32) if want_latent else None
        _lib.check(_lib.lib().wv_generator_forward(net.handle, _ptr(x), _ptr(m), B, T, _ptr(wm), _ptr(y),
                                                   _ptr(lat), _stream(x.device)), "wv_generator_forward")
        return wm, y, lat

    def encode(self, audio_data: torch.Tensor, msg: torch.Tensor) -> torch.Tensor:
        """model/generator.py:290-332 -> latent [B, dimension, ceil(T/hop)]."""
        try:
            return self.embed_batch(audio_data, msg, want_wm=False, want_y=False, want_latent=True)[2]
        except Exception as e:  # noqa: BLE001
            raise RuntimeError(f"Encoding failed: {e}") from e

    @torch.no_grad()
    def decode(self, latent_codes: torch.Tensor) -> torch.Tensor:
        """model/generator.py:334-358 -> audio [B, 1, F*hop]."""
        try:
            z = latent_codes
            if z.dim() != 3 or z.shape[1] != self.dimension:
                raise ValueError(f"latent must be [B, {self.dimension}, F], got {tuple(z.shape)}")
            if z.device != self.device:
                raise RuntimeError(f"latent is on {z.device} but the model is on {self.device}")
            z = z.detach().float().contiguous()
            B, _, F = z.shape
            out = torch.empty(B, 1, F * self.hop_length, device=z.device, dtype=torch.float32)
            _lib.check(_lib.lib().wv_generator_decode(self._native().handle, _ptr(z), B, F, _ptr(out),
                                                      _stream(z.device)), "wv_generator_decode")
            return out
        except Exception as e:  # noqa: BLE001
            raise RuntimeError(f"Decoding failed: {e}") from e

    def forward(self, audio_signal, msg: torch.Tensor, sample_rate: Optional[int] = None):
        """model/generator.py:360-423: returns an AudioSignal holding the watermark residual."""
        try:
            if not is_signal(audio_signal):
                raise ValueError("Input must be an AudioSignal object")
            wm, _, _ = self.embed_batch(audio_signal.audio_data, msg, want_y=False)
            return make_like(audio_signal, wm)
        except Exception as e:  # noqa: BLE001 - the reference re-raises everything as RuntimeError
            logger.error("Error in forward pass: %s", e)
            raise RuntimeError(f"Forward pass failed: {e}") from e


class _HeadModel(_B200Module):
    def __init__(self, **kwargs):
        super().__init__(**kwargs)
        self.ratios = self.cfg.strides
        self.dimension = self.cfg.dimension
        self.output_dim = self.cfg.output_dim
        self.sample_rate = self.cfg.sample_rate
        self.hop_length = self.cfg.hop_length
        self.stride = self.kernel_size = self.cfg.hop_length

    def _pad(self, audio_data, sample_rate):
        if sample_rate is None:
            sample_rate = self.sample_rate
        n = audio_data.shape[-1]
        pad = math.ceil(n / self.hop_length) * self.hop_length - n
        return n, (nn.functional.pad(audio_data, (0, pad)) if pad > 0 else audio_data)


class Detector(_HeadModel):
    """Drop-in for model/detector.py:Detector (dense SEANet encoder + ConvT head; the reference
    ships no mixture-of-experts)."""
    KIND = "detector"

    # Bit decisions (mean sigmoid >= 0.5) of the fp16 fast path are re-evaluated with the fp32-accurate net
    # wherever a clip has a bit whose mean lies within this distance of 0.5: 5x the measured fast-path
    # error on the mean (4e-5 for clips / masks of >= 4000 samples; short clips do not average the per-sample
    # logit error (<= 0.008) down: |d sigmoid| <= 0.25 * 0.008 = 2e-3).
    EXACT_TAU = 2e-4
    EXACT_TAU_SHORT = 5e-3
    EXACT_SHORT_SAMPLES = 4000

    def __init__(self, **kwargs):
        if kwargs.get("nbits", 16) <= 0:
            raise ValueError(f"Invalid nbits: {kwargs['nbits']}. Must be positive.")
        super().__init__(**kwargs)
        self.nbits = self.cfg.nbits
        self.exact_bits = True      # False: fp16 fast path only (bits exact outside a 3e-4 guard band)
        self.refine_slots = 0       # clips per pass of the precise net (0 = max(2, B/16))
        self._refine_counters = None   # device int32[2]: {clips re-evaluated, precise passes} (diagnostic)

    @property
    def recheck_count(self) -> int:
        """Clips re-evaluated by the precise net so far (reads a device counter: synchronises)."""
        return 0 if self._refine_counters is None else int(self._refine_counters[0].item())

    @property
    def recheck_passes(self) -> int:
        return 0 if self._refine_counters is None else int(self._refine_counters[1].item())

    def preprocess(self, audio_data: torch.Tensor, sample_rate: Optional[int] = None):
        """model/detector.py:222-276."""
        if sample_rate is None:
            sample_rate = self.sample_rate
        assert sample_rate == self.sample_rate, \
            f"Sample rate mismatch: expected {self.sample_rate}, got {sample_rate}"
        if audio_data.dim() != 3:
            raise ValueError(f"Expected 3D tensor, got {audio_data.dim()}D")
        return self._pad(audio_data, sample_rate)

    def _run(self, x, pm, want_logits, precise):
        B, _, T = x.shape
        dev = x.device
        nb = self.nbits
        logits = torch.empty(B, nb, T, device=dev, dtype=torch.float32) if want_logits else None
        bits = torch.empty(B, nb, device=dev, dtype=torch.uint8)
        avg = torch.empty(B, nb, device=dev, dtype=torch.float32)
        conf = torch.empty(B, device=dev, dtype=torch.float32)
        valid = torch.empty(B, nb, device=dev, dtype=torch.uint8)
        _lib.check(_lib.lib().wv_detector_forward(self._native(precise).handle, _ptr(x), B, T, _ptr(logits), _ptr(bits),
                                                  _ptr(avg), _ptr(conf), _ptr(valid), _ptr(pm), _stream(dev)),
                   "wv_detector_forward")
        return dict(bits=bits, avg=avg, conf=conf, valid=valid, logits=logits)

    @torch.no_grad()
    def detect_batch(self, audio: torch.Tensor, presence: Optional[torch.Tensor] = None,
                     want_logits: bool = False, precise: Optional[bool] = None):
        """audio [B,1,T] -> dict(bits u8 [B,nbits], avg, conf [B], valid, logits?).  Bit decode =
        sigmoid -> (masked) time mean -> >= 0.5 (waveverify/core.py:577-586, evaluate.py:471-494).

        The batch runs on the fp16 fast path; with `exact_bits` (default) every clip that has a bit whose mean
        lies within EXACT_TAU of the 0.5 threshold is re-evaluated by the fp32-accurate net and its
        bits / avg / conf / valid (and logits) are replaced, so the decoded bits equal the fp32 reference's.
        The re-evaluation is enqueued on the stream like everything else (no host synchronisation).
        precise=True runs the whole batch on the fp32-accurate net."""
        x = self._check_audio(audio)
        B, _, T = x.shape
        dev = x.device
        pm = None
        if presence is not None:
            pm = presence.to(dev).reshape(B, -1)
            if pm.shape[1] != T:
                raise ValueError(f"presence mask must have {T} samples per clip, got {pm.shape[1]}")
            pm = (pm != 0).to(torch.uint8).contiguous()
        if precise:
            return self._run(x, pm, want_logits, True)
        out = self._run(x, pm, want_logits, False)
        if not self.exact_bits:
            return out
        # device-side re-evaluation (no host round trip): selection, the precise net over the selected clips
        # `slots` at a time and the write-back are one CUDA graph with a WHILE node (wv_detector_refine)
        slots = self.refine_slots if self.refine_slots > 0 else max(2, -(-B // 16))
        if self._refine_counters is None or self._refine_counters.device != dev:
            self._refine_counters = torch.zeros(2, dtype=torch.int32, device=dev)
        _lib.check(_lib.lib().wv_detector_refine(
            self._native(True).handle, _ptr(x), B, T, _ptr(out["logits"]), _ptr(out["bits"]), _ptr(out["avg"]),
            _ptr(out["conf"]), _ptr(out["valid"]), _ptr(pm), float(self.EXACT_TAU), float(self.EXACT_TAU_SHORT),
            int(self.EXACT_SHORT_SAMPLES), int(min(slots, B)), _ptr(self._refine_counters), _stream(dev)),
            "wv_detector_refine")
        return out

    def decode(self, audio_data: torch.Tensor, orig_nframes: int) -> torch.Tensor:
        """model/detector.py:278-318 -> raw logits [B, nbits, orig_nframes]."""
        x = self._check_audio(audio_data)
        if orig_nframes != x.shape[-1]:
            # the reference encodes the (padded) input and trims the head output
            out = self.detect_batch(x, want_logits=True)["logits"]
            return out[:, :, :orig_nframes]
        return self.detect_batch(x, want_logits=True)["logits"]

    def postprocess(self, result: torch.Tensor, message_threshold: float = 0.5) -> torch.Tensor:
        """model/detector.py:320-364 (degenerate: always all-ones; kept for parity)."""
        if not 0 <= message_threshold <= 1:
            raise ValueError(f"message_threshold must be in [0, 1], got {message_threshold}")
        r = torch.softmax(result, dim=1).mean(dim=-1)
        return torch.gt(torch.sigmoid(r), message_threshold).int()

    def forward(self, audio_signal) -> torch.Tensor:
        """model/detector.py:366-391 -> raw logits [B, nbits, T]."""
        return self.detect_batch(audio_signal.audio_data, want_logits=True)["logits"]

    def detect(self, audio_signal, verbose: bool = False) -> torch.Tensor:
        """model/detector.py:393-434."""
        with torch.no_grad():
            bits = self.postprocess(self(audio_signal))
            if verbose:
                for i in range(bits.shape[0]):
                    print(f"Detection complete for batch {i}: {''.join(map(str, bits[i].cpu().numpy()))}")
        return bits


class Locator(_HeadModel):
    """Drop-in for model/locator.py:Locator.  Runs on the fp32-accurate ("precise") net by default: its mask is a
    threshold on raw logits (model/watermarking.py:717), which the fp16 fast path reproduces only outside a
    +-0.004 band; `exact = False` selects the fast path."""
    KIND = "locator"

    def __init__(self, **kwargs):
        super().__init__(**kwargs)
        self.exact = True

    @property
    def PRECISE_DEFAULT(self):   # noqa: N802 - overrides the class attribute of _B200Module
        return bool(getattr(self, "exact", True))

    def preprocess(self, audio_data: torch.Tensor, sample_rate: Optional[int] = None):
        """model/locator.py:186-226."""
        if sample_rate is None:
            sample_rate = self.sample_rate
        if sample_rate != self.sample_rate:
            raise ValueError(f"Sample rate mismatch: expected {self.sample_rate}, got {sample_rate}")
        return self._pad(audio_data, sample_rate)

    @torch.no_grad()
    def locate_batch(self, audio: torch.Tensor, want_logits: bool = False, want_mask: bool = True,
                     want_probs: bool = False):
        """audio [B,1,T] -> dict(mask u8 [B,1,T] = logit > 0.5 (model/watermarking.py:717),
        probs = sigmoid(logit) (waveverify/core.py:632), logits)."""
        x = self._check_audio(audio)
        B, _, T = x.shape
        dev = x.device
        logits = torch.empty(B, 1, T, device=dev, dtype=torch.float32) if want_logits else None
        mask = torch.empty(B, 1, T, device=dev, dtype=torch.uint8) if want_mask else None
        probs = torch.empty(B, 1, T, device=dev, dtype=torch.float32) if want_probs else None
        _lib.check(_lib.lib().wv_locator_forward(self._native().handle, _ptr(x), B, T, _ptr(logits), _ptr(mask),
                                                 _ptr(probs), _stream(dev)), "wv_locator_forward")
        return dict(mask=mask, probs=probs, logits=logits)

    def decode(self, audio_data: torch.Tensor, original_frame_count: int) -> torch.Tensor:
        """model/locator.py:228-265."""
        try:
            out = self.locate_batch(audio_data, want_logits=True, want_mask=False)["logits"]
            return out[:, :, :original_frame_count]
        except Exception as e:  # noqa: BLE001
            raise RuntimeError(f"Failed to decode audio: {e}")

    def forward(self, audio_signal) -> torch.Tensor:
        """model/locator.py:268-299 -> raw logits [B, 1, T]."""
        if audio_signal is None or audio_signal.audio_data is None:
            raise ValueError("Invalid audio signal input")
        return self.decode(audio_signal.audio_data, audio_signal.audio_data.shape[-1])


class AudioWatermarking(nn.Module):
    """The inference slice of model/watermarking.py:AudioWatermarking: phase 'audio_sample'
    (:423-441).  Training / validation phases (augmentations, effects) are out of scope."""

    def __init__(self, generator: Generator, detector: Detector, locator: Locator, **_unused):
        super().__init__()
        self.generator = generator
        self.detector = detector
        self.locator = locator

    def forward(self, signal, msg: torch.Tensor, phase: str = "audio_sample"):
        if phase != "audio_sample":
            raise NotImplementedError(
                f"phase={phase!r}: only 'audio_sample' (inference) is implemented by the B200 path")
        try:
            wm, y, _ = self.generator.embed_batch(signal.audio_data, msg)   # y = x + wm fused
        except Exception as e:  # noqa: BLE001
            raise RuntimeError(f"Forward pass failed: {e}") from e
        return make_like(signal, wm), make_like(signal, y)


def metric_counters(bits: torch.Tensor, valid: Optional[torch.Tensor], msg: torch.Tensor,
                    pred_mask: Optional[torch.Tensor], gt_mask: Optional[torch.Tensor],
                    counters: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Accumulate the six exact BER / MIoU counters (scripts/evaluate.py:498-505, 636-656) on the
    device: [bit_errors, valid_bits, I_fg, U_fg, I_bg, U_bg] (int64)."""
    dev = bits.device if bits is not None else pred_mask.device
    if counters is None:
        counters = torch.zeros(6, dtype=torch.int64, device=dev)

    def as_u8(t):
        """uint8 view of a flag tensor; the kernel treats any non-zero byte as set, so uint8 / bool inputs are passed as they are"""
        t = t.to(dev)
        if t.dtype == torch.bool:
            t = t.view(torch.uint8)
        elif t.dtype != torch.uint8:
            t = (t != 0).to(torch.uint8)
        return t.contiguous()

    b8 = v8 = m8 = p8 = g8 = None
    B = nb = 0
    if bits is not None:
        b8 = as_u8(bits)
        B, nb = b8.shape
        m8 = as_u8(msg)
        v8 = as_u8(valid) if valid is not None else None
    n_mask = 0
    if pred_mask is not None:
        p8 = as_u8(pred_mask)
        g8 = as_u8(gt_mask)
        if p8.numel() != g8.numel():
            raise ValueError(f"Shape mismatch: predicted={tuple(p8.shape)}, ground_truth={tuple(g8.shape)}")
        n_mask = p8.numel()
    _lib.check(_lib.lib().wv_metrics_accumulate(_ptr(b8), _ptr(v8), _ptr(m8), B, nb, _ptr(p8), _ptr(g8), n_mask,
                                                _ptr(counters), _stream(dev)), "wv_metrics_accumulate")
    return counters


def ber_miou(counters) -> Tuple[float, float]:
    c = [int(v) for v in (counters.tolist() if torch.is_tensor(counters) else counters)]
    ber = c[0] / c[1] if c[1] > 0 else 0.0
    fg = c[2] / c[3] if c[3] > 0 else 1.0
    bg = c[4] / c[5] if c[5] > 0 else 1.0
    return ber, 0.5 * (fg + bg)
