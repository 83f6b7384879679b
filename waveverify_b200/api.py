"""Public API: the B200 drop-in for `waveverify.WaveVerify` (waveverify/core.py) plus the batched
tensor entry points the reference lacks (embed_batch / detect_batch / locate_batch) and an exact
streaming embedder for long-form audio.

File I/O: the reference loads through torchaudio (no codec available in the build image); here
torchaudio is used when importable and plain PCM/float WAV is read and written with the standard
library otherwise, so the file-path API stays usable."""
from __future__ import annotations

import logging
import math
import wave
from pathlib import Path
from typing import List, Optional, Tuple, Union

import numpy as np
import torch

from .audio import AudioSignal
from .models import AudioWatermarking, Detector, Generator, Locator
from .watermark_id import WatermarkID

logger = logging.getLogger(__name__)
DEFAULT_SAMPLE_RATE = 16000
DEFAULT_BITS = 16


# ---- waveverify/utils.py glue -------------------------------------------------------------------
def message_to_tensor(message: Union[str, List[int]], bits: int = DEFAULT_BITS) -> torch.Tensor:
    """waveverify/utils.py:290-353 -> fp32 [1, bits]."""
    if bits <= 0:
        raise ValueError(f"Bits must be positive, got {bits}")
    if isinstance(message, str):
        if set(message) - {"0", "1"}:
            raise ValueError("Message string must contain only '0' and '1'")
        if len(message) != bits:
            raise ValueError(f"Message must be {bits} bits, got {len(message)}")
        vals = [int(c) for c in message]
    elif isinstance(message, list):
        if not all(isinstance(v, int) and v in (0, 1) for v in message):
            raise ValueError("Message list must contain only 0 and 1")
        if len(message) != bits:
            raise ValueError(f"Message must be {bits} elements, got {len(message)}")
        vals = message
    else:
        raise TypeError(f"Message must be str or list, got {type(message)}")
    return torch.tensor(vals, dtype=torch.float32).unsqueeze(0)


def tensor_to_message(tensor: torch.Tensor, threshold: float = 0.5) -> str:
    """waveverify/utils.py:356-412: time-mean (3-D), first batch item, `>= threshold`."""
    if not isinstance(tensor, torch.Tensor):
        raise TypeError(f"Expected torch.Tensor, got {type(tensor)}")
    if not 0 <= threshold <= 1:
        raise ValueError(f"Threshold must be between 0 and 1, got {threshold}")
    shape = tensor.shape
    if tensor.dim() == 3:
        tensor = tensor.mean(dim=2)
    if tensor.dim() == 2:
        tensor = tensor[0]
    if tensor.dim() != 1:
        raise ValueError(f"Cannot process tensor with shape {shape}")
    return "".join(str(int(b)) for b in (tensor >= threshold).int().tolist())


def _resample(w: torch.Tensor, sr: int, target: int) -> torch.Tensor:
    """The reference resamples with `torchaudio.transforms.Resample(sr, target)` (waveverify/utils.py:211-213: windowed-sinc
    polyphase filter, pure torch - needs no codec backend).  Linear interpolation only when torchaudio is not importable."""
    try:
        import torchaudio  # type: ignore
        return torchaudio.transforms.Resample(sr, target)(w)
    except ImportError:
        n = int(round(w.shape[-1] * target / sr))
        return torch.nn.functional.interpolate(w[None], size=n, mode="linear", align_corners=False)[0]


def load_audio(audio_path: Union[str, Path], target_sr: int = DEFAULT_SAMPLE_RATE) -> Tuple[torch.Tensor, int]:
    """waveverify/utils.py:170-224 -> (mono fp32 [1, T], sample_rate)."""
    p = Path(audio_path)
    if not p.exists():
        raise FileNotFoundError(f"Audio file not found: {p}")
    if not p.is_file():
        raise ValueError(f"Path is not a file: {p}")
    wav = None
    try:
        import torchaudio  # type: ignore
        wav, sr = torchaudio.load(str(p))
        if sr != target_sr:
            wav = torchaudio.transforms.Resample(sr, target_sr)(wav.mean(0, keepdim=True) if wav.shape[0] > 1 else wav)
            sr = target_sr
    except Exception:  # noqa: BLE001 - no codec backend: fall back to stdlib WAV
        wav = None
    if wav is None:
        try:
            with wave.open(str(p), "rb") as f:
                sr, nch, sw, n = f.getframerate(), f.getnchannels(), f.getsampwidth(), f.getnframes()
                raw = f.readframes(n)
        except Exception as e:  # noqa: BLE001
            raise RuntimeError(f"Cannot load audio file: {e}")
        if sw == 2:
            a = np.frombuffer(raw, "<i2").astype(np.float32) / 32768.0
        elif sw == 4:
            a = np.frombuffer(raw, "<i4").astype(np.float32) / 2147483648.0
        elif sw == 1:
            a = (np.frombuffer(raw, "u1").astype(np.float32) - 128.0) / 128.0
        else:
            raise RuntimeError(f"Cannot load audio file: unsupported sample width {sw}")
        wav = torch.from_numpy(a.reshape(-1, nch).T.copy())
        if sr != target_sr:
            wav = _resample(wav.mean(0, keepdim=True), sr, target_sr)
            sr = target_sr
    if wav.shape[0] > 1:
        wav = wav.mean(dim=0, keepdim=True)
    return wav.float(), sr


def save_audio(audio: torch.Tensor, path: Union[str, Path], sample_rate: int = DEFAULT_SAMPLE_RATE) -> None:
    """waveverify/utils.py:227-287: clamp to [-1, 1] and write (16-bit PCM WAV without torchaudio)."""
    if not isinstance(audio, torch.Tensor):
        raise ValueError(f"Audio must be torch.Tensor, got {type(audio)}")
    if sample_rate <= 0:
        raise ValueError(f"Sample rate must be positive, got {sample_rate}")
    p = Path(path)
    p.parent.mkdir(parents=True, exist_ok=True)
    shape = audio.shape
    if audio.dim() == 1:
        audio = audio.unsqueeze(0)
    elif audio.dim() == 3:
        audio = audio.squeeze(0)
    elif audio.dim() != 2:
        raise ValueError(f"Audio must be 1D, 2D, or 3D tensor, got shape {shape}")
    audio = torch.clamp(audio.detach().float().cpu(), -1.0, 1.0)
    try:
        import torchaudio  # type: ignore
        torchaudio.save(str(p), audio, sample_rate)
        return
    except Exception:  # noqa: BLE001
        pass
    pcm = (audio.numpy().T * 32767.0).round().astype("<i2")
    with wave.open(str(p), "wb") as f:
        f.setnchannels(audio.shape[0]); f.setsampwidth(2); f.setframerate(sample_rate)
        f.writeframes(pcm.tobytes())


# ---- exact streaming embed (BASELINE config 5) ----------------------------------------------------
GENERATOR_HALO = 5440      # 17 frames >= the generator's left receptive field


@torch.no_grad()
def embed_streaming(generator: Generator, audio: torch.Tensor, msg: torch.Tensor, chunk_samples: int = 320000,
                    halo: int = GENERATOR_HALO) -> torch.Tensor:
    """Embed a long clip chunk by chunk.  Every conv is causal with a finite left receptive field, so
    processing [s - halo, e) and discarding the first `halo` outputs reproduces the whole-clip result
    to rounding (chunk edges and halo are multiples of the 320-sample hop; the first chunk has no
    halo: true zero-padding start).  audio [1,1,T] (cuda) -> watermarked [1,1,T]."""
    hop = generator.hop_length
    if chunk_samples % hop or halo % hop:
        raise ValueError(f"chunk_samples and halo must be multiples of the hop ({hop})")
    if audio.dim() != 3 or audio.shape[0] != 1:
        raise ValueError("embed_streaming takes one clip [1, 1, T]")
    T = audio.shape[-1]
    out = torch.empty_like(audio, dtype=torch.float32)
    s = 0
    while s < T:
        e = min(T, s + chunk_samples)
        h = min(halo, s)
        _, y, _ = generator.embed_batch(audio[:, :, s - h:e].contiguous(), msg, want_wm=False)
        out[:, :, s:e] = y[:, :, h:]
        s = e
    return out


# ---- checkpoints (waveverify/core.py:141-168, 225-469) -------------------------------------------
_LOC_DEFAULT_KW = dict(dimension=64, channels_enc=32, n_residual_enc=1, strides=[8, 4])
_CFG_KEYS = ("sample_rate", "channels_audio", "dimension", "msg_dimension", "channels_enc", "channels_dec", "n_fft_base",
             "n_residual_enc", "n_residual_dec", "res_scale_enc", "res_scale_dec", "res_scale", "strides", "kernel_size",
             "last_kernel_size", "residual_kernel_size", "norm", "bias", "zero_init", "nbits", "output_dim", "embedding_dim",
             "embedding_layers", "freq_bands",
             # reference kwargs with one supported value: passed on so that an unsupported setting is rejected loudly
             "activation", "activation_kwargs", "norm_kwargs", "dilation_base", "skip", "act_all", "expansion", "groups",
             "encoder_l2norm", "spec", "spec_compression", "pad_mode", "causal", "inout_norm", "final_activation",
             "spec_layer", "spec_learnable")


def find_atomic_checkpoint(path: Path) -> Optional[Path]:
    """core.py:141-168 / 295-322: a .pth file, or the best.pth / latest.pth / first *.pth of a directory that holds
    a dict with a 'models' entry.  None = not an atomic checkpoint (legacy per-component layout)."""
    if path.is_file():
        return path
    if not path.is_dir():
        raise FileNotFoundError(f"checkpoint path not found: {path}")
    files = sorted(path.glob("*.pth"))
    ordered = [f for n in ("best.pth", "latest.pth") for f in files if f.name == n] + \
              [f for f in files if f.name not in ("best.pth", "latest.pth")]
    for f in ordered:
        try:
            ck = torch.load(str(f), map_location="cpu", weights_only=False)
        except Exception:  # noqa: BLE001 - the reference skips unreadable files too
            continue
        if isinstance(ck, dict) and "models" in ck:
            return f
    return None


def kwargs_from_checkpoint_config(config: Optional[dict], cls_name: str) -> dict:
    """The reference builds G / D / L from checkpoint['config'] (an argbind scope: 'Generator.channels_enc': 64, ...;
    core.py:230-276) before loading the weights."""
    if not config:
        return {}
    out = {}
    for k, v in config.items():
        if isinstance(k, str) and k.startswith(cls_name + "."):
            name = k[len(cls_name) + 1:]
            if name in _CFG_KEYS:
                out[name] = v
    if cls_name == "Locator":
        out.pop("nbits", None)            # conf/base.yml lists it, model/locator.py has no such kwarg
    return out


def infer_kwargs_from_state_dict(sd: dict) -> dict:
    """Without a stored config: `bias` and `zero_init` change the parameter inventory (modules/seanet.py:151-243: the
    *scale_param entries exist only with zero_init, conv biases only with bias), so they are read off the keys.
    Loading a zero_init=False checkpoint into a zero_init=True model would leave every scale parameter at 0."""
    keys = list(sd.keys())
    return dict(bias=any(k.startswith("encoder.conv_pre.") and k.endswith(".bias") for k in keys),
                zero_init=any(k.endswith("scale_param") for k in keys))


def _load_component(m, sd: dict, name: str, strict: bool):
    res = m.load_state_dict(sd, strict=strict)
    missing = [k for k in getattr(res, "missing_keys", []) if not k.endswith(".spec.weight")]   # fixed DFT buffers
    if missing:
        raise RuntimeError(f"checkpoint does not match the {name} architecture: {len(missing)} parameters missing "
                           f"(e.g. {missing[0]}); outputs would be silently wrong")
    unexpected = list(getattr(res, "unexpected_keys", []))
    if unexpected:
        logger.warning("%s: %d unexpected keys in the checkpoint (e.g. %s)", name, len(unexpected), unexpected[0])


class WaveVerify:
    """Drop-in for waveverify/core.py:WaveVerify.  `checkpoint` may be a path to an "atomic" .pth
    (dict with models/{generator,detector,locator} state dicts, parametrizations removed or not, optional 'config';
    waveverify/core.py:324-408), a directory holding best.pth / latest.pth / any atomic *.pth (core.py:141-168),
    a directory with generator/detector/locator sub-folders holding model.pth / weights.pth (legacy layout,
    core.py:428-469), or None for random-init models.  The networks are built from the checkpoint's 'config' when it
    has one (core.py:230-276), else `bias` / `zero_init` are inferred from the stored keys; explicit *_kwargs win.
    The reference's default "base" checkpoint has no published URL (waveverify/utils.py:45-52)."""

    def __init__(self, checkpoint: Optional[Union[str, Path]] = None, device: str = "auto",
                 generator_kwargs: Optional[dict] = None, detector_kwargs: Optional[dict] = None,
                 locator_kwargs: Optional[dict] = None):
        if device == "auto":
            device = "cuda" if torch.cuda.is_available() else "cpu"
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("waveverify_b200.WaveVerify needs a CUDA device (sm_100a); there is no CPU fallback")
        self.sample_rate = DEFAULT_SAMPLE_RATE
        self.watermark_bits = DEFAULT_BITS
        if str(checkpoint) == "base":
            raise RuntimeError("the reference publishes no 'base' checkpoint URL (waveverify/utils.py:45-52); "
                               "pass a checkpoint path or None for random-init models")
        user_kw = {"generator": generator_kwargs or {}, "detector": detector_kwargs or {}, "locator": locator_kwargs or {}}
        g, d, l = self.build_models(None if checkpoint is None else Path(checkpoint), user_kw)
        self.model = AudioWatermarking(g.to(self.device), d.to(self.device), l.to(self.device)).eval()

    @staticmethod
    def build_models(path: Optional[Path], user_kw: Optional[dict] = None):
        """Construct Generator / Detector / Locator for a checkpoint (config-driven) and load it.  Host only."""
        user_kw = user_kw or {}
        classes = (("generator", Generator, "Generator"), ("detector", Detector, "Detector"), ("locator", Locator, "Locator"))
        states, config, strict = {}, None, False
        if path is not None:
            atomic = find_atomic_checkpoint(path)
            if atomic is not None:
                ck = torch.load(str(atomic), map_location="cpu", weights_only=False)
                models = ck.get("models", ck) if isinstance(ck, dict) else ck
                config = ck.get("config") if isinstance(ck, dict) else None
                for name, _, _ in classes:
                    if name not in models:
                        raise RuntimeError(f"checkpoint {atomic} has no '{name}' state dict")
                    states[name] = models[name]
            else:                                     # legacy: <dir>/<component>/model.pth, strict (core.py:428-469)
                strict = True
                for name, _, _ in classes:
                    for fn in ("model.pth", "weights.pth"):
                        f = path / name / fn
                        if f.exists():
                            sd = torch.load(str(f), map_location="cpu", weights_only=False)
                            states[name] = sd.get("state_dict", sd) if isinstance(sd, dict) else sd
                            break
                    else:
                        raise FileNotFoundError(f"no {name}/model.pth under {path}")
        out = []
        for name, cls, cls_name in classes:
            kw = dict(_LOC_DEFAULT_KW) if name == "locator" else {}
            if name in states:
                kw.update(infer_kwargs_from_state_dict(states[name]))
            kw.update(kwargs_from_checkpoint_config(config, cls_name))
            kw.update(user_kw.get(name, {}))
            m = cls(**kw)
            if name in states:
                _load_component(m, states[name], name, strict)
            out.append(m)
        return tuple(out)

    # ---- batched tensor API (new) -------------------------------------------------------------
    @torch.no_grad()
    def embed_batch(self, audio: torch.Tensor, msg: torch.Tensor) -> torch.Tensor:
        """audio [B,1,T] cuda fp32, msg [B,16] -> watermarked audio [B,1,T]."""
        return self.model.generator.embed_batch(audio, msg, want_wm=False)[1]

    @torch.no_grad()
    def detect_batch(self, audio: torch.Tensor, presence: Optional[torch.Tensor] = None):
        """-> (bits u8 [B,16], confidence [B])."""
        d = self.model.detector.detect_batch(audio, presence=presence)
        return d["bits"], d["conf"]

    @torch.no_grad()
    def locate_batch(self, audio: torch.Tensor) -> torch.Tensor:
        """-> mask u8 [B,T] (locator logit > 0.5)."""
        return self.model.locator.locate_batch(audio)["mask"][:, 0]

    # ---- reference file API (waveverify/core.py:476-705) ----------------------------------------
    def _validate_watermark_id(self, watermark_id) -> WatermarkID:
        if isinstance(watermark_id, WatermarkID):
            return watermark_id
        try:
            return WatermarkID.custom(watermark_id)
        except (ValueError, TypeError) as e:
            raise ValueError(f"Invalid watermark_id: {e}. Use WatermarkID.for_creator(), .for_timestamp(), etc. "
                             f"or provide a 16-bit binary string, int (0-65535), or 2 bytes.")

    def embed(self, audio_path, watermark_id, output_path=None):
        try:
            wid = self._validate_watermark_id(watermark_id)
            audio, sr = load_audio(audio_path, self.sample_rate)
            msg = message_to_tensor(wid.to_bits(), self.watermark_bits).to(self.device)
            sig = AudioSignal(audio.unsqueeze(0).to(self.device), self.sample_rate)
            with torch.no_grad():
                _, wm_sig = self.model(sig, msg, phase="audio_sample")
            y = wm_sig.audio_data.squeeze(0)
            if output_path:
                save_audio(y, output_path, self.sample_rate)
            return y.cpu().numpy().squeeze(), self.sample_rate, wid
        except Exception as e:  # noqa: BLE001
            raise RuntimeError(f"Failed to embed watermark: {e}") from e

    def detect(self, audio_path):
        try:
            audio, sr = load_audio(audio_path, self.sample_rate)
            d = self.model.detector.detect_batch(audio.unsqueeze(0).to(self.device))
            bits = "".join(str(int(b)) for b in d["bits"][0].tolist())      # avg >= 0.5, utils.py:401
            return WatermarkID.custom(bits), float(d["avg"][0].mean().item())   # core.py:577-583
        except Exception as e:  # noqa: BLE001
            raise RuntimeError(f"Failed to detect watermark: {e}") from e

    def locate(self, audio_path) -> np.ndarray:
        try:
            audio, sr = load_audio(audio_path, self.sample_rate)
            r = self.model.locator.locate_batch(audio.unsqueeze(0).to(self.device), want_mask=False, want_probs=True)
            return r["probs"].squeeze().cpu().numpy()                        # sigmoid(logits), core.py:632
        except Exception as e:  # noqa: BLE001
            raise RuntimeError(f"Failed to locate watermark: {e}") from e

    def verify(self, audio_path, expected_watermark) -> bool:
        try:
            expected = self._validate_watermark_id(expected_watermark)
            detected, _ = self.detect(audio_path)
            return detected == expected
        except Exception as e:  # noqa: BLE001
            raise RuntimeError(f"Failed to verify watermark: {e}") from e
