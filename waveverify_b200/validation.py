"""Validation path on the GPU.

Host-side mirror of the reference's validation step `AudioWatermarking._forward_valid`
(model/watermarking.py:443-525, 757-806): embed -> temporal augmentations -> effect -> Detector +
Locator -> BER / mIoU, with every tensor resident in HBM (the reference bounces each effect
through the CPU, watermarking.py:777-790).

* `LocalizationAugmentation` / `SequenceAugmentation` keep the reference's class names, constructor
  arguments, `forward` signatures and return tuples (utils/localization_augmentation.py:73-325,
  utils/seq_augmentation.py:42-277).  The random PLAN (which segments, which operation, which shift /
  permutation) is drawn on the host from `numpy.random` / `torch.randperm` in exactly the reference's
  call order, so a seeded run selects the same segments as the reference; the data movement is one
  CUDA gather pass (`wv_augment_gather`).  `augment()` applies both plans in a single pass.
* `apply_effect` covers the effects that need no third-party codec: identity, amplitude_scaling,
  quantization, random_noise / white_noise, sample_suppression, median_filter and the julius
  low / high / band-pass FIR filters (utils/effect_augmentation.py).  The rest (sox / ffmpeg / encodec,
  resampling, echo, smooth) raise NotImplementedError: they are out of scope.

No CPU fallback: tensors must live on a CUDA device.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Any, Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
from torch import nn

from . import _lib
from .audio import AudioSignal, make_like
from .models import _ptr, _stream, ber_miou, metric_counters

# utils/localization_augmentation.py:36-38
ORIGINAL_REVERT_PROB = 0.33
ZERO_REPLACE_PROB = 0.66
TARGET_AUGMENTATION_RATIO = 0.20
# utils/seq_augmentation.py:29-35
REVERSE_PROBABILITY = 0.3
CIRCULAR_SHIFT_PROBABILITY = 0.4
SHUFFLE_PROBABILITY = 0.3
DEFAULT_SEGMENT_DURATION = 0.5
DEFAULT_CHUNK_DIVISIONS = 4

OP_KEEP, OP_REVERT, OP_ZERO, OP_CROSS = 0, 1, 2, 3
SEQ_IDENTITY, SEQ_REVERSE, SEQ_SHIFT, SEQ_SHUFFLE, SEQ_CHUNK_SWAP = 0, 1, 2, 3, 4


@dataclass
class LocalizationPlan:
    seg_op: np.ndarray      # [B, S] uint8
    seg_src: np.ndarray     # [B, S] int32 (clip to copy from, for OP_CROSS)
    seg_len: int
    stats: Dict[str, float]


@dataclass
class SequencePlan:
    method: str             # the reference's method string ('reverse', 'circular_shift', 'shuffle', 'unchanged')
    kind: int = SEQ_IDENTITY
    a: int = 0
    b: int = 0
    c: int = 0
    perm: Optional[np.ndarray] = None   # [n] segment order (SEQ_SHUFFLE)

    def out_length(self, T: int) -> int:
        return len(self.perm) * self.c if self.kind == SEQ_SHUFFLE else T


def _need_cuda(t: torch.Tensor, what: str) -> torch.Tensor:
    if not torch.is_tensor(t):
        raise ValueError(f"{what} must be a tensor")
    if t.device.type != "cuda":
        raise RuntimeError(f"{what} is on {t.device}: waveverify_b200 has no CPU path")
    return t


def _gather(original: Optional[torch.Tensor], watermarked: torch.Tensor, gt_in: Optional[torch.Tensor],
            loc: Optional[LocalizationPlan], seq: Optional[SequencePlan], want_orig: bool = True,
            want_gt: bool = True):
    """One `wv_augment_gather` launch.  Tensors are [B, C, T] fp32 (rows = B*C)."""
    wm = _need_cuda(watermarked, "watermarked").float().contiguous()
    dev = wm.device
    B, Cn, T = wm.shape
    rows = B * Cn
    og = None
    if original is not None:
        og = _need_cuda(original, "original").float().contiguous()
        if og.shape != wm.shape:
            raise ValueError(f"Shape mismatch: original {tuple(og.shape)} != watermarked {tuple(wm.shape)}")
    gi = None
    if gt_in is not None:
        gi = _need_cuda(gt_in, "ground_truth_presence").float().contiguous()
        if gi.shape != wm.shape:
            raise ValueError("ground_truth_presence must have the shape of the audio")
    seg_op = seg_src = None
    seg_len = n_seg = 0
    if loc is not None:
        if Cn != 1:
            raise NotImplementedError("localization augmentation: mono audio only (conf/base.yml)")
        seg_op = torch.from_numpy(np.ascontiguousarray(loc.seg_op, dtype=np.uint8)).to(dev)
        seg_src = torch.from_numpy(np.ascontiguousarray(loc.seg_src, dtype=np.int32)).to(dev)
        seg_len, n_seg = int(loc.seg_len), int(loc.seg_op.shape[1])
    seq = seq or SequencePlan("unchanged")
    perm = None
    n_perm = 0
    if seq.kind == SEQ_SHUFFLE:
        perm = torch.from_numpy(np.ascontiguousarray(seq.perm, dtype=np.int32)).to(dev)
        n_perm = int(perm.numel())
    T_out = seq.out_length(T)
    out_wm = torch.empty(B, Cn, T_out, device=dev, dtype=torch.float32)
    out_og = torch.empty_like(out_wm) if (want_orig and og is not None) else None
    out_gt = torch.empty_like(out_wm) if want_gt else None
    _lib.check(_lib.lib().wv_augment_gather(
        _ptr(og), _ptr(wm), _ptr(gi), rows, T, _ptr(seg_op), _ptr(seg_src), seg_len, n_seg,
        int(seq.kind), int(seq.a), int(seq.b), int(seq.c), _ptr(perm), n_perm, T_out,
        _ptr(out_wm), _ptr(out_og), _ptr(out_gt), _stream(dev)), "wv_augment_gather")
    return out_wm, out_og, out_gt


class LocalizationAugmentation(nn.Module):
    """utils/localization_augmentation.py:73-325."""

    def __init__(self, sample_rate: int, window_duration: float) -> None:
        super().__init__()
        if sample_rate <= 0:
            raise ValueError(f"Sample rate must be positive, got {sample_rate}")
        if window_duration <= 0:
            raise ValueError(f"Window duration must be positive, got {window_duration}")
        self.sample_rate = sample_rate
        self.window_duration = window_duration
        self.segment_length = int(sample_rate * window_duration)        # :109
        self.stats: Dict[str, float] = {}

    def plan(self, batch_size: int, num_samples: int) -> LocalizationPlan:
        """The random draws of forward() (:258-309), in the reference's order: per clip one
        `np.random.choice(starts, k, replace=False)`, then per selected segment one `np.random.rand()`
        and, for a cross substitution, one `np.random.choice(other clips)`."""
        L = self.segment_length
        total_segments = int(np.ceil(num_samples / L))
        k = int(total_segments * TARGET_AUGMENTATION_RATIO)
        seg_op = np.zeros((batch_size, total_segments), np.uint8)
        seg_src = np.tile(np.arange(batch_size, dtype=np.int32)[:, None], (1, total_segments))
        counts = {"original_revert": 0, "zero_replace": 0, "cross_substitute": 0}
        for b in range(batch_size):
            starts = np.arange(0, num_samples, L)
            for start in np.random.choice(starts, k, replace=False):
                s = int(start) // L
                n = min(int(start) + L, num_samples) - int(start)
                p = np.random.rand()
                if p < ORIGINAL_REVERT_PROB:
                    seg_op[b, s] = OP_REVERT
                    counts["original_revert"] += n
                elif p < ZERO_REPLACE_PROB:
                    seg_op[b, s] = OP_ZERO
                    counts["zero_replace"] += n
                elif batch_size >= 2:
                    others = [j for j in range(batch_size) if j != b]
                    seg_op[b, s] = OP_CROSS
                    seg_src[b, s] = int(np.random.choice(others))
                    counts["cross_substitute"] += n
        total = batch_size * num_samples
        counts["unchanged"] = total - sum(counts.values())
        stats = {key: float(v / total * 100) for key, v in counts.items()} if total else dict.fromkeys(counts, 0.0)
        return LocalizationPlan(seg_op, seg_src, L, stats)

    @torch.no_grad()
    def forward(self, original: torch.Tensor, watermarked: torch.Tensor):
        if original.shape != watermarked.shape:
            raise ValueError(f"Shape mismatch: original {original.shape} != watermarked {watermarked.shape}")
        B, _, T = watermarked.shape
        plan = self.plan(B, T)
        wm, upd, gt = _gather(original, watermarked, None, plan, None)
        self.stats = plan.stats
        return AudioSignal(wm, self.sample_rate), gt, upd, self.stats


class SequenceAugmentation(nn.Module):
    """utils/seq_augmentation.py:42-277."""

    VALID = ["reverse", "circular_shift", "shuffle", "chunk_shuffle"]

    def __init__(self, sample_rate: int, methods: Optional[List[str]] = None) -> None:
        super().__init__()
        if sample_rate <= 0:
            raise ValueError(f"Sample rate must be positive, got {sample_rate}")
        self.sample_rate = sample_rate
        if methods is None:
            self.methods = list(self.VALID)
        else:
            invalid = set(methods) - set(self.VALID)
            if invalid:
                raise ValueError(f"Invalid augmentation methods: {invalid}. Valid methods: {self.VALID}")
            self.methods = methods
        self.stats = {m: 0 for m in self.methods}
        self.stats["unchanged"] = 0

    def plan(self, num_samples: int) -> SequencePlan:
        """The draws of forward() (:154-205): one `np.random.rand()` picks the method (the
        probabilities sum to 1, so 'chunk_shuffle' is never drawn, as in the reference), then
        `np.random.randint(1, T)` for the shift or `torch.randperm(n)` for the shuffle."""
        r = np.random.rand()
        if r < REVERSE_PROBABILITY:
            return SequencePlan("reverse", SEQ_REVERSE)
        if r < REVERSE_PROBABILITY + CIRCULAR_SHIFT_PROBABILITY:
            return SequencePlan("circular_shift", SEQ_SHIFT, a=int(np.random.randint(1, num_samples)))
        if r < REVERSE_PROBABILITY + CIRCULAR_SHIFT_PROBABILITY + SHUFFLE_PROBABILITY:
            seg = int(DEFAULT_SEGMENT_DURATION * self.sample_rate)
            if num_samples >= 2 * seg:
                n = num_samples // seg
                return SequencePlan("shuffle", SEQ_SHUFFLE, c=seg, perm=torch.randperm(n).numpy().astype(np.int32))
            return SequencePlan("unchanged_short")        # counted as 'shuffle' by the reference (:207)
        return SequencePlan("unchanged")

    @staticmethod
    def chunk_swap_plan(chunk1_start: int, chunk2_start: int, chunk_size: int) -> SequencePlan:
        """'chunk_shuffle' (:209-246) with explicit positions (the reference never draws it)."""
        return SequencePlan("chunk_shuffle", SEQ_CHUNK_SWAP, a=int(chunk1_start), b=int(chunk2_start), c=int(chunk_size))

    @torch.no_grad()
    def forward(self, updated_original: torch.Tensor, watermarked: torch.Tensor, ground_truth_presence: torch.Tensor):
        try:   # the reference wraps its own shape check too: the outermost type is RuntimeError (:275-277)
            if not (updated_original.shape == watermarked.shape == ground_truth_presence.shape):
                raise ValueError("Input tensors must have the same shape. Got: "
                                 f"updated_original={updated_original.shape}, watermarked={watermarked.shape}, "
                                 f"ground_truth_presence={ground_truth_presence.shape}")
            B, _, T = watermarked.shape
            plan = self.plan(T)
            wm, og, gt = _gather(updated_original, watermarked, ground_truth_presence, None, plan)
            self.stats = {k: 0.0 for k in self.stats}
            self.stats["unchanged"] = 0.0
            key = {"unchanged_short": "shuffle"}.get(plan.method, plan.method)
            self.stats[key] = 100.0
            method = "unchanged" if plan.method == "unchanged_short" else plan.method
            return AudioSignal(wm, self.sample_rate), og, gt, self.stats, method
        except Exception as e:  # noqa: BLE001
            raise RuntimeError(f"Failed to apply augmentation: {e}") from e


@torch.no_grad()
def augment(original: torch.Tensor, watermarked: torch.Tensor, loc: Optional[LocalizationPlan],
            seq: Optional[SequencePlan]):
    """`AudioWatermarking._apply_augmentations` (model/watermarking.py:485-521) as ONE pass: the
    localization table and the sequence map compose inside the gather.  Returns (watermarked_augmented,
    mask, updated_original), all [B, 1, T_out] fp32 on the device."""
    wm, og, gt = _gather(original, watermarked, None, loc, seq)
    return wm, gt, og


# --------------------------------------------------------------------------------------------- effects
SUPPORTED_EFFECTS = ("identity", "amplitude_scaling", "quantization", "random_noise", "white_noise",
                     "sample_suppression", "median_filter", "lowpass_filter", "highpass_filter", "bandpass_filter",
                     "resample", "speed")
EPSILON = 1e-5   # utils/effect_augmentation.py:92


def julius_lowpass_taps(cutoffs: Sequence[float], zeros: float = 8) -> torch.Tensor:
    """The windowed-sinc low-pass filters of `julius.LowPassFilters(cutoffs, zeros=8)` (julius is a
    third-party dependency of the reference, unpinned in requirements.txt and absent here; this restates
    its published design): half = int(zeros / min(cutoff > 0) / 2) taps on each side, Hann window
    (periodic=False), h = 2 f_c * window * sinc(2 pi f_c n), normalised to sum 1.  Same fp32 torch ops as
    julius, so the taps are bit-identical to the ones it would build.  Returns [len(cutoffs), 2*half+1]."""
    cutoffs = [float(c) for c in cutoffs]
    if min(cutoffs) < 0:
        raise ValueError("Minimum cutoff must be larger than zero.")
    if max(cutoffs) > 0.5:
        raise ValueError("A cutoff above 0.5 does not make sense.")
    half = int(zeros / min(c for c in cutoffs if c > 0) / 2)     # ValueError on an empty sequence, like julius
    window = torch.hann_window(2 * half + 1, periodic=False)
    time = torch.arange(-half, half + 1)
    out = []
    for c in cutoffs:
        if c == 0:
            out.append(torch.zeros(2 * half + 1))
            continue
        arg = 2 * c * math.pi * time
        sinc = torch.where(arg == 0, torch.tensor(1.0), torch.sin(arg) / arg)
        f = 2 * c * window * sinc
        out.append(f / f.sum())
    return torch.stack(out)


def sinc_resample_kernel(orig_freq: int, new_freq: int, lowpass_filter_width: int = 6, rolloff: float = 0.99):
    """The polyphase filter bank of `torchaudio.transforms.Resample(orig_freq, new_freq)` (sinc_interp_hann, the
    defaults the reference uses, utils/effect_augmentation.py:1481-1493): restates torchaudio's
    `_get_sinc_resample_kernel` with the same float64 torch ops, so the fp32 taps are bit-identical to torchaudio's
    (tests/test_validation_oracle.py checks that when torchaudio is importable).
    Returns (taps [new, 2*width + orig] fp32, width, orig, new) with the rates divided by their gcd."""
    g = math.gcd(int(orig_freq), int(new_freq))
    orig, new = int(orig_freq) // g, int(new_freq) // g
    base_freq = min(orig, new) * rolloff
    width = math.ceil(lowpass_filter_width * orig / base_freq)
    idx = torch.arange(-width, width + orig, dtype=torch.float64)[None, None] / orig
    t = torch.arange(0, -new, -1)[:, None, None] / new + idx     # integer arange / int -> float32 phases, as in torchaudio
    t *= base_freq
    t = t.clamp_(-lowpass_filter_width, lowpass_filter_width)
    window = torch.cos(t * math.pi / lowpass_filter_width / 2) ** 2
    t *= math.pi
    scale = base_freq / orig
    kernels = torch.where(t == 0, torch.tensor(1.0, dtype=torch.float64), t.sin() / t)
    kernels *= window * scale
    return kernels.to(torch.float32)[:, 0, :].contiguous(), width, orig, new


def _resample(x: torch.Tensor, orig_freq: int, new_freq: int, stretch_to: Optional[int] = None) -> torch.Tensor:
    """x [B, C, T] -> torchaudio-style resampled [B, C, ceil(T*new/orig)]; stretch_to = L additionally interpolates the
    result linearly to L samples (the `speed` effect's length restore) in the same kernel."""
    xin = _need_cuda(x, "audio").float().contiguous()
    B, Cn, T = xin.shape
    if orig_freq == new_freq and stretch_to in (None, T):
        return xin
    taps, width, orig, new = sinc_resample_kernel(orig_freq, new_freq)
    T_mid = -(-T * new // orig)
    T_out = T_mid if stretch_to is None else int(stretch_to)
    out = torch.empty(B, Cn, T_out, device=xin.device, dtype=torch.float32)
    h = taps.to(xin.device)
    _lib.check(_lib.lib().wv_effect_resample(_ptr(xin), _ptr(h), B * Cn, T, orig, new, width, T_mid, T_out,
                                             0 if stretch_to is None else 1, _ptr(out), _stream(xin.device)),
               "wv_effect_resample")
    return out


def _fir(x: torch.Tensor, taps: torch.Tensor, subtract: bool) -> torch.Tensor:
    xin = _need_cuda(x, "audio").float().contiguous()
    B, Cn, T = xin.shape
    h = taps.to(device=xin.device, dtype=torch.float32).contiguous()
    out = torch.empty_like(xin)
    _lib.check(_lib.lib().wv_effect_fir(_ptr(xin), _ptr(h), int(h.numel()), B * Cn, T, 1 if subtract else 0, _ptr(out),
                                        _stream(xin.device)), "wv_effect_fir")
    return out


def _pointwise(effect: int, x: torch.Tensor, p0: float, noise: Optional[torch.Tensor] = None, seed: int = 0):
    x = _need_cuda(x, "audio").float().contiguous()
    out = torch.empty_like(x)
    nz = None
    if noise is not None:
        nz = _need_cuda(noise, "noise").float().contiguous()
        if nz.shape != x.shape:
            raise ValueError("noise must have the shape of the audio")
    _lib.check(_lib.lib().wv_effect_pointwise(effect, _ptr(x), x.numel(), float(p0), _ptr(nz),
                                              int(seed) & 0xFFFFFFFFFFFFFFFF, _ptr(out), _stream(x.device)),
               "wv_effect_pointwise")
    return out


@torch.no_grad()
def apply_effect(audio: torch.Tensor, effect_type: str, sample_rate: int = 16000,
                 mask: Optional[torch.Tensor] = None, **params: Any) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
    """utils/effect_augmentation.py:2409-2636 `apply_effect(audio, effect_type, sample_rate, mask, **params)`
    -> (audio, mask) for the codec-free effects.  Extra keyword arguments of this implementation:
    `noise` (the N(0,1) draw to use instead of the in-kernel generator), `seed`, `indices` (the
    [B, k] samples to suppress instead of a fresh `torch.randperm` draw)."""
    if effect_type not in SUPPORTED_EFFECTS:
        raise NotImplementedError(
            f"effect {effect_type!r} is not implemented on the GPU path (supported: {', '.join(SUPPORTED_EFFECTS)})")
    x = _need_cuda(audio, "audio")
    if x.dim() != 3:
        raise ValueError(f"Expected [batch, channels, time] audio, got {tuple(x.shape)}")
    if effect_type == "identity":                                           # :1364-1379
        return audio, mask
    if effect_type == "amplitude_scaling":                                  # :2000-2028
        return _pointwise(1, x, float(params.get("scale", 1.0))), mask
    if effect_type == "quantization":                                       # :2030-2059, 1090-1111
        bit_depth = int(params.get("bit_depth", 16))
        if not 1 <= bit_depth <= 32:
            raise ValueError(f"Bit depth must be between 1 and 32, got {bit_depth}")
        if bit_depth == 1:
            raise ValueError("bit_depth=1 divides by zero in the reference (max_val = 0)")
        return _pointwise(2, x, float(2 ** (bit_depth - 1) - 1)), mask
    if effect_type in ("random_noise", "white_noise"):                      # :2105-2133, 2338-2368
        std = float(params.get("noise_std", 0.001 if effect_type == "random_noise" else 0.01))
        if std < 0:
            raise ValueError(f"Noise std must be non-negative, got {std}")
        noise = params.get("noise")
        if noise is not None:
            return _pointwise(3, x, std, noise=noise), mask
        seed = params.get("seed")
        if seed is None:
            seed = int(torch.randint(0, 2 ** 62, (1,)).item())
        return _pointwise(4, x, std, seed=seed), mask
    if effect_type == "sample_suppression":                                 # :2061-2103
        frac = float(params.get("suppression_percentage", 0.1))
        if not 0 <= frac <= 1:
            raise ValueError(f"Suppression percentage must be between 0 and 1, got {frac}")
        B, Cn, T = x.shape
        k = int(T * frac)
        idx = params.get("indices")
        if idx is None:   # one torch.randperm per (clip, channel), as the reference draws them
            idx = torch.stack([torch.randperm(T)[:k] for _ in range(B * Cn)]) if k else torch.zeros(B * Cn, 0, dtype=torch.long)
        idx = torch.as_tensor(idx, dtype=torch.long).reshape(B * Cn, -1).to(x.device).contiguous()
        out = x.float().clone()
        m = None
        if mask is not None:
            m = _need_cuda(mask, "mask")
            if m.dtype != torch.float32 or not m.is_contiguous():
                raise ValueError("mask must be a contiguous fp32 tensor (it is updated in place, as in the reference)")
        _lib.check(_lib.lib().wv_effect_suppress(_ptr(out), _ptr(m), _ptr(idx), B * Cn, T, int(idx.shape[1]),
                                                 _stream(x.device)), "wv_effect_suppress")
        return out, mask
    if effect_type == "resample":                                           # :1451-1502
        new_sr = params.get("new_sample_rate")
        if not isinstance(new_sr, int) or isinstance(new_sr, bool) or new_sr <= 0:
            raise ValueError(f"new_sample_rate must be positive int, got {new_sr}")
        # down to new_sample_rate and back up: two torchaudio.transforms.Resample passes (windowed-sinc polyphase)
        return _resample(_resample(x, sample_rate, new_sr), new_sr, sample_rate), mask
    if effect_type == "speed":                                              # :1381-1449
        speed = params.get("speed", 1.0)
        if isinstance(speed, tuple):
            import random
            speed = random.uniform(*speed)
        if speed <= 0:            # the reference raises inside its try block, logs and returns the input unchanged
            return audio, mask
        # SoX `speed s` + `rate sr` = resampling from sr*s to sr (T/s samples), then the reference stretches the result
        # back to the input length by linear interpolation (mode 'stretch', :187-215); the mask keeps its length.
        # SoX is a third-party binary absent here (PARITY UNPINNED for its `rate` filter): the resampling step uses the
        # same windowed-sinc design as torchaudio.transforms.Resample(round(sr*s), sr).
        src_rate = int(round(sample_rate * float(speed)))
        if src_rate == sample_rate:
            return audio, mask
        return _resample(x, src_rate, sample_rate, stretch_to=x.shape[-1]), mask
    if effect_type in ("lowpass_filter", "highpass_filter"):               # :1684-1770
        # the reference divides by the Nyquist frequency although julius expects cycles per sample: reproduced
        cutoff_freq = float(params.get("cutoff_freq", 3000 if effect_type == "lowpass_filter" else 500))
        nyquist = sample_rate / 2
        cutoff = max(0.0, min(cutoff_freq, nyquist - EPSILON)) / nyquist
        try:
            taps = julius_lowpass_taps([cutoff])[0]
        except Exception:  # noqa: BLE001  (julius raises -> the reference logs and returns the input unchanged)
            return audio, mask
        return _fir(x, taps, subtract=effect_type == "highpass_filter"), mask
    if effect_type == "bandpass_filter":                                    # :1772-1871 (ValueErrors propagate)
        if "cutoff_freq_low" not in params or "cutoff_freq_high" not in params:
            raise TypeError("bandpass_filter needs cutoff_freq_low and cutoff_freq_high")
        lo_f, hi_f = float(params["cutoff_freq_low"]), float(params["cutoff_freq_high"])
        nyquist = sample_rate / 2
        if lo_f < 0:
            raise ValueError(f"Low cutoff frequency must be non-negative, got {lo_f} Hz")
        if hi_f < 0:
            raise ValueError(f"High cutoff frequency must be non-negative, got {hi_f} Hz")
        lo_a, hi_a = max(0.0, min(lo_f, nyquist - EPSILON)), max(0.0, min(hi_f, nyquist - EPSILON))
        if lo_a >= hi_a:
            raise ValueError(f"Low cutoff {lo_a} Hz must be less than high cutoff {hi_a} Hz")
        lo, hi = lo_a / nyquist, hi_a / nyquist
        if not (0.0 < lo < 1.0) or not (0.0 < hi < 1.0):
            raise ValueError(f"Normalized cutoffs must be between 0 and 1. Got low: {lo}, high: {hi}")
        taps = julius_lowpass_taps([lo, hi])          # julius.BandPassFilter: lowpass(high) - lowpass(low), shared length
        return _fir(x, taps[1] - taps[0], subtract=False), mask
    # median_filter                                                         # :1873-1902, 1246-1312
    k = int(params.get("kernel_size", 3))
    if k < 1:
        raise ValueError(f"Kernel size must be positive, got {k}")
    if k % 2 == 0:
        k += 1
    xin = x.float().contiguous()
    out = torch.empty_like(xin)
    B, Cn, T = xin.shape
    _lib.check(_lib.lib().wv_effect_median(_ptr(xin), B * Cn, T, k, _ptr(out), _stream(x.device)), "wv_effect_median")
    return out, mask


# --------------------------------------------------------------------------------------------- driver
class ValidationPipeline(nn.Module):
    """`AudioWatermarking._forward_valid` (model/watermarking.py:443-483) on the device:
    wm = G(x, msg); y = x + wm; (y_aug, mask, x_upd) = augment(x, y); for each effect:
    y_e, mask_e = effect(y_aug, mask); detector bits (masked decode) + locator mask ->
    BER (scripts/evaluate.py:442-516) and mIoU (:591-665) from six exact integer counters."""

    def __init__(self, generator, detector, locator, sample_rate: int = 16000, window_duration: float = 0.1,
                 effects: Sequence[Tuple[str, Dict[str, Any]]] = (("identity", {}),)):
        super().__init__()
        self.generator, self.detector, self.locator = generator, detector, locator
        self.sample_rate = sample_rate
        self.localization_augmenter = LocalizationAugmentation(sample_rate, window_duration)
        self.seq_augmenter = SequenceAugmentation(sample_rate)
        self.effects = list(effects)

    @torch.no_grad()
    def forward(self, signal, msg: torch.Tensor):
        x = _need_cuda(signal.audio_data, "audio")
        B, _, T = x.shape
        wm, y, _ = self.generator.embed_batch(x, msg)
        loc = self.localization_augmenter.plan(B, T)
        seq = self.seq_augmenter.plan(T)
        y_aug, mask, x_upd = augment(x, y, loc, seq)
        results: Dict[str, Dict[str, Any]] = {}
        for name, params in self.effects:
            m_in = mask.clone() if name == "sample_suppression" else mask
            y_e, m_e = apply_effect(y_aug, name, sample_rate=self.sample_rate, mask=m_in, **params)
            det = self.detector.detect_batch(y_e, presence=m_e)
            locm = self.locator.locate_batch(y_e)["mask"]
            counters = metric_counters(det["bits"], det["valid"], msg, locm, m_e)
            ber, miou = ber_miou(counters)
            results[name] = dict(bits=det["bits"], locator_mask=locm, mask=m_e, counters=counters, ber=ber, miou=miou)
        stats = {**loc.stats, "sequence_method": seq.method}
        return make_like(signal, wm), make_like(signal, y), results, stats
