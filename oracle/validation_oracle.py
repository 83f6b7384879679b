"""CPU oracle for the validation-path augmentations and effects.

TEST INFRASTRUCTURE ONLY (see oracle/wv_oracle.py): numpy restatement, sequential slice operations
in the reference's own order; only `tests/` may import it.

Parity pinning: pinned against OUTPUTS OF THE REFERENCE ITSELF (utils/localization_augmentation.py,
utils/seq_augmentation.py, utils/effect_augmentation.py imported unmodified in the build container
by oracle/make_golden_validation.py, with stand-ins for the absent matplotlib / julius /
audiotools imports that these functions never call) -> tests/golden/validation/validation.npz, re-checked
by tests/test_validation_oracle.py.  Random draws follow the reference's call order on the
global numpy / torch generators, so seeding those reproduces the reference's selections.
"""
from __future__ import annotations

import numpy as np
import torch


def localization_augment(original: np.ndarray, watermarked: np.ndarray, segment_length: int):
    """utils/localization_augmentation.py:212-325.  [B, C, T] float32 in; returns
    (watermarked', ground_truth, updated_original, stats%)."""
    original = original.astype(np.float32).copy()
    upd = original.copy()
    wm = watermarked.astype(np.float32).copy()
    B, _, T = wm.shape
    gt = np.ones_like(wm)
    stats = dict(original_revert=0, zero_replace=0, cross_substitute=0, unchanged=0)
    total_segments = int(np.ceil(T / segment_length))                      # :259
    k = int(total_segments * 0.20)                                         # :260
    for b in range(B):
        starts = np.arange(0, T, segment_length)                           # :269
        for start in np.random.choice(starts, k, replace=False):           # :272-276
            end = min(start + segment_length, T)
            p = np.random.rand()                                           # :281
            if p < 0.33:                                                   # :126-149
                wm[b, :, start:end] = original[b, :, start:end]
                gt[b, :, start:end] = 0
                stats["original_revert"] += end - start
            elif p < 0.66:                                                 # :151-175
                wm[b, :, start:end] = 0
                upd[b, :, start:end] = 0
                gt[b, :, start:end] = 0
                stats["zero_replace"] += end - start
            elif B >= 2:                                                   # :177-210
                other = np.random.choice([j for j in range(B) if j != b])
                wm[b, :, start:end] = original[other, :, start:end]
                upd[b, :, start:end] = original[other, :, start:end]
                gt[b, :, start:end] = 0
                stats["cross_substitute"] += end - start
    total = B * T
    stats["unchanged"] = total - stats["original_revert"] - stats["zero_replace"] - stats["cross_substitute"]
    return wm, gt, upd, {k_: float(v / total * 100) for k_, v in stats.items()}


def sequence_augment(updated_original: np.ndarray, watermarked: np.ndarray, gt: np.ndarray, sample_rate: int):
    """utils/seq_augmentation.py:100-277.  Returns (watermarked', updated_original', gt', method)."""
    B, C, T = watermarked.shape
    r = np.random.rand()                                                   # :154
    tensors = [watermarked, updated_original, gt]
    if r < 0.3:                                                            # :165-170
        return (*[np.ascontiguousarray(t[:, :, ::-1]) for t in tensors], "reverse")
    if r < 0.7:                                                            # :172-178
        shift = np.random.randint(1, T)
        return (*[np.roll(t, shift, axis=2) for t in tensors], "circular_shift")
    if r < 1.0:                                                            # :181-207
        seg = int(0.5 * sample_rate)
        if T >= 2 * seg:
            n = T // seg
            perm = torch.randperm(n).numpy()
            outs = [t[:, :, :n * seg].reshape(B, C, n, seg)[:, :, perm].reshape(B, C, n * seg) for t in tensors]
            return (*outs, "shuffle")
        return (*tensors, "unchanged")
    return (*tensors, "unchanged")


def chunk_swap(x: np.ndarray, c1: int, c2: int, size: int) -> np.ndarray:
    """utils/seq_augmentation.py:236-240 (the swap itself; the reference never selects this method)."""
    y = x.copy()
    tmp = y[:, :, c1:c1 + size].copy()
    y[:, :, c1:c1 + size] = y[:, :, c2:c2 + size]
    y[:, :, c2:c2 + size] = tmp
    return y


def amplitude_scaling(x: np.ndarray, scale: float) -> np.ndarray:
    """utils/effect_augmentation.py:2022."""
    return (x.astype(np.float32) * np.float32(scale)).astype(np.float32)


def quantization(x: np.ndarray, bit_depth: int) -> np.ndarray:
    """utils/effect_augmentation.py:1103-1109: round-half-even(x * m) / m in fp32."""
    m = np.float32(2 ** (bit_depth - 1) - 1)
    return (np.rint(x.astype(np.float32) * m) / m).astype(np.float32)


def add_noise(x: np.ndarray, noise: np.ndarray, std: float) -> np.ndarray:
    """utils/effect_augmentation.py:2127-2128 / 2362-2363 with the N(0,1) draw given."""
    return (x.astype(np.float32) + noise.astype(np.float32) * np.float32(std)).astype(np.float32)


def sample_suppression(x: np.ndarray, mask, idx: np.ndarray):
    """utils/effect_augmentation.py:2084-2098; idx [B*C, k] = the randperm heads."""
    y = x.copy()
    B, C, T = x.shape
    m = None if mask is None else mask.copy()
    for b in range(B):
        for c in range(C):
            y[b, c, idx[b * C + c]] = 0
            if m is not None:
                m[b, c, idx[b * C + c]] = 0
    return y, m


def median_filter(x: np.ndarray, kernel_size: int) -> np.ndarray:
    """utils/effect_augmentation.py:1278-1307 -> scipy.signal.medfilt (zero-padded ends)."""
    if kernel_size % 2 == 0:
        kernel_size += 1
    h = kernel_size // 2
    B, C, T = x.shape
    out = np.empty_like(x)
    for b in range(B):
        for c in range(C):
            p = np.concatenate([np.zeros(h, x.dtype), x[b, c], np.zeros(h, x.dtype)])
            w = np.lib.stride_tricks.sliding_window_view(p, kernel_size)
            out[b, c] = np.sort(w, axis=1)[:, h]
    return out


# ---- julius FIR filters (PARITY UNPINNED: julius is absent here; restated from its published algorithm) ----
def julius_lowpass_taps(cutoffs, zeros=8):
    """julius.LowPassFilters.__init__ (julius/lowpass.py, v0.2.x), float32 numpy."""
    if min(cutoffs) < 0:
        raise ValueError("Minimum cutoff must be larger than zero.")
    if max(cutoffs) > 0.5:
        raise ValueError("A cutoff above 0.5 does not make sense.")
    half = int(zeros / min(c for c in cutoffs if c > 0) / 2)
    n = 2 * half + 1
    k = np.arange(n, dtype=np.float64)
    window = (0.5 - 0.5 * np.cos(2 * np.pi * k / (n - 1))) if n > 1 else np.ones(1)   # hann, periodic=False
    t = np.arange(-half, half + 1, dtype=np.float64)
    out = []
    for c in cutoffs:
        if c == 0:
            out.append(np.zeros(n))
            continue
        f = 2 * c * window * np.sinc(2 * c * t)          # np.sinc(x) = sin(pi x) / (pi x)
        out.append(f / f.sum())
    return np.stack(out)


def julius_fir(x: np.ndarray, taps: np.ndarray) -> np.ndarray:
    """julius.LowPassFilters.forward: replicate-pad by half on both sides, cross-correlate (float64 here)."""
    half = len(taps) // 2
    B, C, T = x.shape
    out = np.empty((B, C, T), np.float64)
    for b in range(B):
        for c in range(C):
            p = np.pad(x[b, c].astype(np.float64), (half, half), mode="edge")
            out[b, c] = np.correlate(p, taps.astype(np.float64), mode="valid")
    return out


def lowpass_filter(x, cutoff_freq, sample_rate):
    """utils/effect_augmentation.py:1728-1770 (cutoff normalised by the Nyquist frequency, as the reference does)."""
    nyq = sample_rate / 2
    c = max(0.0, min(cutoff_freq, nyq - 1e-5)) / nyq
    try:
        taps = julius_lowpass_taps([c])[0]
    except Exception:
        return x.astype(np.float64)
    return julius_fir(x, taps)


def highpass_filter(x, cutoff_freq, sample_rate):
    """utils/effect_augmentation.py:1684-1726; julius: input - lowpass(input)."""
    nyq = sample_rate / 2
    c = max(0.0, min(cutoff_freq, nyq - 1e-5)) / nyq
    try:
        taps = julius_lowpass_taps([c])[0]
    except Exception:
        return x.astype(np.float64)
    return x.astype(np.float64) - julius_fir(x, taps)


def bandpass_filter(x, lo_freq, hi_freq, sample_rate):
    """utils/effect_augmentation.py:1772-1871; julius.BandPassFilter: lowpass(high) - lowpass(low), shared length."""
    nyq = sample_rate / 2
    lo, hi = max(0.0, min(lo_freq, nyq - 1e-5)) / nyq, max(0.0, min(hi_freq, nyq - 1e-5)) / nyq
    taps = julius_lowpass_taps([lo, hi])
    return julius_fir(x, taps[1]) - julius_fir(x, taps[0])


# ---- resample / speed (utils/effect_augmentation.py:1381-1502) ------------------------------------------------
def sinc_resample_taps(orig_freq: int, new_freq: int, lowpass_filter_width: int = 6, rolloff: float = 0.99):
    """torchaudio.functional._get_sinc_resample_kernel (sinc_interp_hann) in float64 numpy ->
    (taps [new, 2*width + orig] float64, width, orig, new), rates divided by their gcd."""
    import math
    g = math.gcd(int(orig_freq), int(new_freq))
    orig, new = int(orig_freq) // g, int(new_freq) // g
    base = min(orig, new) * rolloff
    width = math.ceil(lowpass_filter_width * orig / base)
    idx = np.arange(-width, width + orig, dtype=np.float64)[None, :] / orig
    t = (np.arange(0, -new, -1, dtype=np.float32)[:, None] / np.float32(new)).astype(np.float64) + idx   # torchaudio: fp32 phases
    t = np.clip(t * base, -lowpass_filter_width, lowpass_filter_width)
    window = np.cos(t * math.pi / lowpass_filter_width / 2) ** 2
    t = t * math.pi
    with np.errstate(divide="ignore", invalid="ignore"):
        k = np.where(t == 0, 1.0, np.sin(t) / t)
    return (k * window * (base / orig)).astype(np.float32).astype(np.float64), width, orig, new


def sinc_resample(x: np.ndarray, orig_freq: int, new_freq: int) -> np.ndarray:
    """torchaudio.functional.resample: zero-pad (width, width + orig), strided correlation with the filter bank,
    interleave the `new` phases, keep ceil(new * T / orig) samples.  float64 accumulation."""
    if orig_freq == new_freq:
        return x.astype(np.float32)
    h, width, orig, new = sinc_resample_taps(orig_freq, new_freq)
    T = x.shape[-1]
    xp = np.pad(x.astype(np.float64), [(0, 0)] * (x.ndim - 1) + [(width, width + orig)])
    n_fr = (xp.shape[-1] - h.shape[1]) // orig + 1
    out = np.zeros(x.shape[:-1] + (n_fr, new))
    for q in range(n_fr):
        seg = xp[..., q * orig:q * orig + h.shape[1]]
        out[..., q, :] = seg @ h.T
    T_out = -(-T * new // orig)
    return out.reshape(x.shape[:-1] + (n_fr * new,))[..., :T_out].astype(np.float32)


def resample_effect(x: np.ndarray, new_sample_rate: int, sample_rate: int) -> np.ndarray:
    return sinc_resample(sinc_resample(x, sample_rate, new_sample_rate), new_sample_rate, sample_rate)


def linear_stretch(y: np.ndarray, T_out: int) -> np.ndarray:
    """torch.nn.functional.interpolate(mode='linear', align_corners=False) with its fp32 index arithmetic."""
    T_in = y.shape[-1]
    scale = np.float32(T_in) / np.float32(T_out)
    src = (np.arange(T_out, dtype=np.float32) + np.float32(0.5)) * scale - np.float32(0.5)
    src = np.maximum(src, np.float32(0))
    i0 = src.astype(np.int64)
    i1 = i0 + (i0 < T_in - 1)
    l1 = (src - i0.astype(np.float32)).astype(np.float64)
    return ((1 - l1) * y[..., i0].astype(np.float64) + l1 * y[..., i1].astype(np.float64)).astype(np.float32)


def speed_effect(x: np.ndarray, speed: float, sample_rate: int) -> np.ndarray:
    """SoX speed + rate as a windowed-sinc resample from sr*speed to sr (PARITY UNPINNED against SoX itself), then the
    reference's linear stretch back to the input length (utils/effect_augmentation.py:187-215, 583-590)."""
    src = int(round(sample_rate * speed))
    if src == sample_rate:
        return x.astype(np.float32)
    return linear_stretch(sinc_resample(x, src, sample_rate), x.shape[-1])

