"""Achieved HBM GB/s of the validation-path kernels at 512 x 10 s clips
(BASELINE config 3 size per GPU at N = 8 is 64 x 10 s; 512 x 10 s = 328 MB per tensor > L2).
CUDA events on the launching stream, 3 warm-ups + 10 timed launches each."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from waveverify_b200 import validation as V  # noqa: E402

PEAK = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))


def timed(fn, n=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e-3


def main():
    B, T = 512, 160000
    x = 0.1 * torch.randn(B, 1, T, device="cuda")
    y = x + 0.01 * torch.randn_like(x)
    nbytes = x.numel() * 4
    np.random.seed(0)
    torch.manual_seed(0)
    loc = V.LocalizationAugmentation(16000, 0.1).plan(B, T)
    rows = []
    for name, seq in (("augment(loc + reverse)", V.SequencePlan("reverse", V.SEQ_REVERSE)),
                      ("augment(loc + shift)", V.SequencePlan("circular_shift", V.SEQ_SHIFT, a=12345)),
                      ("augment(loc + shuffle)", V.SequencePlan("shuffle", V.SEQ_SHUFFLE, c=8000,
                                                               perm=np.random.permutation(T // 8000).astype(np.int32))),
                      ("augment(loc only)", None)):
        t = timed(lambda: V.augment(x, y, loc, seq))
        rows.append((name, 5 * nbytes, t))          # reads x, y; writes wm, gt, original
    for name, kw, traffic in (("amplitude_scaling", dict(scale=0.5), 2), ("quantization", dict(bit_depth=8), 2),
                              ("white_noise (in-kernel Philox)", dict(noise_std=0.01, seed=1), 2),
                              ("random_noise (supplied draw)", dict(noise_std=0.01, noise=x), 3),
                              ("median_filter k=3", dict(kernel_size=3), 2), ("median_filter k=9", dict(kernel_size=9), 2),
                              ("lowpass_filter 3 kHz (21 taps)", dict(cutoff_freq=3000), 2),
                              ("highpass_filter 500 Hz (129 taps)", dict(cutoff_freq=500), 2),
                              ("bandpass_filter 1-3 kHz (65 taps)", dict(cutoff_freq_low=1000, cutoff_freq_high=3000), 2)):
        eff = name.split(" ")[0]
        t = timed(lambda: V.apply_effect(y, eff, **kw))
        rows.append((name, traffic * nbytes, t))
    peak = PEAK.get("hbm_gbs") or 6451.5
    print(f"| kernel ({B} x {T // 16000} s clips, {nbytes / 1e6:.0f} MB per tensor) | algorithmic MB | us | GB/s | of measured copy peak ({peak:.0f} GB/s) |")
    print("|---|---|---|---|---|")
    for name, bts, t in rows:
        print(f"| {name} | {bts / 1e6:.0f} | {t * 1e6:.1f} | {bts / t / 1e9:.0f} | {100 * bts / t / 1e9 / peak:.0f} % |")
    print("\n(times include the torch.empty output allocations of the Python shim; the kernels are the only device work)")


if __name__ == "__main__":
    main()
