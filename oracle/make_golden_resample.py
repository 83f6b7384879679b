"""Generate tests/golden/validation/resample.npz (run in the build container only).

`resample`: outputs of the REFERENCE ITSELF - utils/effect_augmentation.py loaded unmodified by path (as in
make_golden_validation.py), `apply_effect(x, "resample", new_sample_rate=...)`, which calls
torchaudio.transforms.Resample twice (:1451-1502).
`speed`: the reference shells out to SoX through torchaudio.sox_effects, which this image does not have (the
reference then logs the failure and returns its input unchanged, :1443-1449) - PARITY UNPINNED against SoX.  The stored
vectors are torchaudio.functional.resample(x, round(sr*speed), sr) followed by the reference's own
AudioProcessor.adjust_audio_length(..., mode='stretch') (:187-215), i.e. the reference's post-processing applied to
torchaudio's windowed-sinc resampler.
"""
import os
import sys

import numpy as np
import torch
import torchaudio

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from make_golden_validation import inputs, load_ref  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "validation", "resample.npz")


def main():
    FX = load_ref()["effect_augmentation"]
    out = {}
    cases = [(31, 2, 1000, 16000, 32000), (32, 2, 997, 16000, 8000), (33, 1, 640, 16000, 22050), (34, 3, 333, 16000, 12000)]
    out["resample_cases"] = np.array(cases, np.int64)
    for i, (seed, B, T, sr, new_sr) in enumerate(cases):
        x, _ = inputs(seed, B, T)
        y, _ = FX.apply_effect(torch.from_numpy(x), "resample", sample_rate=sr, new_sample_rate=new_sr)
        out[f"resample{i}"] = y.numpy()
    sp_cases = [(41, 2, 1000, 16000, 0.8), (42, 1, 777, 16000, 1.25), (43, 2, 500, 16000, 0.9)]
    out["speed_cases"] = np.array([(a, b, c, d, int(round(e * 1000))) for a, b, c, d, e in sp_cases], np.int64)
    for i, (seed, B, T, sr, sp) in enumerate(sp_cases):
        x, _ = inputs(seed, B, T)
        y1 = torchaudio.functional.resample(torch.from_numpy(x), int(round(sr * sp)), sr)
        y = FX.AudioProcessor.adjust_audio_length(y1, T, mode="stretch")
        out[f"speed{i}"] = y.numpy()
    np.savez_compressed(OUT, **out)
    print("wrote", os.path.normpath(OUT), os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
