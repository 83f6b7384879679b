"""CPU: host glue of the public API - WatermarkID mappings (golden values produced by the
reference's own waveverify/watermark_id.py), message <-> tensor helpers, WAV round trip."""
import json
import os
from datetime import datetime

import numpy as np
import pytest
import torch

from helpers import GOLDEN
from waveverify_b200 import WatermarkID, load_audio, message_to_tensor, save_audio, tensor_to_message


def test_watermark_id_matches_reference_mappings():
    cases = json.load(open(os.path.join(GOLDEN, "watermark_id_cases.json")))
    assert len(cases) >= 15
    for fn, args, bits, text in cases:
        a = [bytes.fromhex(x[6:]) if isinstance(x, str) and x.startswith("bytes:") else x for x in args]
        if fn == "for_timestamp":
            a = [datetime.fromisoformat(a[0])]
        w = getattr(WatermarkID, fn)(*a)
        assert w.bits == bits and str(w) == text, (fn, args)
        assert WatermarkID.custom(w.to_int()) == w and WatermarkID.custom(w.to_bytes()) == w
        assert len({w, WatermarkID(bits)}) == 1


def test_watermark_id_validation():
    with pytest.raises(ValueError):
        WatermarkID.custom("101")
    with pytest.raises(ValueError):
        WatermarkID.custom(70000)
    with pytest.raises(ValueError):
        WatermarkID.custom(b"\x01")
    with pytest.raises(TypeError):
        WatermarkID.custom(1.5)
    with pytest.raises(ValueError):
        WatermarkID.for_timestamp(datetime(2023, 1, 1))
    with pytest.raises(ValueError):
        WatermarkID.for_creator("")


def test_message_tensor_round_trip():
    t = message_to_tensor("1010000011110001")
    assert t.shape == (1, 16) and t.dtype == torch.float32
    assert tensor_to_message(t) == "1010000011110001"
    assert message_to_tensor([1, 0] * 8).tolist() == [[1.0, 0.0] * 8]
    probs = torch.full((2, 16, 50), 0.5)                 # ties decode to 1 (>=), first batch item only
    probs[0, 3] = 0.49
    assert tensor_to_message(probs) == "1110" + "1" * 12
    for bad in ("10", "1010000011110002"):
        with pytest.raises(ValueError):
            message_to_tensor(bad)
    with pytest.raises(TypeError):
        message_to_tensor(5)
    with pytest.raises(ValueError):
        tensor_to_message(torch.zeros(2, 2, 2, 2))


def test_wav_round_trip(tmp_path):
    x = torch.from_numpy((0.3 * np.sin(np.arange(16000) * 0.05)).astype(np.float32))
    p = tmp_path / "a" / "tone.wav"
    save_audio(x * 5, p, 16000)                            # clamps to [-1, 1] like the reference
    y, sr = load_audio(p, 16000)
    assert sr == 16000 and y.shape == (1, 16000)
    assert float((y[0] - torch.clamp(x * 5, -1, 1)).abs().max()) < 2e-4
    with pytest.raises(FileNotFoundError):
        load_audio(tmp_path / "missing.wav")


def test_load_audio_resamples_like_the_reference(tmp_path):
    """waveverify/utils.py:205-216: mono mix-down, then torchaudio.transforms.Resample(sr, 16000)."""
    torchaudio = pytest.importorskip("torchaudio")
    import wave
    from waveverify_b200 import load_audio
    rng = np.random.RandomState(2)
    sr, n = 22050, 4410
    pcm = (rng.standard_normal((n, 2)) * 3000).astype("<i2")
    p = tmp_path / "stereo.wav"
    with wave.open(str(p), "wb") as f:
        f.setnchannels(2); f.setsampwidth(2); f.setframerate(sr); f.writeframes(pcm.tobytes())
    wav, out_sr = load_audio(p)
    mono = torch.from_numpy(pcm.astype(np.float32) / 32768.0).T.mean(0, keepdim=True)
    want = torchaudio.transforms.Resample(sr, 16000)(mono)
    assert out_sr == 16000 and wav.shape == want.shape
    assert torch.allclose(wav, want, atol=1e-6)
