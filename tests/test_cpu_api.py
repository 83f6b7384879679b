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


# ---- checkpoint formats (waveverify/core.py:141-168, 225-469): host-side model construction + load ----------
def _fixture_models(bias, zero_init):
    from waveverify_b200 import Detector, Generator, Locator, fixture_state_dict
    loc = dict(dimension=64, channels_enc=32, n_residual_enc=1, strides=[8, 4])
    out = {}
    for name, cls, kw in (("generator", Generator, {}), ("detector", Detector, {}), ("locator", Locator, loc)):
        m = cls(**{**kw, "bias": bias, "zero_init": zero_init})
        m.load_state_dict(fixture_state_dict(m.cfg, 3))
        out[name] = m
    return out


def _same_params(a, b):
    sa, sb = a.state_dict(), b.state_dict()
    assert list(sa.keys()) == list(sb.keys())
    for k in sa:
        assert torch.equal(sa[k], sb[k]), k


def test_atomic_checkpoint_builds_models_from_its_config(tmp_path):
    """conf/base.yml trains with zero_init=False / bias as configured: the architecture must come from
    checkpoint['config'] (core.py:230-276), not from the class defaults (zero_init=True adds scale parameters)."""
    from waveverify_b200 import WaveVerify
    src = _fixture_models(bias=True, zero_init=False)
    config = {"Generator.zero_init": False, "Generator.bias": True, "Generator.channels_enc": 64,
              "Detector.zero_init": False, "Detector.bias": True, "Detector.nbits": 16,
              "Locator.zero_init": False, "Locator.bias": True, "Locator.nbits": 16, "Locator.strides": [8, 4],
              "Locator.channels_enc": 32, "Locator.dimension": 64, "Locator.n_residual_enc": 1, "train.lr": 1e-4}
    f = tmp_path / "latest.pth"
    torch.save({"models": {k: m.state_dict() for k, m in src.items()}, "config": config, "step": 7}, str(f))
    g, d, l = WaveVerify.build_models(f)
    assert g.cfg.zero_init is False and not any(k.endswith("scale_param") for k in g.state_dict())
    for m, name in ((g, "generator"), (d, "detector"), (l, "locator")):
        _same_params(m, src[name])
    # a directory holding best.pth / latest.pth is the reference's normal layout (core.py:141-168): best.pth wins
    other = _fixture_models(bias=True, zero_init=True)
    torch.save({"models": {k: m.state_dict() for k, m in other.items()}}, str(tmp_path / "best.pth"))
    g2, d2, l2 = WaveVerify.build_models(tmp_path)
    assert g2.cfg.zero_init is True                       # no config stored: inferred from the *scale_param keys
    _same_params(g2, other["generator"])
    _same_params(l2, other["locator"])


def test_checkpoint_without_config_infers_bias_and_zero_init(tmp_path):
    from waveverify_b200 import WaveVerify
    from waveverify_b200.fold import fold_state_dict
    src = _fixture_models(bias=True, zero_init=False)
    # parametrizations removed before saving (scripts/train.py:1624-1629): plain `weight` keys
    plain = {k: fold_state_dict(m.state_dict()) for k, m in src.items()}
    f = tmp_path / "ck.pth"
    torch.save({"models": plain}, str(f))
    g, d, l = WaveVerify.build_models(f)
    assert g.cfg.zero_init is False and g.cfg.bias is True
    a, b = fold_state_dict(g.state_dict()), plain["generator"]
    assert set(a) == set(b)
    for k in a:
        assert torch.allclose(a[k], b[k], atol=1e-6), k
    # a checkpoint that does not match the architecture fails loudly instead of leaving parameters at their init
    broken = {k: dict(v) for k, v in plain.items()}
    drop = next(k for k in broken["detector"] if k.startswith("encoder.blocks.0.0.block.1"))
    del broken["detector"][drop]
    torch.save({"models": broken}, str(tmp_path / "broken.pth"))
    with pytest.raises(RuntimeError, match="parameters missing"):
        WaveVerify.build_models(tmp_path / "broken.pth")


def test_legacy_checkpoint_layout(tmp_path):
    """<dir>/{generator,detector,locator}/model.pth, strict loading (core.py:428-469)."""
    from waveverify_b200 import WaveVerify
    src = _fixture_models(bias=True, zero_init=True)
    for name, m in src.items():
        (tmp_path / name).mkdir()
        torch.save(m.state_dict(), str(tmp_path / name / "model.pth"))
    g, d, l = WaveVerify.build_models(tmp_path)
    for m, name in ((g, "generator"), (d, "detector"), (l, "locator")):
        _same_params(m, src[name])
    (tmp_path / "locator" / "model.pth").unlink()
    with pytest.raises(FileNotFoundError):
        WaveVerify.build_models(tmp_path)


def test_plain_weight_with_zero_rows_folds_to_zeros():
    """zero-initialised '1x1_zero' / pruned channels: g*v/||v|| must stay 0, not 0*0/0 = NaN."""
    from waveverify_b200.fold import _WN_G, _WN_V, fold_state_dict, unfold_plain_weight
    w = torch.randn(6, 4, 5)
    w[2] = 0
    g, v = unfold_plain_weight(w)
    out = fold_state_dict({"a" + _WN_G: g, "a" + _WN_V: v})["a.weight"]
    assert torch.isfinite(out).all() and torch.allclose(out, w, atol=1e-6) and float(out[2].abs().max()) == 0.0
