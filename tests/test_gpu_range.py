"""GPU: the fp16 range check (wv_net_set_range_check): every stored fp16 tensor of a forward is scanned for values at the
saturation bound of cvt.rn.satfinite (65504) and for non-finite values."""
import numpy as np
import pytest
import torch

from helpers import BASE_KW, fixture_weights

pytestmark = pytest.mark.gpu


def build(kind):
    from waveverify_b200 import Detector, Generator, Locator
    cls = {"generator": Generator, "detector": Detector, "locator": Locator}[kind]
    _, sd = fixture_weights(kind, False, 0)
    m = cls(**{**BASE_KW[kind], "bias": True, "zero_init": False})
    m.load_state_dict(sd)
    return m.cuda(), sd


@pytest.mark.parametrize("kind", ["generator", "detector", "locator"])
def test_no_saturation_on_full_scale_audio(kind):
    """Full-scale noise and a full-scale square wave: nothing saturates, the largest stored value stays far below 65504."""
    m, _ = build(kind)
    rng = np.random.RandomState(0)
    T = 16000
    x = np.stack([np.clip(rng.standard_normal(T), -1, 1), np.sign(np.sin(2 * np.pi * 440 * np.arange(T) / 16000)),
                  0.1 * rng.standard_normal(T)]).astype(np.float32)[:, None, :]
    x = torch.from_numpy(x).cuda()
    m.set_range_check(True)
    if kind == "generator":
        m.embed_batch(x, torch.ones(3, 16, device="cuda"))
    elif kind == "detector":
        m.exact_bits = False
        m.detect_batch(x)
    else:
        m.exact = False
        m.locate_batch(x)
    r = m.range_read()
    m.set_range_check(False)
    assert r["saturated"] == 0 and r["nonfinite"] == 0, r
    assert 0.0 < r["max_abs"] < 65504.0 / 64, r


def test_saturation_is_detected():
    """conv_pre taps scaled by 1e6: its fp16 output clips at 65504 and the check counts it."""
    from waveverify_b200 import Detector
    _, sd = fixture_weights("detector", False, 0)
    key = [k for k in sd if k.startswith("encoder.conv_pre") and k.endswith("original0")]
    assert key, "conv_pre weight-norm gain not found"
    sd = {k: v.clone() for k, v in sd.items()}
    sd[key[0]] = sd[key[0]] * 1e6
    m = Detector(**{**BASE_KW["detector"], "bias": True, "zero_init": False})
    m.load_state_dict(sd)
    m = m.cuda()
    m.exact_bits = False
    m.set_range_check(True)
    m.detect_batch(0.1 * torch.randn(1, 1, 4000, device="cuda"))
    r = m.range_read()
    assert r["saturated"] > 0 or r["nonfinite"] > 0, r
    assert m.range_read()["saturated"] == 0           # the read resets the counters
