"""CPU: the oracle restatement reproduces the committed outputs of the reference itself."""
import os

import numpy as np
import pytest
import torch

import wv_oracle as O
from helpers import fixture_weights, golden_cases, load_case, oracle_cfg

CASES = golden_cases()


def test_golden_fixtures_present():
    assert len(CASES) >= 5


@pytest.mark.parametrize("path", CASES, ids=[os.path.basename(p)[:-4] for p in CASES])
def test_oracle_matches_reference_outputs(path):
    z = load_case(path)
    if z["x"].shape[-1] * z["x"].shape[0] > 40000:
        torch.set_num_threads(8)
    zi, ws = bool(z["zero_init"]), int(z["wseed"])
    x = torch.from_numpy(z["x"]); msg = torch.from_numpy(z["msg"])
    W = {}
    cfgs = {}
    for kind in ("generator", "detector", "locator"):
        c, sd = fixture_weights(kind, zi, ws)
        W[kind] = O.fold_state_dict(sd); cfgs[kind] = oracle_cfg(c)
    with torch.no_grad():
        taps = {}
        wm = O.generator_forward(x, msg, W["generator"], cfgs["generator"], taps)
        y = x + wm
        det = O.detector_forward(torch.from_numpy(z["y"]), W["detector"], cfgs["detector"])
        loc = O.locator_forward(torch.from_numpy(z["y"]), W["locator"], cfgs["locator"])
    # fp32 tolerance: both sides are fp32 with different summation order
    np.testing.assert_allclose(wm.numpy(), z["wm"], atol=2e-6, rtol=0)
    np.testing.assert_allclose(y.numpy(), z["y"], atol=2e-6, rtol=0)
    np.testing.assert_allclose(taps["latent"].numpy(), z["latent"], atol=5e-5, rtol=0)
    d = int(z["det_decim"])
    np.testing.assert_allclose(det[:, :, ::d].numpy(), z["det_logits_decim"], atol=1e-4, rtol=1e-5)
    np.testing.assert_allclose(loc.numpy(), z["loc_logits"], atol=2e-5, rtol=1e-5)
    bits, avg, conf, valid = O.decode_bits(det)
    np.testing.assert_allclose(avg.numpy(), z["det_avg"], atol=2e-6)
    np.testing.assert_allclose(conf.numpy(), z["det_conf"], atol=2e-6)
    # bits/masks: exact wherever the reference value is not within fp32 noise of the threshold
    safe = np.abs(z["det_avg"] - 0.5) > 1e-5
    assert (bits.numpy() == z["det_bits"])[safe].all()
    safe = np.abs(z["loc_logits"] - 0.5) > 1e-4
    assert (O.locator_mask(loc).numpy() == z["loc_mask"])[safe].all()
    assert (O.detector_postprocess(det).numpy() == z["det_post"]).all()


def test_ber_known_answer():
    """scripts/evaluate.py:672-787 quasi-KAT: +-2 logits + 0.5 noise => BER == 0, full and half mask."""
    g = torch.Generator().manual_seed(0)
    B, Wb, T = 4, 16, 2000
    msg = torch.randint(0, 2, (B, Wb), generator=g)
    logits = (msg.float() * 4 - 2).unsqueeze(-1) + 0.5 * torch.randn(B, Wb, T, generator=g)
    for mask in (None, torch.cat([torch.ones(B, 1, T // 2), torch.zeros(B, 1, T - T // 2)], 2)):
        bits, avg, conf, valid = O.decode_bits(logits, mask)
        c = O.metric_counters(bits, valid, msg, torch.zeros(B, 1, T), torch.zeros(B, 1, T))
        assert c[0] == 0 and c[1] == B * Wb
    # empty mask => no valid bits => BER defined as 0
    bits, avg, conf, valid = O.decode_bits(logits, torch.zeros(B, 1, T))
    c = O.metric_counters(bits, valid, msg, torch.zeros(B, 1, T), torch.zeros(B, 1, T))
    assert c[1] == 0 and O.ber_miou_from_counters(c)[0] == 0.0


def test_miou_counters():
    p = torch.tensor([[1, 1, 0, 0, 1, 0]], dtype=torch.uint8)
    g = torch.tensor([[1, 0, 0, 1, 1, 0]], dtype=torch.uint8)
    c = O.metric_counters(torch.zeros(1, 1, dtype=torch.uint8), torch.ones(1, 1, dtype=torch.bool),
                          torch.zeros(1, 1), p, g)
    assert c[2:] == [2, 4, 2, 4]
    assert abs(O.ber_miou_from_counters(c)[1] - 0.5) < 1e-12
    # empty-union rule (scripts/evaluate.py:640-653)
    z = torch.zeros(1, 4, dtype=torch.uint8)
    c = O.metric_counters(torch.zeros(1, 1, dtype=torch.uint8), torch.ones(1, 1, dtype=torch.bool),
                          torch.zeros(1, 1), z, z)
    assert O.ber_miou_from_counters(c)[1] == 1.0
