"""GPU: the CUDA path against the CPU oracle DIRECTLY at the BASELINE.json configuration sizes (the per-rank slices
of the multi-GPU configs), plus drift of the error with clip length at fixed weights.

  config 2: 64 x 1 s, embed + detect + locate                          (whole batch vs the oracle)
  config 3: 8 x 10 s = the per-rank slice of 64 x 10 s on 8 GPUs        (whole slice vs the oracle)
  config 4: Detector + Locator on 16 x 5 s clips                        (config 4 clip shape)
  config 5: a 60 s clip streamed through the Generator in 20 s chunks   (vs the oracle's whole-clip output, full length)

Tolerances as in tests/test_gpu_parity.py: bits / masks EXACT outside the fp32 band, waveforms / logits within the
stated SNR.  The oracle runs on the box's host cores (seconds per case)."""
import os

import numpy as np
import pytest
import torch

import wv_oracle as O
from helpers import BASE_KW, fixture_weights, oracle_cfg, snr_db

pytestmark = pytest.mark.gpu
_M = {}


def models():
    if not _M:
        from waveverify_b200 import Detector, Generator, Locator
        torch.set_num_threads(os.cpu_count() or 8)
        for kind, cls in (("generator", Generator), ("detector", Detector), ("locator", Locator)):
            c, sd = fixture_weights(kind, False, 0)
            m = cls(**{**BASE_KW[kind], "bias": True, "zero_init": False})
            m.load_state_dict(sd)
            _M[kind] = (m.cuda(), O.fold_state_dict(sd), oracle_cfg(c))
    return _M


def synth(B, T, seed):
    rng = np.random.RandomState(seed)
    x = torch.from_numpy((0.1 * rng.standard_normal((B, 1, T))).astype(np.float32))
    msg = torch.from_numpy(rng.randint(0, 2, (B, 16)).astype(np.int64))
    return x, msg


def check_detect_locate(y_cpu, d, l, chunk=8):
    """Detector bits / Locator mask of the CUDA path against the oracle on the same watermarked audio."""
    m = models()
    n_bits_bad = n_mask_bad = 0
    worst_avg = worst_ll = 0.0
    n_ll = n_ll_off = 0
    for b0 in range(0, y_cpu.shape[0], chunk):          # the oracle materialises [B,16,T] logits: bounded slices
        yc = y_cpu[b0:b0 + chunk]
        with torch.no_grad():
            lg_o = O.detector_forward(yc, m["detector"][1], m["detector"][2])
            ll_o = O.locator_forward(yc, m["locator"][1], m["locator"][2])
        bits_o, avg_o, _, _ = O.decode_bits(lg_o)
        avg = d["avg"][b0:b0 + chunk].cpu().numpy(); bits = d["bits"][b0:b0 + chunk].cpu().numpy()
        safe = np.abs(avg_o.numpy() - 0.5) > 1e-5
        n_bits_bad += int(((bits != bits_o.numpy()) & safe).sum())
        worst_avg = max(worst_avg, float(np.abs(avg - avg_o.numpy()).max()))
        ll = l["logits"][b0:b0 + chunk].cpu().numpy(); mask = l["mask"][b0:b0 + chunk].cpu().numpy()
        safe = np.abs(ll_o.numpy() - 0.5) > 1e-4
        n_mask_bad += int(((mask != O.locator_mask(ll_o).numpy()) & safe).sum())
        dl = np.abs(ll - ll_o.numpy())
        worst_ll = max(worst_ll, float(dl.max()))
        n_ll += dl.size; n_ll_off += int((dl > 1e-4).sum())
    assert n_bits_bad == 0, f"{n_bits_bad} decoded bits differ from the oracle outside the fp32 band"
    assert n_mask_bad == 0, f"{n_mask_bad} mask samples differ from the oracle outside the fp32 band"
    # Locator logits of the fp32-accurate net: <= 1e-4 from the oracle on all but a handful of samples.  The exceptions
    # (measured: 248 of 1 024 000 samples on the 64 x 1 s batch, max 2.0e-3, profiles/r02_precise_locator_outliers.md)
    # sit where an STFT bin nearly cancels: the log-magnitude amplifies the tensor cores' truncating fp32 accumulation
    # (the fp32 oracle is within 2e-5 of a float64 evaluation there); the masks above are still exact outside 1e-4.
    assert worst_avg <= 2e-4 and worst_ll <= 5e-3 and n_ll_off <= 1e-3 * n_ll, (worst_avg, worst_ll, n_ll_off, n_ll)


def run_embed_detect_locate(B, T, seed, chunk):
    m = models()
    x, msg = synth(B, T, seed)
    wm, y, _ = m["generator"][0].embed_batch(x.cuda(), msg.cuda())
    d = m["detector"][0].detect_batch(y)
    l = m["locator"][0].locate_batch(y, want_logits=True)
    snrs = []
    for b0 in range(0, B, chunk):
        with torch.no_grad():
            wm_o = O.generator_forward(x[b0:b0 + chunk], msg[b0:b0 + chunk], m["generator"][1], m["generator"][2])
        got = wm[b0:b0 + chunk].cpu().numpy()
        snrs.append(snr_db(wm_o.numpy(), got))
        assert np.abs(wm_o.numpy() - got).max() <= 4e-4
    assert min(snrs) >= 46.0, snrs
    assert torch.equal(y.cpu(), x + wm.cpu())
    check_detect_locate(y.cpu(), d, l, chunk)
    return min(snrs)


def test_config2_64x1s_against_oracle():
    run_embed_detect_locate(64, 16000, 21, 16)


def test_config3_rank_slice_8x10s_against_oracle():
    run_embed_detect_locate(8, 160000, 22, 2)


def test_config4_shape_detector_locator_16x5s_against_oracle():
    m = models()
    x, _ = synth(16, 80000, 23)
    y = x.cuda()                                          # config 4: any fixed input y
    d = m["detector"][0].detect_batch(y)
    l = m["locator"][0].locate_batch(y, want_logits=True)
    check_detect_locate(x, d, l, 4)


def test_config5_60s_streamed_against_oracle_full_length():
    """chunks of 20 s with a 5440-sample left halo, chunk starts multiples of the 320 hop."""
    from waveverify_b200 import embed_streaming
    m = models()
    T = 16000 * 60
    x, msg = synth(1, T, 24)
    y = embed_streaming(m["generator"][0], x.cuda(), msg[:1].cuda(), chunk_samples=320000)
    with torch.no_grad():
        wm_o = O.generator_forward(x, msg[:1], m["generator"][1], m["generator"][2])
    got = (y.cpu() - x).numpy()
    assert snr_db(wm_o.numpy(), got) >= 46.0
    # no drift along the clip: the error of the last 5 s is that of the first 5 s
    seg = 16000 * 5
    head = snr_db(wm_o.numpy()[..., :seg], got[..., :seg]); tail = snr_db(wm_o.numpy()[..., -seg:], got[..., -seg:])
    assert abs(head - tail) <= 3.0, (head, tail)


def test_error_does_not_drift_with_clip_length():
    """Same weights, clips of 1 s / 10 s / 60 s: the residual's SNR against the oracle stays within 3 dB."""
    m = models()
    snrs = {}
    for T in (16000, 160000, 960000):
        x, msg = synth(1, T, 31)
        wm, _, _ = m["generator"][0].embed_batch(x.cuda(), msg.cuda())
        with torch.no_grad():
            wm_o = O.generator_forward(x, msg, m["generator"][1], m["generator"][2])
        snrs[T] = snr_db(wm_o.numpy(), wm.cpu().numpy())
    print("wm SNR vs T:", snrs)
    assert min(snrs.values()) >= 46.0 and max(snrs.values()) - min(snrs.values()) <= 3.0, snrs
