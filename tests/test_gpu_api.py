"""GPU: public API (waveverify/core.py surface), checkpoints, exact streaming (BASELINE config 5),
sub-batched detector+locator (config 4 shape at reduced batch)."""
import numpy as np
import pytest
import torch

import wv_oracle as O
from helpers import BASE_KW, fixture_weights, oracle_cfg, snr_db

pytestmark = pytest.mark.gpu


def _fixture_models(zero_init=False, seed=0):
    from waveverify_b200 import Detector, Generator, Locator
    out = {}
    for kind, cls in (("generator", Generator), ("detector", Detector), ("locator", Locator)):
        c, sd = fixture_weights(kind, zero_init, seed)
        m = cls(**{**BASE_KW[kind], "bias": True, "zero_init": zero_init})
        m.load_state_dict(sd)
        out[kind] = (m.cuda(), sd, c)
    return out


def test_waveverify_file_api_and_atomic_checkpoint(tmp_path):
    from waveverify_b200 import WatermarkID, WaveVerify, load_audio, save_audio
    mods = _fixture_models()
    # "atomic" checkpoint with parametrizations removed (scripts/train.py:1589-1676)
    ck = {"step": 1, "models": {k: O.fold_state_dict(v[1]) for k, v in mods.items()}}
    ckp = tmp_path / "ck.pth"
    torch.save(ck, ckp)
    kw = dict(bias=True, zero_init=False)
    wv = WaveVerify(checkpoint=ckp, device="cuda", generator_kwargs=kw, detector_kwargs=kw, locator_kwargs=kw)
    rng = np.random.RandomState(0)
    x = torch.from_numpy((0.1 * rng.standard_normal(24000)).astype(np.float32))
    src = tmp_path / "in.wav"; dst = tmp_path / "out" / "wm.wav"
    save_audio(x, src, 16000)
    wid = WatermarkID.for_creator("beyonce_2024")
    y, sr, wid2 = wv.embed(src, wid, dst)
    assert sr == 16000 and wid2 == wid and y.shape == (24000,) and dst.exists()
    xq, _ = load_audio(src)                                 # 16-bit quantised input actually embedded
    wm_direct, y_direct, _ = mods["generator"][0].embed_batch(xq[None].cuda(), torch.tensor([[int(b) for b in wid.bits]]).cuda())
    # checkpoint path == direct weights, up to fp16 flips of weights whose re-folded fp32 value moved by 1 ulp
    assert snr_db(y_direct[0, 0].cpu().numpy(), y) > 55.0
    det, conf = wv.detect(dst)
    assert isinstance(det, WatermarkID) and 0.0 <= conf <= 1.0
    assert wv.verify(dst, det) is True
    assert wv.verify(dst, WatermarkID.custom(det.to_int() ^ 1)) is False
    probs = wv.locate(dst)
    assert probs.shape == (24000,) and probs.min() >= 0.0 and probs.max() <= 1.0
    with pytest.raises(RuntimeError, match="Failed to embed"):
        wv.embed(tmp_path / "nope.wav", wid)
    with pytest.raises(RuntimeError):
        wv.embed(src, "not-bits")
    # batched tensor API
    xb = (0.1 * torch.randn(3, 1, 8000)).cuda()
    yb = wv.embed_batch(xb, torch.randint(0, 2, (3, 16)).cuda())
    bits, cf = wv.detect_batch(yb)
    mask = wv.locate_batch(yb)
    assert yb.shape == xb.shape and bits.shape == (3, 16) and cf.shape == (3,) and mask.shape == (3, 8000)


def test_streaming_embed_equals_whole_clip_and_oracle_prefix():
    """Long-form clip through the Generator in chunks with a 5440-sample causal halo: identical to
    one whole-clip pass, and equal to the CPU oracle on a prefix within the fp16 tolerance."""
    from waveverify_b200 import embed_streaming
    mods = _fixture_models()
    G, sd, c = mods["generator"]
    rng = np.random.RandomState(4)
    T = 16000 * 40 + 123                                     # 40 s, ragged tail
    x = torch.from_numpy((0.1 * rng.standard_normal((1, 1, T))).astype(np.float32)).cuda()
    msg = torch.from_numpy(rng.randint(0, 2, (1, 16))).cuda()
    _, whole, _ = G.embed_batch(x, msg, want_wm=False)
    for chunk in (320 * 100, 320 * 500):
        y = embed_streaming(G, x, msg, chunk_samples=chunk)
        assert float((y - whole).abs().max()) <= 1e-6, f"chunk {chunk}"
    Tp = 16000 * 4
    with torch.no_grad():
        wm_o = O.generator_forward(x[:, :, :Tp].cpu(), msg.cpu(), O.fold_state_dict(sd), oracle_cfg(c))
    got = (y[:, :, :Tp] - x[:, :, :Tp]).cpu().numpy()
    assert snr_db(wm_o.numpy(), got) >= 46.0
    with pytest.raises(ValueError):
        embed_streaming(G, x, msg, chunk_samples=1000)


def test_subbatched_detector_locator_config4_shape():
    """BASELINE config 4 shape (5 s clips, Detector+Locator only) at a reduced batch: internal
    sub-batching must be invisible in bits / masks / averages."""
    mods = _fixture_models()
    D, L = mods["detector"][0], mods["locator"][0]
    g = torch.Generator().manual_seed(9)
    B, T = 24, 80000
    y = (0.1 * torch.randn(B, 1, T, generator=g)).cuda()
    d0 = D.detect_batch(y); l0 = L.locate_batch(y)
    D.set_chunk_samples(5 * T); L.set_chunk_samples(7 * T)   # 5 + 5 + 5 + 5 + 4 clips, 7 + 7 + 7 + 3
    d1 = D.detect_batch(y); l1 = L.locate_batch(y)
    D.set_chunk_samples(0); L.set_chunk_samples(0)
    assert torch.equal(d0["bits"], d1["bits"]) and torch.equal(d0["avg"], d1["avg"]) and torch.equal(d0["conf"], d1["conf"])
    assert torch.equal(l0["mask"], l1["mask"])
    assert D.launches(B, T) > 0 and D.workspace_bytes() > 0


def test_metric_counters_allreduce_single_process():
    from waveverify_b200 import metric_counters
    from waveverify_b200.dist import allreduce_counters
    bits = torch.randint(0, 2, (4, 16), dtype=torch.uint8).cuda()
    c = metric_counters(bits, None, bits.clone(), torch.ones(4, 1, 100, dtype=torch.uint8).cuda(),
                        torch.ones(4, 1, 100, dtype=torch.uint8).cuda())
    assert allreduce_counters(c).tolist() == [0, 64, 400, 400, 0, 0]


def test_request_batcher_equals_single_requests():
    """serving.RequestBatcher: clips of different lengths coalesced into one padded batch give the results of
    one call per clip (embed: bit-exact - causal nets, independent clips; detect: bits equal, confidence to 1e-6;
    locate: mask exact)."""
    from waveverify_b200 import WaveVerify
    from waveverify_b200.serving import RequestBatcher
    kw = dict(bias=True, zero_init=False)
    wv = WaveVerify(checkpoint=None, device="cuda", generator_kwargs=kw, detector_kwargs=kw, locator_kwargs=kw)
    for kind, m in (("generator", wv.model.generator), ("detector", wv.model.detector), ("locator", wv.model.locator)):
        _, sd = fixture_weights(kind, False, 3)
        m.load_state_dict(sd)
    rng = np.random.RandomState(8)
    lens = [16000, 12345, 1600, 777, 16000, 3840, 12345, 32000]
    clips = [(0.1 * rng.standard_normal(n)).astype(np.float32) for n in lens]
    msgs = [rng.randint(0, 2, 16).astype(np.float32) for _ in lens]
    with RequestBatcher(wv, max_batch=8, max_wait_s=0.5, hop=320) as rb:
        fe = [rb.embed(c, m) for c, m in zip(clips, msgs)]
        ys = [f.result(timeout=120) for f in fe]
        fd = [rb.detect(y) for y in ys]
        fl = [rb.locate(y) for y in ys]
        dets = [f.result(timeout=120) for f in fd]
        locs = [f.result(timeout=120) for f in fl]
    assert max(rb.batches) >= 5                       # the five hop-aligned clips ran as one padded batch
    dev = torch.device("cuda")
    for c, m, y, (bits, conf), mask in zip(clips, msgs, ys, dets, locs):
        x1 = torch.from_numpy(c).view(1, 1, -1).to(dev)
        y1 = wv.embed_batch(x1, torch.from_numpy(m).view(1, -1).to(dev))
        assert np.array_equal(y1.cpu().numpy()[0, 0], y)
        b1, c1 = wv.detect_batch(y1)
        assert np.array_equal(b1.cpu().numpy()[0], bits) and abs(float(c1[0]) - conf) < 1e-6
        assert np.array_equal(wv.locate_batch(y1).cpu().numpy()[0], mask)
