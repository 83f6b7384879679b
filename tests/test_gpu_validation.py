"""GPU parity of the validation-path kernels through the C-ABI: bit-exact
against outputs of the reference itself (tests/golden/validation/validation.npz) and against the
numpy oracle on fresh seeded inputs, plus size-independent properties at full clip sizes."""
import os

import numpy as np
import pytest
import torch

import validation_oracle as VO
from waveverify_b200 import validation as V

from helpers import GOLDEN

pytestmark = pytest.mark.gpu
PATH = os.path.join(GOLDEN, "validation", "validation.npz")


@pytest.fixture(scope="module")
def gold():
    z = np.load(PATH)
    return {k: z[k] for k in z.files}


def inputs(seed, B, T):
    r = np.random.RandomState(seed)
    x = (0.1 * r.standard_normal((B, 1, T))).astype(np.float32)
    y = (x + 0.01 * r.standard_normal((B, 1, T))).astype(np.float32)
    return x, y


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def test_localization_matches_reference_golden(gold):
    sr = int(gold["sample_rate"])
    aug = V.LocalizationAugmentation(sr, 0.1)
    for i, (seed, B, T) in enumerate(gold["loc_cases"]):
        x, y = inputs(seed, B, T)
        np.random.seed(seed)
        sig, gt, upd, stats = aug(dev(x), dev(y))
        assert np.array_equal(sig.audio_data.cpu().numpy(), gold[f"loc{i}_wm"])
        assert np.array_equal(gt.cpu().numpy().astype(np.uint8), gold[f"loc{i}_gt"])
        assert np.array_equal(upd.cpu().numpy(), gold[f"loc{i}_upd"])
        assert np.allclose([stats[k] for k in ("original_revert", "zero_replace", "cross_substitute", "unchanged")],
                           gold[f"loc{i}_stats"], rtol=0, atol=1e-9)


def test_sequence_matches_reference_golden(gold):
    sr = int(gold["sample_rate"])
    aug = V.SequenceAugmentation(sr)
    for i, (seed, B, T) in enumerate(gold["seq_cases"]):
        x, y = inputs(seed, B, T)
        gt = (np.random.RandomState(seed + 1).rand(B, 1, T) < 0.7).astype(np.float32)
        np.random.seed(seed); torch.manual_seed(seed)
        sig, upd, gto, stats, method = aug(dev(x), dev(y), dev(gt))
        assert method == gold["seq_methods"][i]
        assert np.array_equal(sig.audio_data.cpu().numpy(), gold[f"seq{i}_wm"])
        assert np.array_equal(upd.cpu().numpy(), gold[f"seq{i}_upd"])
        assert np.array_equal(gto.cpu().numpy().astype(np.uint8), gold[f"seq{i}_gt"])
        assert abs(sum(stats.values()) - 100.0) < 1e-9


def test_effects_match_reference_golden(gold):
    x, _ = inputs(21, 3, 437)
    xd = dev(x)
    sr = int(gold["sample_rate"])
    out, _ = V.apply_effect(xd, "amplitude_scaling", sample_rate=sr, scale=0.5)
    assert np.array_equal(out.cpu().numpy(), gold["fx_scale"])
    for bd in (8, 16):
        out, _ = V.apply_effect(xd, "quantization", sample_rate=sr, bit_depth=bd)
        assert np.array_equal(out.cpu().numpy(), gold[f"fx_quant{bd}"])
    for k in (3, 5, 9):
        out, _ = V.apply_effect(xd, "median_filter", sample_rate=sr, kernel_size=k)
        assert np.array_equal(out.cpu().numpy(), gold[f"fx_median{k}"])
    out, _ = V.apply_effect(xd, "random_noise", sample_rate=sr, noise_std=0.01, noise=dev(gold["fx_noise_draw"]))
    assert np.array_equal(out.cpu().numpy(), gold["fx_noise"])
    mask = torch.ones_like(xd)
    out, m = V.apply_effect(xd, "sample_suppression", sample_rate=sr, mask=mask, suppression_percentage=0.1,
                            indices=gold["fx_suppress_idx"])
    assert np.array_equal(out.cpu().numpy(), gold["fx_suppress"])
    assert np.array_equal(m.cpu().numpy().astype(np.uint8), gold["fx_suppress_mask"])
    assert m is mask                                   # updated in place, like the reference
    out, m2 = V.apply_effect(xd, "identity", mask=mask)
    assert out is xd and m2 is mask


@pytest.mark.parametrize("B,T,sr", [(4, 16000, 16000), (3, 16001, 16000), (2, 33333, 16000), (5, 2, 16000), (2, 1599, 16000)])
def test_fused_augment_equals_sequential_oracle(B, T, sr):
    for seed in range(3):
        x, y = inputs(100 * B + seed, B, T)
        np.random.seed(seed + T); torch.manual_seed(seed + T)
        seg = int(sr * 0.1)
        wm1, gt1, upd1, _ = VO.localization_augment(x, y, seg)
        wm2, upd2, gt2, method = VO.sequence_augment(upd1, wm1, gt1, sr)
        np.random.seed(seed + T); torch.manual_seed(seed + T)
        loc = V.LocalizationAugmentation(sr, 0.1).plan(B, T)
        seq = V.SequenceAugmentation(sr).plan(T)
        wm, gt, og = V.augment(dev(x), dev(y), loc, seq)
        assert np.array_equal(wm.cpu().numpy(), wm2)
        assert np.array_equal(og.cpu().numpy(), upd2)
        assert np.array_equal(gt.cpu().numpy(), gt2)


def test_chunk_swap_and_error_codes():
    x, y = inputs(7, 2, 1000)
    plan = V.SequenceAugmentation.chunk_swap_plan(100, 600, 250)
    wm, og, gt = V._gather(dev(x), dev(y), None, None, plan)
    assert np.array_equal(wm.cpu().numpy(), VO.chunk_swap(y, 100, 600, 250))
    assert np.array_equal(og.cpu().numpy(), VO.chunk_swap(x, 100, 600, 250))
    assert float(gt.min()) == 1.0
    with pytest.raises(ValueError):   # overlapping chunks are rejected by the C-ABI (WV_ERR_INVALID)
        V._gather(dev(x), dev(y), None, None, V.SequenceAugmentation.chunk_swap_plan(100, 200, 250))
    with pytest.raises(ValueError):
        V.apply_effect(dev(x), "median_filter", kernel_size=33)
    with pytest.raises(ValueError):
        V.apply_effect(dev(x), "quantization", bit_depth=40)


def test_full_size_properties():
    """BASELINE config-2 size (64 x 1 s): properties that need no oracle run."""
    B, T = 64, 16000
    g = torch.Generator(device="cuda").manual_seed(3)
    x = 0.1 * torch.randn(B, 1, T, device="cuda", generator=g)
    y = x + 0.01 * torch.randn(B, 1, T, device="cuda", generator=g)
    # reverse twice / shift by a then T - a / shuffle then inverse permutation = identity
    rev = V.SequencePlan("reverse", V.SEQ_REVERSE)
    a, _, _ = V._gather(x, y, None, None, rev)
    b, _, _ = V._gather(x, a, None, None, rev)
    assert torch.equal(b, y)
    s1, s2 = V.SequencePlan("circular_shift", V.SEQ_SHIFT, a=1234), V.SequencePlan("circular_shift", V.SEQ_SHIFT, a=T - 1234)
    b, _, _ = V._gather(x, V._gather(x, y, None, None, s1)[0], None, None, s2)
    assert torch.equal(b, y)
    assert torch.equal(V._gather(x, y, None, None, s1)[0], torch.roll(y, 1234, 2))
    perm = np.random.RandomState(0).permutation(2).astype(np.int32)
    sh = V.SequencePlan("shuffle", V.SEQ_SHUFFLE, c=8000, perm=perm)
    inv = V.SequencePlan("shuffle", V.SEQ_SHUFFLE, c=8000, perm=np.argsort(perm).astype(np.int32))
    b, _, _ = V._gather(x, V._gather(x, y, None, None, sh)[0], None, None, inv)
    assert torch.equal(b, y)
    # localization: ground truth zero exactly where the output differs from the watermarked input
    np.random.seed(5)
    loc = V.LocalizationAugmentation(16000, 0.1).plan(B, T)
    wm, gt, og = V.augment(x, y, loc, None)
    assert torch.equal(wm[gt == 1], y[gt == 1])
    assert int((gt == 0).sum()) == int((loc.seg_op != 0).sum()) * 1600
    assert abs(loc.stats["unchanged"] - 100.0 * float(gt.mean())) < 1e-3
    # quantization is idempotent; the median of a constant interior is the constant; noise statistics
    q, _ = V.apply_effect(y, "quantization", bit_depth=8)
    q2, _ = V.apply_effect(q, "quantization", bit_depth=8)
    assert torch.equal(q, q2)
    assert torch.equal(q.cpu(), torch.round(y.cpu() * 127.0) / 127.0)   # the reference runs its effects on the CPU (true division)
    m, _ = V.apply_effect(torch.full_like(y, 0.25), "median_filter", kernel_size=5)
    assert torch.equal(m[..., 2:-2], torch.full_like(m[..., 2:-2], 0.25)) and float(m[0, 0, 0]) == 0.25
    m3, _ = V.apply_effect(y, "median_filter", kernel_size=3)
    pad = torch.nn.functional.pad(y, (1, 1))
    assert torch.equal(m3, pad.unfold(2, 3, 1).median(dim=-1).values)
    n1, _ = V.apply_effect(torch.zeros_like(y), "white_noise", noise_std=0.5, seed=11)
    n2, _ = V.apply_effect(torch.zeros_like(y), "white_noise", noise_std=0.5, seed=11)
    n3, _ = V.apply_effect(torch.zeros_like(y), "white_noise", noise_std=0.5, seed=12)
    assert torch.equal(n1, n2) and not torch.equal(n1, n3)
    z = (n1 / 0.5).double().flatten()
    assert abs(float(z.mean())) < 5e-3 and abs(float(z.std()) - 1.0) < 5e-3
    assert abs(float((z ** 4).mean()) - 3.0) < 0.05 and abs(float((z[:-1] * z[1:]).mean())) < 5e-3


def test_validation_pipeline_runs_and_counts():
    from helpers import fixture_weights
    from waveverify_b200 import Detector, Generator, Locator, AudioSignal
    from helpers import BASE_KW
    nets = {}
    for kind, cls in (("generator", Generator), ("detector", Detector), ("locator", Locator)):
        c, sd = fixture_weights(kind, False, 123)
        m = cls(**{**BASE_KW[kind], "bias": True, "zero_init": False})
        m.load_state_dict(sd)
        nets[kind] = m.to(torch.device("cuda:0"))
    B, T = 4, 16000
    g = torch.Generator(device="cuda").manual_seed(9)
    x = 0.1 * torch.randn(B, 1, T, device="cuda", generator=g)
    msg = torch.randint(0, 2, (B, 16), device="cuda", generator=g).float()
    effects = [("identity", {}), ("amplitude_scaling", {"scale": 0.5}), ("quantization", {"bit_depth": 8}),
               ("random_noise", {"noise_std": 0.001, "seed": 1}), ("median_filter", {"kernel_size": 3}),
               ("sample_suppression", {"suppression_percentage": 0.001})]
    pipe = V.ValidationPipeline(nets["generator"], nets["detector"], nets["locator"], effects=effects)
    np.random.seed(1); torch.manual_seed(1)
    wm, y, results, stats = pipe(AudioSignal(x, 16000), msg)
    assert set(results) == {e[0] for e in effects}
    # identity effect: the counters equal a by-hand evaluation of the same augmented batch
    np.random.seed(1); torch.manual_seed(1)
    loc = pipe.localization_augmenter.plan(B, T)
    seq = pipe.seq_augmenter.plan(T)
    y_aug, mask, _ = V.augment(x, y.audio_data, loc, seq)
    det = nets["detector"].detect_batch(y_aug, presence=mask)
    lm = nets["locator"].locate_batch(y_aug)["mask"]
    bits, valid = det["bits"].cpu().numpy(), det["valid"].cpu().numpy().astype(bool)
    mb = (msg.cpu().numpy() != 0)
    p, q = lm.cpu().numpy().astype(bool), mask.cpu().numpy().astype(bool)
    want = [int(((bits != mb) & valid).sum()), int(valid.sum()), int((p & q).sum()), int((p | q).sum()),
            int((~p & ~q).sum()), int((~p | ~q).sum())]
    assert results["identity"]["counters"].tolist() == want
    for r in results.values():
        assert 0.0 <= r["ber"] <= 1.0 and 0.0 <= r["miou"] <= 1.0
        assert r["mask"].shape == r["locator_mask"].shape


def test_fir_effects_within_2e_6_of_the_published_julius_algorithm():
    """julius low / high / band-pass (parity unpinned: julius is absent; oracle = its published algorithm in
    float64).  Tolerance: fp32 accumulation of <= 129 taps, |x| <= 0.5 -> 2e-6 absolute."""
    sr = 16000
    for B, T in ((3, 16000), (2, 4097), (1, 50), (2, 1)):
        x, _ = inputs(31 + T, B, T)
        xd = dev(x)
        for name, kw, ref in (
                ("lowpass_filter", dict(cutoff_freq=3000), lambda: VO.lowpass_filter(x, 3000, sr)),
                ("lowpass_filter", dict(cutoff_freq=1000), lambda: VO.lowpass_filter(x, 1000, sr)),
                ("highpass_filter", dict(cutoff_freq=500), lambda: VO.highpass_filter(x, 500, sr)),
                ("highpass_filter", dict(cutoff_freq=1000), lambda: VO.highpass_filter(x, 1000, sr)),
                ("bandpass_filter", dict(cutoff_freq_low=1000, cutoff_freq_high=3000), lambda: VO.bandpass_filter(x, 1000, 3000, sr))):
            out, m = V.apply_effect(xd, name, sample_rate=sr, mask=None, **kw)
            assert out.shape == xd.shape and m is None
            assert np.abs(out.cpu().numpy().astype(np.float64) - ref()).max() <= 2e-6, (name, kw, B, T)
    x, _ = inputs(5, 2, 8000)
    xd = dev(x)
    # the reference's quirks: a cutoff above half the Nyquist frequency makes julius raise -> low / high-pass return the
    # input unchanged, band-pass re-raises the ValueError
    out, _ = V.apply_effect(xd, "lowpass_filter", sample_rate=sr, cutoff_freq=5000)
    assert out is xd
    out, _ = V.apply_effect(xd, "highpass_filter", sample_rate=sr, cutoff_freq=0)
    assert out is xd
    with pytest.raises(ValueError):
        V.apply_effect(xd, "bandpass_filter", sample_rate=sr, cutoff_freq_low=1000, cutoff_freq_high=5000)
    with pytest.raises(ValueError):
        V.apply_effect(xd, "bandpass_filter", sample_rate=sr, cutoff_freq_low=3000, cutoff_freq_high=1000)
    # properties: DC gain 1 for the low-pass (taps sum to 1), low + high = identity
    c = torch.full((1, 1, 4000), 0.25, device="cuda")
    lo, _ = V.apply_effect(c, "lowpass_filter", sample_rate=sr, cutoff_freq=1000)
    assert float((lo - 0.25).abs().max()) < 1e-6
    lo, _ = V.apply_effect(xd, "lowpass_filter", sample_rate=sr, cutoff_freq=1000)
    hi, _ = V.apply_effect(xd, "highpass_filter", sample_rate=sr, cutoff_freq=1000)
    assert float((lo + hi - xd).abs().max()) < 1e-6


def test_resample_and_speed_effects():
    """`resample` against outputs of the reference itself (tests/golden/validation/resample.npz: the reference's
    apply_effect -> torchaudio.transforms.Resample down and up) and `speed` against torchaudio's resampler + the reference's
    linear stretch (SoX itself: parity unpinned); then the float64 oracle at full clip length.  Tolerance 2e-6 absolute
    (fp32 FIR of <= 30 taps, |x| <= 0.5)."""
    z = np.load(os.path.join(GOLDEN, "validation", "resample.npz"))
    for i, (seed, B, T, sr, new_sr) in enumerate(z["resample_cases"]):
        x, _ = inputs(int(seed), int(B), int(T))
        m = torch.ones(int(B), 1, int(T), device="cuda")
        out, m2 = V.apply_effect(dev(x), "resample", sample_rate=int(sr), mask=m, new_sample_rate=int(new_sr))
        assert m2 is m and out.shape == z[f"resample{i}"].shape
        assert np.abs(out.cpu().numpy() - z[f"resample{i}"]).max() <= (2e-6 if int(new_sr) in (32000, 8000) else 1e-5), i
    for i, (seed, B, T, sr, sp) in enumerate(z["speed_cases"]):
        x, _ = inputs(int(seed), int(B), int(T))
        out, _ = V.apply_effect(dev(x), "speed", sample_rate=int(sr), speed=float(sp) / 1000.0)
        assert out.shape == x.shape
        assert np.abs(out.cpu().numpy() - z[f"speed{i}"]).max() <= 5e-6, i
    x, _ = inputs(77, 4, 16000)
    xd = dev(x)
    out, _ = V.apply_effect(xd, "resample", sample_rate=16000, new_sample_rate=32000)       # conf/effects_config.yml:70-91
    assert np.abs(out.cpu().numpy() - VO.resample_effect(x, 32000, 16000)).max() <= 2e-6
    out, _ = V.apply_effect(xd, "speed", sample_rate=16000, speed=0.8)
    assert np.abs(out.cpu().numpy() - VO.speed_effect(x, 0.8, 16000)).max() <= 2e-6
    # error behaviour of the reference: bad rate -> ValueError; non-positive speed -> logged, input returned unchanged
    with pytest.raises(ValueError):
        V.apply_effect(xd, "resample", sample_rate=16000, new_sample_rate=0)
    with pytest.raises(ValueError):
        V.apply_effect(xd, "resample", sample_rate=16000, new_sample_rate=8000.0)
    out, _ = V.apply_effect(xd, "speed", sample_rate=16000, speed=-1.0)
    assert out is xd
    out, _ = V.apply_effect(xd, "speed", sample_rate=16000, speed=1.0)
    assert out is xd


def test_validation_pipeline_runs_the_references_eval_effects():
    """conf/effects_config.yml:70-91 `eval_effects` end to end on the device (model/watermarking.py:443-483)."""
    from helpers import BASE_KW, fixture_weights
    from waveverify_b200 import AudioSignal, Detector, Generator, Locator
    mods = {}
    for kind, cls in (("generator", Generator), ("detector", Detector), ("locator", Locator)):
        _, sd = fixture_weights(kind, False, 0)
        m = cls(**{**BASE_KW[kind], "bias": True, "zero_init": False})
        m.load_state_dict(sd)
        mods[kind] = m.cuda()
    eval_effects = [("identity", {}), ("resample", {"new_sample_rate": 32000}), ("speed", {"speed": 0.8}),
                    ("random_noise", {"noise_std": 0.001}), ("lowpass_filter", {"cutoff_freq": 2000}),
                    ("highpass_filter", {"cutoff_freq": 3500}), ("bandpass_filter", {"cutoff_freq_low": 300, "cutoff_freq_high": 4000})]
    pipe = V.ValidationPipeline(mods["generator"], mods["detector"], mods["locator"], effects=eval_effects)
    rng = np.random.RandomState(3)
    x = torch.from_numpy((0.1 * rng.standard_normal((4, 1, 16000))).astype(np.float32)).cuda()
    msg = torch.from_numpy(rng.randint(0, 2, (4, 16))).cuda()
    np.random.seed(1); torch.manual_seed(1)
    wm, y, results, stats = pipe(AudioSignal(x, 16000), msg)
    assert list(results) == [n for n, _ in eval_effects]
    for name, r in results.items():
        assert 0.0 <= r["ber"] <= 1.0 and 0.0 <= r["miou"] <= 1.0, name
        assert r["locator_mask"].shape == r["mask"].shape
