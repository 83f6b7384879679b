"""GPU: the device-side bit re-evaluation (wv_detector_refine: selection kernel + CUDA-graph WHILE node around
{gather, precise net, scatter}) against whole-batch runs of the two nets it combines."""
import numpy as np
import pytest
import torch

from helpers import BASE_KW, fixture_weights

pytestmark = pytest.mark.gpu


def detector():
    from waveverify_b200 import Detector
    c, sd = fixture_weights("detector", False, 0)
    m = Detector(**{**BASE_KW["detector"], "bias": True, "zero_init": False})
    m.load_state_dict(sd)
    return m.cuda()


def clips(B, T, seed):
    rng = np.random.RandomState(seed)
    return torch.from_numpy((0.1 * rng.standard_normal((B, 1, T))).astype(np.float32)).cuda()


@pytest.mark.parametrize("B,T,slots", [(5, 8000, 2), (3, 16000, 4), (9, 4097, 1)])
def test_every_clip_selected_equals_the_precise_net(B, T, slots):
    """tau = 1: every clip is 'near' -> ceil(B / slots) passes of the WHILE body; all outputs equal the precise net's."""
    D = detector()
    y = clips(B, T, 3)
    ref = D.detect_batch(y, want_logits=True, precise=True)
    D.EXACT_TAU = D.EXACT_TAU_SHORT = 1.0
    D.refine_slots = slots
    n0, p0 = D.recheck_count, D.recheck_passes
    out = D.detect_batch(y, want_logits=True)
    for k in ("bits", "avg", "conf", "valid", "logits"):
        assert torch.equal(out[k], ref[k]), k
    assert D.recheck_count - n0 == B and D.recheck_passes - p0 == -(-B // min(slots, B))


def test_no_clip_selected_leaves_the_fast_path_untouched():
    D = detector()
    y = clips(4, 8000, 4)
    D.exact_bits = False
    fast = D.detect_batch(y, want_logits=True)
    D.exact_bits = True
    D.EXACT_TAU = D.EXACT_TAU_SHORT = 0.0
    n0 = D.recheck_count
    out = D.detect_batch(y, want_logits=True)
    for k in ("bits", "avg", "conf", "valid", "logits"):
        assert torch.equal(out[k], fast[k]), k
    assert D.recheck_count == n0


def test_partial_selection_and_masked_decode():
    """A band that selects some clips: the selected ones carry the precise net's values, the others the fast path's;
    with a presence mask the band is applied to valid bits only (scripts/evaluate.py:471-494)."""
    D = detector()
    B, T = 8, 8000
    y = clips(B, T, 5)
    pres = torch.ones(B, 1, T, dtype=torch.uint8, device="cuda")
    pres[1] = 0                      # no valid bit: never selected
    pres[2, :, : T // 2] = 0
    for pm in (None, pres):
        D.exact_bits = False
        fast = D.detect_batch(y, presence=pm)
        prec = D.detect_batch(y, presence=pm, precise=True)
        D.exact_bits = True
        margin = (fast["avg"] - 0.5).abs()
        if pm is not None:
            margin = torch.where(fast["valid"] != 0, margin, torch.full_like(margin, 9.0))
        m = margin.min(dim=1).values
        tau = float(m.sort().values[B // 2 - 1]) * 1.0001 + 1e-9     # selects about half of the clips
        D.EXACT_TAU = D.EXACT_TAU_SHORT = tau
        D.refine_slots = 3
        out = D.detect_batch(y, presence=pm)
        sel = m < tau
        assert 0 < int(sel.sum()) < B
        for k in ("bits", "avg", "conf", "valid"):
            want = torch.where(sel.reshape(-1, *([1] * (fast[k].dim() - 1))), prec[k], fast[k])
            assert torch.equal(out[k], want), (k, pm is not None)


def test_refine_survives_workspace_growth_and_chunking():
    """The refine graphs hold pointers into the precise net's workspace: a later, larger precise call moves the workspace
    (plans and graphs are rebuilt), and a sub-batch limit clamps the number of slots."""
    D = detector()
    y = clips(6, 8000, 7)
    ref = D.detect_batch(y, precise=True)
    D.EXACT_TAU = D.EXACT_TAU_SHORT = 1.0
    D.refine_slots = 2
    a = D.detect_batch(y)
    big = clips(48, 16000, 8)
    D.detect_batch(big, precise=True)                  # grows the precise net's workspace
    b = D.detect_batch(y)
    D.set_chunk_samples(8000)                          # one clip per sub-batch / slot
    c = D.detect_batch(y)
    D.set_chunk_samples(0)
    for out in (a, b, c):
        for k in ("bits", "avg", "conf", "valid"):
            assert torch.equal(out[k], ref[k]), k
