"""Host logic of the request batcher (waveverify_b200/serving.py) with a stand-in backend: grouping, padding to
the bucket, per-request trimming, presence masks, error propagation, shutdown."""
import threading

import numpy as np
import pytest
import torch

from waveverify_b200.serving import RequestBatcher


class FakeBackend:
    device = torch.device("cpu")

    def __init__(self):
        self.calls = []

    def embed_batch(self, audio, msg):
        self.calls.append(("embed", tuple(audio.shape)))
        return audio + msg.sum(dim=1).view(-1, 1, 1)

    def detect_batch(self, audio, presence=None):
        self.calls.append(("detect", tuple(audio.shape)))
        n = presence.float().sum(dim=(1, 2))
        mean = (audio * presence.float()).sum(dim=(1, 2)) / n
        bits = (mean.view(-1, 1) > torch.linspace(-1, 1, 16).view(1, -1)).to(torch.uint8)
        return bits, n

    def locate_batch(self, audio):
        self.calls.append(("locate", tuple(audio.shape)))
        if audio.shape[0] == 3:
            raise RuntimeError("boom")
        return (audio[:, 0] > 0).to(torch.uint8)


def test_batches_pad_and_trim():
    be = FakeBackend()
    rng = np.random.RandomState(0)
    clips = [rng.standard_normal(n).astype(np.float32) for n in (320, 1600, 1601, 37, 1601, 640)]
    msgs = [np.full(16, i, np.float32) for i in range(6)]
    with RequestBatcher(be, max_batch=16, max_wait_s=0.2, hop=320) as rb:
        fe = [rb.embed(c, m) for c, m in zip(clips, msgs)]
        fd = [rb.detect(c) for c in clips]
        outs = [f.result(timeout=10) for f in fe]
        dets = [f.result(timeout=10) for f in fd]
    for c, m, o in zip(clips, msgs, outs):
        assert o.shape == c.shape and np.allclose(o, c + m.sum())
    for c, (bits, n) in zip(clips, dets):
        assert n == len(c)                                   # the presence mask covers exactly the clip's samples
        want = (c.mean() > np.linspace(-1, 1, 16)).astype(np.uint8)
        assert np.array_equal(bits, want)
    assert sum(rb.batches) == 12 and max(rb.batches) > 1    # requests were coalesced
    # hop-aligned clips (320, 1600, 640) share one padded batch; 1601 pairs with 1601; 37 runs alone
    assert sorted(shape for op, shape in be.calls if op == "embed") == [(1, 1, 37), (2, 1, 1601), (3, 1, 1600)]


def test_concurrent_clients_and_errors():
    be = FakeBackend()
    results = {}
    with RequestBatcher(be, max_batch=4, max_wait_s=0.05, hop=1) as rb:
        def client(i):
            x = np.full(50 + i, 1.0 if i % 2 else -1.0, np.float32)
            results[i] = rb.locate(x).result(timeout=10)
        th = [threading.Thread(target=client, args=(i,)) for i in range(8)]
        [t.start() for t in th]
        [t.join() for t in th]
        for i in range(8):
            assert results[i].shape == (50 + i,) and results[i].max() == (1 if i % 2 else 0)
        # a failing batch (the fake raises for batches of three) fails its own requests only
        f = [rb.locate(np.ones(10, np.float32)) for _ in range(3)]
        errs = 0
        for x in f:
            try:
                x.result(timeout=10)
            except RuntimeError:
                errs += 1
        assert errs in (0, 3)          # 3 when the three requests landed in one batch
        assert rb.locate(np.ones(10, np.float32)).result(timeout=10).sum() == 10
    with pytest.raises(RuntimeError):
        rb.embed(np.ones(4, np.float32), np.zeros(16, np.float32))
    with pytest.raises(ValueError):
        RequestBatcher(be, max_batch=0)
