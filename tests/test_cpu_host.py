"""CPU: host-side logic - parameter inventory, folding, config validation, C-ABI symbols,
sharding + counter reduction over gloo (world_size 2)."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

import wv_oracle as O
from helpers import BASE_KW, ROOT, fixture_weights, net_config
from waveverify_b200 import Detector, Generator, Locator, _lib, params as P
from waveverify_b200.dist import shard_range
from waveverify_b200.fold import fold_state_dict


def test_library_loads_and_exports_every_declared_symbol():
    from waveverify_b200 import build
    build.build()
    header = open(os.path.join(ROOT, "include", "wv_b200.h")).read()
    declared = set(re.findall(r"\b(wv_[a-z0-9_]+)\s*\(", header))
    declared -= {"wv_net", "wv_tensor", "wv_net_config"}
    assert len(declared) >= 18
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in include/wv_b200.h but not exported"
    assert set(_lib.EXPORTED_SYMBOLS) == declared
    assert _lib.lib().wv_version() >= 1


@pytest.mark.parametrize("kind,cls", [("generator", Generator), ("detector", Detector), ("locator", Locator)])
@pytest.mark.parametrize("zero_init", [False, True])
def test_state_dict_keys_and_fold_match_oracle(kind, cls, zero_init):
    m = cls(**{**BASE_KW[kind], "bias": True, "zero_init": zero_init})
    c, sd = fixture_weights(kind, zero_init, 3)
    assert list(m.state_dict().keys()) == list(sd.keys())
    m.load_state_dict(sd)
    a = fold_state_dict(m.state_dict())
    b = O.fold_state_dict(sd)
    assert a.keys() == b.keys()
    for k in a:
        np.testing.assert_allclose(a[k].numpy(), b[k].numpy(), rtol=1e-6, atol=1e-7, err_msg=k)


def test_expected_key_counts():
    # 341 keys in G; D 197; L 89 (zero_init=False)
    assert len(P.param_spec(net_config("generator", False))) == 341
    assert len(P.param_spec(net_config("detector", False))) == 197
    assert len(P.param_spec(net_config("locator", False))) == 89


def test_plain_weight_checkpoint_loads():
    """Checkpoints saved with parametrizations removed (scripts/train.py:1624-1629)."""
    m = Locator(**{**BASE_KW["locator"], "bias": True, "zero_init": False})
    c, sd = fixture_weights("locator", False, 5)
    plain = O.fold_state_dict(sd)
    m.load_state_dict(plain)
    again = fold_state_dict(m.state_dict())
    for k in plain:
        np.testing.assert_allclose(again[k].numpy(), plain[k].numpy(), rtol=2e-6, atol=1e-7, err_msg=k)


def test_weight_standardization_fold():
    g = torch.Generator().manual_seed(0)
    v = torch.randn(6, 4, 5, generator=g); gg = torch.rand(6, 1, 1, generator=g) + 0.5
    sc = torch.tensor([1.3])
    sd = {"c.weight_v": v, "c.weight_g": gg, "c.weight_scale": sc}
    a = fold_state_dict(sd)["c.weight"]
    flat = v.flatten(1)
    ref = gg * sc * (v - flat.mean(1).view(-1, 1, 1)) / torch.sqrt(torch.clamp(flat.var(1, unbiased=False).view(-1, 1, 1) * 20, min=1e-7))
    np.testing.assert_allclose(a.numpy(), ref.numpy(), rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(a.numpy(), O.fold_state_dict(sd)["c.weight"].numpy(), rtol=1e-6, atol=1e-7)


def test_unsupported_options_rejected():
    with pytest.raises(NotImplementedError):
        Generator(causal=False)
    with pytest.raises(NotImplementedError):
        Detector(norm="layer_norm")
    with pytest.raises(NotImplementedError):
        Locator(skip="1x1")
    with pytest.raises(ValueError):
        Generator(sample_rate=0)
    with pytest.raises(ValueError):
        Detector(nbits=0)
    with pytest.raises(TypeError):
        Locator(nbits=16)        # the reference Locator has no nbits kwarg


def test_cpu_model_refuses_to_run():
    m = Locator(**{**BASE_KW["locator"]})
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m.locate_batch(torch.zeros(1, 1, 320))


def test_reference_attributes():
    g = Generator()
    assert (g.nbits, g.ratios, g.dimension, g.sample_rate, g.hop_length) == (16, [8, 5, 4, 2], 128, 16000, 320)
    d = Detector()
    assert (d.nbits, d.hop_length, d.output_dim) == (16, 320, 32)
    l = Locator(**BASE_KW["locator"])
    assert (l.hop_length, l.dimension) == (32, 64)
    n, padded = d.preprocess(torch.zeros(2, 1, 1000))
    assert n == 1000 and padded.shape[-1] == 1280
    assert g.preprocess(torch.zeros(1, 1, 321)).shape[-1] == 640
    with pytest.raises(ValueError):
        d.preprocess(torch.zeros(2, 1000))
    assert d.postprocess(torch.randn(2, 16, 50)).shape == (2, 16)


def test_shard_range_partitions_batch():
    for n in (1, 7, 64, 512):
        for w in (1, 2, 3, 8):
            spans = [shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def _gloo_worker(rank, world, port, q):
    import torch.distributed as dist
    from waveverify_b200.dist import allreduce_counters, shard_range
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    g = torch.Generator().manual_seed(5)
    B, T = 9, 700
    bits = torch.randint(0, 2, (B, 16), generator=g, dtype=torch.uint8)
    msg = torch.randint(0, 2, (B, 16), generator=g, dtype=torch.uint8)
    valid = torch.rand(B, 16, generator=g) > 0.2
    pred = torch.randint(0, 2, (B, 1, T), generator=g, dtype=torch.uint8)
    gt = torch.randint(0, 2, (B, 1, T), generator=g, dtype=torch.uint8)
    a, b = shard_range(B, rank, world)
    c = torch.tensor(O.metric_counters(bits[a:b], valid[a:b], msg[a:b], pred[a:b], gt[a:b]), dtype=torch.int64)
    allreduce_counters(c)
    full = O.metric_counters(bits, valid, msg, pred, gt)
    q.put((rank, c.tolist(), full))
    dist.destroy_process_group()


def test_counter_allreduce_matches_global_metrics_gloo_world2():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, got, full in res:
        assert got == full, (rank, got, full)
