"""Generate tests/golden/*.npz by running the UNMODIFIED reference modules (imported from
/root/reference, build container only) on deterministic fixture weights and inputs.

TEST INFRASTRUCTURE.  Run:  python oracle/make_golden.py
The fixtures pin BOTH the oracle (tests/test_oracle_golden.py) and the CUDA path
(tests/test_gpu_parity.py).  Weights are not stored: they are regenerated from
`waveverify_b200.params.fixture_state_dict(cfg, seed)` (numpy MT19937, machine independent) and
the script asserts that the reference accepts them with load_state_dict(strict=True), i.e. that
key names and shapes are the reference's.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))

import ref_import  # noqa: E402
from waveverify_b200 import params as P  # noqa: E402

CASES = [
    # name, zero_init, B, T, weight seed, input seed
    ("a_zi0_B2_T16000", False, 2, 16000, 0, 7),
    ("b_zi1_B3_T4097", True, 3, 4097, 1, 8),
    ("c_zi0_B1_T16001", False, 1, 16001, 0, 9),
    ("d_zi0_B2_T100", False, 2, 100, 0, 10),
    ("e_zi1_B1_T320", True, 1, 320, 1, 11),
    ("f_zi0_B1_T50000", False, 1, 50000, 2, 12),
]
DET_DECIM = 5


def make_inputs(B, T, seed):
    rng = np.random.RandomState(seed)
    # speech-like level with a slow envelope so STFT bins span some dynamic range
    x = 0.1 * rng.standard_normal((B, 1, T)).astype(np.float32)
    env = 0.35 + 0.65 * np.abs(np.sin(np.arange(T, dtype=np.float32) * (2 * np.pi / 6000.0)))
    x = x * env[None, None, :]
    msg = rng.randint(0, 2, size=(B, 16)).astype(np.int64)
    return x.astype(np.float32), msg


def main():
    torch.manual_seed(0)
    torch.set_num_threads(8)
    G, D, L, AS, cfg = ref_import.load()
    out_dir = os.path.join(os.path.dirname(HERE), "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    kws = dict(generator=cfg["Generator"], detector=cfg["Detector"],
               locator={k: v for k, v in cfg["Locator"].items() if k != "nbits"})
    for name, zi, B, T, wseed, iseed in CASES:
        g, d, l, _ = ref_import.build(zi)
        mods = dict(generator=g, detector=d, locator=l)
        for kind, m in mods.items():
            c = P.config_from_kwargs(kind, {**kws[kind], "bias": True, "zero_init": zi})
            m.load_state_dict(P.fixture_state_dict(c, wseed), strict=True)
        x_np, msg_np = make_inputs(B, T, iseed)
        x = torch.from_numpy(x_np); msg = torch.from_numpy(msg_np)
        with torch.no_grad():
            wm = g(AS(x, 16000), msg).audio_data                 # model/generator.py:360
            y = AS(x, 16000) + wm                                 # model/watermarking.py:440
            y = y.audio_data
            det = d(AS(y, 16000))                                 # model/detector.py:366
            loc = l(AS(y, 16000))                                 # model/locator.py:268
            latent = g.encode(x, msg)
            p = torch.sigmoid(det)                                # waveverify/core.py:577-583
            avg = p.mean(dim=2)
            bits = (avg >= 0.5).to(torch.uint8)                   # waveverify/utils.py:401
            conf = avg.mean(dim=1)
            mask = (loc > 0.5).to(torch.uint8)                    # model/watermarking.py:717
            post = d.postprocess(det)                             # model/detector.py:320
        np.savez_compressed(
            os.path.join(out_dir, name + ".npz"),
            zero_init=np.array(zi), wseed=np.array(wseed), x=x_np, msg=msg_np,
            wm=wm.numpy(), y=y.numpy(), latent=latent.numpy(),
            det_logits_decim=det[:, :, ::DET_DECIM].numpy(), det_decim=np.array(DET_DECIM),
            det_avg=avg.numpy(), det_bits=bits.numpy(), det_conf=conf.numpy(),
            det_post=post.numpy(), loc_logits=loc.numpy(), loc_mask=mask.numpy())
        print(name, "wm rms %.4f" % float(wm.pow(2).mean().sqrt()),
              "mask frac %.3f" % float(mask.float().mean()), "bits", bits[0].tolist())


if __name__ == "__main__":
    main()
