"""Where does the precise Locator differ from the fp64 oracle?  Runs the 64 x 1 s batch of
tests/test_gpu_baseline_sizes.py (seed 21 audio used directly as the locator input), finds the clip with the worst
logit, and compares every tapped launch output of that clip with the fp64 oracle (and the fp32 oracle's own error)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import wv_oracle as O  # noqa: E402
from helpers import BASE_KW, fixture_weights, oracle_cfg  # noqa: E402
from waveverify_b200 import Locator  # noqa: E402

torch.set_num_threads(os.cpu_count() or 8)
dev = torch.device("cuda:0")
c, sd = fixture_weights("locator", False, 0)
L = Locator(**{**BASE_KW["locator"], "bias": True, "zero_init": False}); L.load_state_dict(sd); L = L.to(dev)
W64 = O.fold_state_dict(sd, dtype=torch.float64); W32 = O.fold_state_dict(sd)
rng = np.random.RandomState(int(os.environ.get("SEED", 21)))
x = torch.from_numpy((0.1 * rng.standard_normal((64, 1, 16000))).astype(np.float32))
ll = L.locate_batch(x.to(dev), want_logits=True)["logits"].cpu().numpy()
with torch.no_grad():
    l64 = O.locator_forward(x.double(), W64, oracle_cfg(c)).numpy()
    l32 = O.locator_forward(x, W32, oracle_cfg(c)).numpy()
e = np.abs(ll - l64); e32 = np.abs(l32 - l64)
print("precise vs fp64: max %.2e p99.99 %.2e p99.9 %.2e median %.2e | fp32 oracle vs fp64: max %.2e p99.99 %.2e median %.2e" %
      (e.max(), np.quantile(e, 0.9999), np.quantile(e, 0.999), np.median(e), e32.max(), np.quantile(e32, 0.9999), np.median(e32)))
b, _, t = np.unravel_index(e.argmax(), e.shape)
print("worst sample: clip %d t %d (frame %d): ours %.6f fp64 %.6f fp32 %.6f" % (b, t, t // 32, ll[b, 0, t], l64[b, 0, t], l32[b, 0, t]))
xb = x[b:b + 1]
taps64, taps32 = {}, {}
with torch.no_grad():
    O.locator_forward(xb.double(), W64, oracle_cfg(c), taps64)
    O.locator_forward(xb, W32, oracle_cfg(c), taps32)


def split_to_f32(buf, C):   # [rows, 2C] fp16 -> hi + lo
    return buf[:, :C].float() + buf[:, C:].float()


def report(name, got, key):
    ref = taps64[key][0].numpy().T          # [T, C]
    r32 = taps32[key][0].numpy().T
    g = got.cpu().numpy().astype(np.float64)[:, :ref.shape[1]]
    d = np.abs(g - ref); d32 = np.abs(r32 - ref)
    i = np.unravel_index(d.argmax(), d.shape)
    print("%-22s ours max %.2e (at row %d ch %d, ref %.5f) rms %.2e | fp32 max %.2e rms %.2e" %
          (name, d.max(), i[0], i[1], ref[i], np.sqrt((d ** 2).mean()), d32.max(), np.sqrt((d32 ** 2).mean())))


T = 16000
Ts = [16000, 4000, 500]
Cs = [32, 64, 128]
for s in range(2):
    C, Tn = Cs[s], Ts[s]
    raw = L.debug_tap(xb.to(dev), None, f"enc.s{s}.r0.out", 0, (Tn, C), torch.float32)
    report(f"s{s} resblock out (raw)", raw, f"enc_s{s}_res")
    ldy = (64 << s) // 2 + 8
    Y = L.debug_tap(xb.to(dev), None, f"enc.s{s}.spec.stft", 0, (Tn, 2 * ldy), torch.float16)
    report(f"s{s} log-spectrogram", split_to_f32(Y, ldy), f"enc_s{s}_y")
    dn = L.debug_tap(xb.to(dev), None, f"enc.s{s}.down", 0, (Ts[s + 1], 2 * C), torch.float32)
    report(f"s{s} down (raw)", dn, f"enc_s{s}")
ldy = 256 // 2 + 8
Y = L.debug_tap(xb.to(dev), None, "enc.post.spec.stft", 0, (500, 2 * ldy), torch.float16)
report("post log-spectrogram", split_to_f32(Y, ldy), "enc_post_y")
lat = L.debug_tap(xb.to(dev), None, "enc.latent", 0, (500, 128), torch.float16)
report("latent", split_to_f32(lat, 64), "enc_latent")
