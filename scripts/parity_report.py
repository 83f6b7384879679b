"""Parity of the CUDA path against the committed outputs of the reference (tests/golden): SNR and
max-abs per output, decoded-bit mismatches, and how close to the threshold every differing locator
mask sample sits.  Prints a markdown table (kept under profiles/)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
from helpers import BASE_KW, fixture_weights, golden_cases, load_case, snr_db  # noqa: E402
from waveverify_b200 import Detector, Generator, Locator  # noqa: E402

dev = torch.device("cuda:0")
cache = {}


def models(zi, seed):
    key = (bool(zi), int(seed))
    if key not in cache:
        out = {}
        for kind, cls in (("generator", Generator), ("detector", Detector), ("locator", Locator)):
            c, sd = fixture_weights(kind, key[0], key[1])
            m = cls(**{**BASE_KW[kind], "bias": True, "zero_init": key[0]})
            m.load_state_dict(sd)
            out[kind] = m.to(dev)
        cache[key] = out
    return cache[key]


print("| case | wm SNR dB | wm max-abs | latent SNR | det logits SNR | det max-abs | avg max-abs | bit mismatches (min margin of ref) | loc logits SNR | loc max-abs | mask mismatches / samples | max margin of a mismatch | samples inside that margin |")
print("|---|---|---|---|---|---|---|---|---|---|---|---|---|")
for path in golden_cases():
    z = load_case(path)
    m = models(z["zero_init"], z["wseed"])
    x = torch.from_numpy(z["x"]).to(dev); msg = torch.from_numpy(z["msg"]).to(dev)
    wm, y, lat = m["generator"].embed_batch(x, msg, want_latent=True)
    yg = torch.from_numpy(z["y"]).to(dev)
    d = m["detector"].detect_batch(yg, want_logits=True)
    l = m["locator"].locate_batch(yg, want_logits=True)
    wm = wm.cpu().numpy(); lat = lat.cpu().numpy()
    lg = d["logits"][:, :, ::int(z["det_decim"])].cpu().numpy()
    avg = d["avg"].cpu().numpy(); bits = d["bits"].cpu().numpy()
    ll = l["logits"].cpu().numpy(); mask = l["mask"].cpu().numpy()
    bad_bits = int((bits != z["det_bits"]).sum())
    mm = mask.reshape(z["loc_mask"].shape) != z["loc_mask"]
    marg = np.abs(z["loc_logits"] - 0.5)
    worst = float(marg[mm.reshape(marg.shape)].max()) if mm.any() else 0.0
    inside = int((marg <= worst).sum()) if mm.any() else 0
    print(f"| {os.path.basename(path)[:-4]} | {snr_db(z['wm'], wm):.1f} | {np.abs(z['wm'] - wm).max():.2e} | {snr_db(z['latent'], lat):.1f} | "
          f"{snr_db(z['det_logits_decim'], lg):.1f} | {np.abs(z['det_logits_decim'] - lg).max():.3f} | {np.abs(avg - z['det_avg']).max():.2e} | "
          f"{bad_bits} ({np.abs(z['det_avg'] - 0.5).min():.1e}) | {snr_db(z['loc_logits'], ll):.1f} | {np.abs(z['loc_logits'] - ll).max():.3f} | "
          f"{int(mm.sum())} / {mm.size} | {worst:.4f} | {inside} |")
