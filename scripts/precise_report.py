"""Precise (fp32-accurate) path against the committed outputs of the reference (tests/golden):
locator logits / mask on the precise net, detector on the precise net (whole batch) and through the
default fast path + re-check.  Prints a markdown table (kept under profiles/).  WV_TAPS=1 also compares
every launch's output of the precise locator with the fp64 oracle (debug)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
from helpers import BASE_KW, fixture_weights, golden_cases, load_case, snr_db  # noqa: E402
from waveverify_b200 import Detector, Locator  # noqa: E402

dev = torch.device("cuda:0")
cache = {}


def models(zi, seed):
    key = (bool(zi), int(seed))
    if key not in cache:
        out = {}
        for kind, cls in (("detector", Detector), ("locator", Locator)):
            c, sd = fixture_weights(kind, key[0], key[1])
            m = cls(**{**BASE_KW[kind], "bias": True, "zero_init": key[0]})
            m.load_state_dict(sd)
            out[kind] = m.to(dev)
        cache[key] = out
    return cache[key]


print("| case | loc max-abs (precise) | mask mismatches / samples | max margin of a mismatch | loc max-abs (fast) | fast mismatches | det logits max-abs (precise) | avg max-abs (precise) | avg max-abs (fast) | bit mismatches default path | clips re-checked | min ref margin |")
print("|---|---|---|---|---|---|---|---|---|---|---|---|")
for path in golden_cases():
    z = load_case(path)
    m = models(z["zero_init"], z["wseed"])
    yg = torch.from_numpy(z["y"]).to(dev)
    L, D = m["locator"], m["detector"]
    L.exact = True
    lp = L.locate_batch(yg, want_logits=True)
    L.exact = False
    lf = L.locate_batch(yg, want_logits=True)
    L.exact = True
    llp = lp["logits"].cpu().numpy(); llf = lf["logits"].cpu().numpy()
    marg = np.abs(z["loc_logits"] - 0.5)
    mm = lp["mask"].cpu().numpy().reshape(z["loc_mask"].shape) != z["loc_mask"]
    mmf = lf["mask"].cpu().numpy().reshape(z["loc_mask"].shape) != z["loc_mask"]
    worst = float(marg[mm.reshape(marg.shape)].max()) if mm.any() else 0.0
    dd = int(z["det_decim"])
    dp = D.detect_batch(yg, want_logits=True, precise=True)
    D.exact_bits = False
    df = D.detect_batch(yg)
    D.exact_bits = True
    n0 = D.recheck_count
    de = D.detect_batch(yg)
    lgp = dp["logits"][:, :, ::dd].cpu().numpy()
    print(f"| {os.path.basename(path)[:-4]} | {np.abs(z['loc_logits'] - llp).max():.2e} | {int(mm.sum())} / {mm.size} | {worst:.2e} | "
          f"{np.abs(z['loc_logits'] - llf).max():.2e} | {int(mmf.sum())} | {np.abs(z['det_logits_decim'] - lgp).max():.2e} | "
          f"{np.abs(dp['avg'].cpu().numpy() - z['det_avg']).max():.2e} | {np.abs(df['avg'].cpu().numpy() - z['det_avg']).max():.2e} | "
          f"{int((de['bits'].cpu().numpy() != z['det_bits']).sum())} | {D.recheck_count - n0} / {yg.shape[0]} | {np.abs(z['det_avg'] - 0.5).min():.1e} |")

if os.environ.get("WV_TAPS"):
    import wv_oracle as O
    from helpers import oracle_cfg
    z = load_case(golden_cases()[0])
    c, sd = fixture_weights("locator", z["zero_init"], z["wseed"])
    L = models(z["zero_init"], z["wseed"])["locator"]
    W = O.fold_state_dict(sd, dtype=torch.float64)
    taps = {}
    y64 = torch.from_numpy(z["y"]).double()
    with torch.no_grad():
        O.locator_forward(y64, W, oracle_cfg(c), taps) if "taps" in O.locator_forward.__code__.co_varnames else None
    print("oracle taps:", sorted(taps.keys())[:40])
