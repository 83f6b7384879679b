"""Per-launch table (tag, device time from CUDA events, algorithmic GB/s and TFLOP/s)."""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--clips", type=int, default=64)
ap.add_argument("--seconds", type=float, default=1.0)
ap.add_argument("--reps", type=int, default=3)
a = ap.parse_args()
dev = torch.device("cuda:0")
mods = bench.make_models(dev)
T = int(a.seconds * 16000)
x_np, msg_np, gt_np = bench.synth(a.clips, T, 100)
x = torch.from_numpy(x_np).to(dev); msg = torch.from_numpy(msg_np).to(dev)
G, D, L = mods["generator"], mods["detector"], mods["locator"]
_, y, _ = G.embed_batch(x, msg); D.detect_batch(y); L.locate_batch(y)
for m in (G, D, L):
    m.set_profile(True)
best = {}
for rep in range(a.reps):
    recs = []
    _, y, _ = G.embed_batch(x, msg, want_wm=False); torch.cuda.synchronize(); recs += [("G", r) for r in G.profile_read()]
    D.detect_batch(y); torch.cuda.synchronize(); recs += [("D", r) for r in D.profile_read()]
    L.locate_batch(y); torch.cuda.synchronize(); recs += [("L", r) for r in L.profile_read()]
    for i, (n, r) in enumerate(recs):
        if i not in best or r["ms"] < best[i][1]["ms"]:
            best[i] = (n, r)
tot = sum(r["ms"] for _, r in best.values())
print(f"total {tot:.3f} ms for {a.clips * a.seconds:g} audio-s -> {a.clips * a.seconds / tot * 1e3:.0f} audio-s/s (sum of per-launch minima)")
for i in sorted(best):
    n, r = best[i]
    print(f"{n} {r['tag']:<22s} cls {r['cls']:3d} {r['ms'] * 1e3:9.1f} us  {r['bytes'] / 1e6 / max(r['ms'], 1e-9):8.1f} GB/s  {r['flops'] / 1e9 / max(r['ms'], 1e-9):8.1f} TF/s  {r['bytes'] / 1e6:9.2f} MB")
