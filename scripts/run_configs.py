"""BASELINE.json configs 3-5 at their full per-GPU sizes on ONE B200 (CUDA-event timing, 1 warm-up
+ 3 timed passes each).  Prints a markdown table (kept under profiles/)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from waveverify_b200 import metric_counters  # noqa: E402
from waveverify_b200.api import embed_streaming  # noqa: E402

dev = torch.device("cuda:0")
mods = bench.make_models(dev)
G, D, L = mods["generator"], mods["detector"], mods["locator"]
SR = 16000


def timed(fn, reps=3):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) / 1e3)
    return min(ts), sum(ts) / len(ts)


rows = []
# config 3: 512 x 10 s over 8 GPUs -> 64 x 10 s per GPU, embed + detect + locate + counters
B, T = 64, 160000
x_np, msg_np, gt_np = bench.synth(B, T, 3)
x = torch.from_numpy(x_np).to(dev); msg = torch.from_numpy(msg_np).to(dev); gt = torch.from_numpy(gt_np).to(dev)
cnt = torch.zeros(6, dtype=torch.int64, device=dev)
for m in (G, D, L):
    m.set_chunk_samples(64 * SR)          # sub-batches of 64 audio-seconds


def c3():
    _, y, _ = G.embed_batch(x, msg, want_wm=False)
    d = D.detect_batch(y); l = L.locate_batch(y)
    metric_counters(d["bits"], d["valid"], msg, l["mask"], gt, counters=cnt)


tmin, tavg = timed(c3)
rows.append(("3: 64 x 10 s per GPU (= 512 x 10 s over 8), embed+detect+locate, 64 audio-s sub-batches", B * T / SR, tmin, tavg))
del x, gt

# config 4: Detector + Locator only on 4096 x 5 s (sub-batched)
B, T = 4096, 80000
y = (0.1 * torch.randn(B, 1, T, device=dev, generator=torch.Generator(device=dev).manual_seed(4)))
bits_acc = []


def c4():
    d = D.detect_batch(y); l = L.locate_batch(y)
    bits_acc.append(int(d["bits"].sum().item()) + int(l["mask"][:8].sum().item()))


tmin, tavg = timed(c4, reps=2)
rows.append(("4: Detector + Locator on 4096 x 5 s, 64 audio-s sub-batches", B * T / SR, tmin, tavg))
assert len(set(bits_acc)) == 1, "non-deterministic detect/locate"
del y

# config 5: one 10-minute clip streamed through the Generator in 20 s chunks with a 5440-sample halo
T = 600 * SR
xl = 0.1 * torch.randn(1, 1, T, device=dev, generator=torch.Generator(device=dev).manual_seed(5))
m1 = msg[:1]
for m in (G, D, L):
    m.set_chunk_samples(0)


def c5():
    return embed_streaming(G, xl, m1, chunk_samples=320000)


tmin, tavg = timed(c5)
rows.append(("5: 10-minute clip streamed through the Generator (20 s chunks, 5440-sample halo)", T / SR, tmin, tavg))

print("| config | audio-s per pass | best pass (s) | mean pass (s) | audio-s/s (best) |")
print("|---|---|---|---|---|")
for name, aud, tmin, tavg in rows:
    print(f"| {name} | {aud:.0f} | {tmin:.4f} | {tavg:.4f} | {aud / tmin:.0f} |")
