"""Target program for ncu: one warm pass, then ONE profiled embed+detect+locate pass between
cudaProfilerStart/Stop (use `ncu --profile-from-start off`).  Same workload as bench.py."""
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
ap.add_argument("--which", default="gdl")
a = ap.parse_args()
dev = torch.device("cuda:0")
mods = bench.make_models(dev)
T = int(a.seconds * 16000)
x_np, msg_np, gt_np = bench.synth(a.clips, T, 100)
x = torch.from_numpy(x_np).to(dev); msg = torch.from_numpy(msg_np).to(dev)


def step():
    y = x
    if "g" in a.which:
        _, y, _ = mods["generator"].embed_batch(x, msg, want_wm=False)
    if "d" in a.which:
        mods["detector"].detect_batch(y)
    if "l" in a.which:
        mods["locator"].locate_batch(y)


step(); step()
torch.cuda.synchronize()
torch.cuda.profiler.start()
step()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("done")
