"""Device time of small-batch calls of the precise Detector / Locator nets (CUDA events, graph path warmed up)."""
import os, sys, torch, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import BASE_KW, fixture_weights
from waveverify_b200 import Detector, Locator
dev = torch.device("cuda:0")
mods = {}
for kind, cls in (("detector", Detector), ("locator", Locator)):
    _, sd = fixture_weights(kind, False, 0)
    m = cls(**{**BASE_KW[kind], "bias": True, "zero_init": False}); m.load_state_dict(sd); mods[kind] = m.to(dev)
D, L = mods["detector"], mods["locator"]
def timeit(fn, n=20):
    for _ in range(4): fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / n
for B in (1, 2, 4, 8, 64):
    x = 0.1 * torch.randn(B, 1, 16000, device=dev)
    tf = timeit(lambda: D._run(x, None, False, False)); tp = timeit(lambda: D._run(x, None, False, True))
    L.exact = False; lf = timeit(lambda: L.locate_batch(x)); L.exact = True; lp = timeit(lambda: L.locate_batch(x))
    print(f"B={B:3d} x 1 s: detector fast {tf*1e3:7.1f} us  precise {tp*1e3:7.1f} us ({D.launches(B,16000)} launches) | locator fast {lf*1e3:7.1f} us  precise {lp*1e3:7.1f} us")
