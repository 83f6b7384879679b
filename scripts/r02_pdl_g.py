"""Device time of the three nets separately on the 64 x 1 s batch (L2 flushed between calls), for A/B runs of launch
options (WV_PDL, WV_GRAPH_MAX_SAMPLES)."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
dev = torch.device("cuda:0")
mods = bench.make_models(dev)
G, D, L = mods["generator"], mods["detector"], mods["locator"]
x_np, msg_np, _ = bench.synth(64, 16000, 100)
x = torch.from_numpy(x_np).to(dev); msg = torch.from_numpy(msg_np).to(dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
_, y, _ = G.embed_batch(x, msg, want_wm=False)
def timeit(fn, n=10):
    for _ in range(3): fn()
    ts = []
    for _ in range(n):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize(); ts.append(s.elapsed_time(e))
    return min(ts), sum(ts) / len(ts)
D.exact_bits = False
for name, fn in (("G", lambda: G.embed_batch(x, msg, want_wm=False)), ("D fast", lambda: D.detect_batch(y)), ("L", lambda: L.locate_batch(y))):
    mn, av = timeit(fn)
    print(f"{name}: min {mn*1e3:.0f} us  avg {av*1e3:.0f} us")
D.exact_bits = True
mn, av = timeit(lambda: D.detect_batch(y)); print(f"D exact: min {mn*1e3:.0f} us avg {av*1e3:.0f} us")
