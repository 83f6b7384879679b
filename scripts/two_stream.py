"""Experiment: Locator on a side stream, concurrent with the Detector (both consume y)."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
dev = torch.device("cuda:0")
mods = bench.make_models(dev)
G, D, L = mods["generator"], mods["detector"], mods["locator"]
B, T = 64, 16000
x_np, msg_np, _ = bench.synth(B, T, 1)
x = torch.from_numpy(x_np).to(dev); msg = torch.from_numpy(msg_np).to(dev)
side = torch.cuda.Stream()
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

def step_serial():
    _, y, _ = G.embed_batch(x, msg, want_wm=False)
    D.detect_batch(y); L.locate_batch(y)

def step_two():
    _, y, _ = G.embed_batch(x, msg, want_wm=False)
    ev = torch.cuda.Event(); ev.record()
    with torch.cuda.stream(side):
        side.wait_event(ev)
        l = L.locate_batch(y)
    d = D.detect_batch(y)
    torch.cuda.current_stream().wait_stream(side)

for name, fn in (("serial", step_serial), ("locator on side stream", step_two), ("serial", step_serial), ("locator on side stream", step_two)):
    for _ in range(3): fn()
    torch.cuda.synchronize(); ts = []
    for _ in range(10):
        flush.zero_(); a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    print(f"{name}: mean {sum(ts)/len(ts):.3f} ms  best {min(ts):.3f} ms", flush=True)
