"""Single-clip latency (BASELINE configs[0] shape: 1 x 1 s): embed -> detect -> locate, CUDA events."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
dev = torch.device("cuda:0")
mods = bench.make_models(dev)
G, D, L = mods["generator"], mods["detector"], mods["locator"]
for B, T in ((1, 16000), (1, 160000), (8, 16000)):
    x_np, msg_np, _ = bench.synth(B, T, 1)
    x = torch.from_numpy(x_np).to(dev); msg = torch.from_numpy(msg_np).to(dev)
    def step():
        _, y, _ = G.embed_batch(x, msg, want_wm=False)
        D.detect_batch(y); L.locate_batch(y)
    for _ in range(5): step()
    torch.cuda.synchronize()
    ts = []
    for _ in range(20):
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(); step(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    ts.sort()
    print(f"B={B} T={T}: median {ts[10]:.3f} ms, best {ts[0]:.3f} ms  (WV_PDL={os.environ.get('WV_PDL', '0')})", flush=True)
