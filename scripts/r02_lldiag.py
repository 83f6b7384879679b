"""Diagnostic: distribution of |locator logit (precise CUDA) - oracle| on the 64 x 1 s batch (test_config2)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import wv_oracle as O
import test_gpu_baseline_sizes as tb
m = tb.models()
x, msg = tb.synth(64, 16000, 21)
wm, y, _ = m["generator"][0].embed_batch(x.cuda(), msg.cuda())
l = m["locator"][0].locate_batch(y, want_logits=True)
yc = y.cpu()
with torch.no_grad():
    ll_o = torch.cat([O.locator_forward(yc[i:i + 16], m["locator"][1], m["locator"][2]) for i in range(0, 64, 16)])
d = (l["logits"].cpu() - ll_o).abs().numpy().reshape(64, -1)
print("max", d.max(), "p99.99", np.quantile(d, 0.9999), "p99.9", np.quantile(d, 0.999), "median", np.median(d))
for thr in (1e-4, 3e-4, 1e-3):
    idx = np.argwhere(d > thr)
    print("count >", thr, len(idx), "clips", sorted(set(idx[:, 0].tolist()))[:10], "first", idx[:6].tolist())
b, t = np.unravel_index(d.argmax(), d.shape)
print("argmax clip", b, "t", t, "neighbourhood", d[b, max(0, t - 40):t + 40:4])
# which of the two is off at the outlier?  fp64 oracle on the affected clips
def to64(o):
    if torch.is_tensor(o): return o.double() if o.is_floating_point() else o
    if isinstance(o, dict): return {k: to64(v) for k, v in o.items()}
    if isinstance(o, (list, tuple)): return type(o)(to64(v) for v in o)
    return o
W64 = to64(m["locator"][1])
for cb in (7, 22):
    with torch.no_grad():
        r64 = O.locator_forward(yc[cb:cb + 1].double(), W64, m["locator"][2])[0, 0].numpy()
    o32 = ll_o[cb, 0].numpy().astype(np.float64); g32 = l["logits"][cb, 0].cpu().numpy().astype(np.float64)
    print("clip", cb, "max |oracle32-fp64|", np.abs(o32 - r64).max(), "at", np.abs(o32 - r64).argmax(),
          "max |cuda-fp64|", np.abs(g32 - r64).max(), "at", np.abs(g32 - r64).argmax())
# fast path on the same input for scale
L = m["locator"][0]; L.exact = False
lf = L.locate_batch(y, want_logits=True)["logits"].cpu()
print("fast path max-abs vs oracle", float((lf - ll_o).abs().max()))
