import ctypes as C, os, sys, torch
import torch.nn.functional as F
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from waveverify_b200 import _lib
L = _lib.lib(); dev = torch.device("cuda:0")
P = lambda t: None if t is None else C.c_void_p(t.data_ptr())
S = lambda: C.c_void_p(torch.cuda.current_stream().cuda_stream)

def run(B, T, Cc, W1, W2, dw1, dw2, b1, b2, X, pre=1.0, s_act=1.0):
    out = torch.full((B, T, Cc), float("nan"), dtype=torch.float16, device=dev)
    k1 = dw1[:, 0, :].t().contiguous().to(dev); k2 = dw2[:, 0, :].t().contiguous().to(dev)
    keep = [X.to(dev), W1.to(dev), b1.to(dev), W2.to(dev), b2.to(dev)]
    rc = L.wv_op_resblock(P(keep[0]), P(keep[1]), P(k1), P(keep[2]), P(keep[3]), P(k2), P(keep[4]), B, T, Cc, pre, P(out), None, s_act, S())
    assert rc == 0, L.wv_last_error()
    torch.cuda.synchronize()
    def dw5(u, w, b): return F.conv1d(F.pad(u.transpose(1, 2), (4, 0)), w, b, groups=Cc).transpose(1, 2)
    x = X.double(); a = F.elu(x * pre)
    h = F.elu(dw5(a @ W1.double().t(), dw1.double(), b1.double()))
    ref = (dw5(h @ W2.double().t(), dw2.double(), b2.double()) + x).float()
    err = (out.cpu().float() - ref).abs()
    return err, ref, out.cpu().float()

def report(name, err):
    B, T, Cc = err.shape
    print(f"{name}: max {err.max():.4f}; per-row max (first 16 rows of clip 0): {[round(v, 3) for v in err[0, :16].max(dim=1).values.tolist()]}")
    bad = (err > 0.02)
    if bad.any():
        rows = bad.any(dim=2)[0].nonzero().flatten().tolist()
        chans = bad.any(dim=1)[0].nonzero().flatten().tolist()
        print(f"   bad rows (clip 0): n={len(rows)} first {rows[:24]}  bad channels n={len(chans)} first {chans[:24]}")

g = torch.Generator().manual_seed(0)
for Cc in (64, 96):
    B, T = 1, 300
    I = torch.eye(Cc).to(torch.float16)
    delta = torch.zeros(Cc, 1, 5); delta[:, 0, 4] = 1.0
    z = torch.zeros(Cc)
    X = torch.randn(B, T, Cc, generator=g).to(torch.float16)
    Wr1 = (torch.randn(Cc, Cc, generator=g) / Cc ** 0.5).to(torch.float16)
    Wr2 = (torch.randn(Cc, Cc, generator=g) / Cc ** 0.5).to(torch.float16)
    dwr1 = torch.randn(Cc, 1, 5, generator=g) * 0.4; dwr2 = torch.randn(Cc, 1, 5, generator=g) * 0.3
    br = torch.randn(Cc, generator=g) * 0.5
    print("C =", Cc)
    report("identity everything", run(B, T, Cc, I, I, delta, delta, z, z, X)[0])
    report("W1 random", run(B, T, Cc, Wr1, I, delta, delta, z, z, X)[0])
    report("W2 random", run(B, T, Cc, I, Wr2, delta, delta, z, z, X)[0])
    report("dw1 random", run(B, T, Cc, I, I, dwr1, delta, z, z, X)[0])
    report("dw2 random", run(B, T, Cc, I, I, delta, dwr2, z, z, X)[0])
    report("bias1", run(B, T, Cc, I, I, delta, delta, br, z, X)[0])
    report("bias2", run(B, T, Cc, I, I, delta, delta, z, br, X)[0])
    report("all random", run(B, T, Cc, Wr1, Wr2, dwr1, dwr2, br, br, X, 0.866)[0])

print("---- constant input")
Cc = 64; B, T = 1, 100
I = torch.eye(Cc).to(torch.float16); delta = torch.zeros(Cc, 1, 5); delta[:, 0, 4] = 1.0; z = torch.zeros(Cc)
X = torch.zeros(B, T, Cc, dtype=torch.float16)
X[0, :, :] = (torch.arange(T).float()[:, None] * 0.01 + torch.arange(Cc).float()[None, :] * 0.001).to(torch.float16)
err, ref, out = run(B, T, Cc, I, I, delta, delta, z, z, X)
torch.set_printoptions(precision=4, linewidth=200)
print("x  ", X[0, 8:12, :8].float())
print("ref", ref[0, 8:12, :8])
print("out", out[0, 8:12, :8])
print("out rows 0..3", out[0, :4, :8])
print("out ch 56..63 row 20", out[0, 20, 56:], "ref", ref[0, 20, 56:])
