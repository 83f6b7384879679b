"""Per-pair timeline of CTA 0 for the fused resblock kernel (-DWV_TIMELINE build, WV_TIMELINE_RB=1)."""
import ctypes as C, os, sys, torch
if "--probe" in sys.argv: os.environ["WV_TIMELINE_RB"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from waveverify_b200 import _lib
L = _lib.lib(); dev = torch.device("cuda:0")
P = lambda t: None if t is None else C.c_void_p(t.data_ptr())
S = lambda: C.c_void_p(torch.cuda.current_stream().cuda_stream)
Cc, T, B = int(sys.argv[1]), int(sys.argv[2]), 64
X = torch.randn(B, T, Cc, device=dev).to(torch.float16)
W1 = (torch.randn(Cc, Cc, device=dev) / Cc ** 0.5).to(torch.float16); W2 = (torch.randn(Cc, Cc, device=dev) / Cc ** 0.5).to(torch.float16)
k1 = torch.randn(5, Cc, device=dev) * 0.3; k2 = torch.randn(5, Cc, device=dev) * 0.3
b1 = torch.randn(Cc, device=dev); b2 = torch.randn(Cc, device=dev)
o1 = torch.empty_like(X); o2 = torch.empty_like(X)
sys.stdout.flush()
A = torch.nn.functional.elu(X.float() * 0.9).to(torch.float16)
L.wv_op_resblock(P(X), P(A), P(W1), P(k1), P(b1), P(W2), P(k2), P(b2), B, T, Cc, 0.9, P(o1), P(o2), 0.7, S())
