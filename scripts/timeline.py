"""Per-tile timeline of CTA 0 for the fused 1x1+dw5 GEMM (WV_TIMELINE=1 build probe)."""
import ctypes as C, os, sys, torch
os.environ["WV_TIMELINE"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from waveverify_b200 import _lib
L = _lib.lib(); dev = torch.device("cuda:0")
P = lambda t: None if t is None else C.c_void_p(t.data_ptr())
S = lambda: C.c_void_p(torch.cuda.current_stream().cuda_stream)
Cc, T, B = int(sys.argv[1]), int(sys.argv[2]), 64
A = torch.randn(B, T, Cc, device=dev).to(torch.float16)
W = (torch.randn(Cc, Cc, device=dev) / Cc ** 0.5).to(torch.float16)
dw = torch.randn(5, Cc, device=dev) * 0.3; bias = torch.randn(Cc, device=dev)
R = torch.randn(B, T, Cc, device=dev).to(torch.float16)
o1 = torch.empty_like(A); o2 = torch.empty_like(A)
print("== h1"); sys.stdout.flush()
L.wv_op_gemm_dw5(P(A), P(W), B, T, Cc, Cc, P(dw), P(bias), None, None, P(o2), 1.0, S())
print("== out"); sys.stdout.flush()
L.wv_op_gemm_dw5(P(A), P(W), B, T, Cc, Cc, P(dw), P(bias), P(R), P(o1), P(o2), 0.8, S())
