import ctypes as C, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from waveverify_b200 import _lib
L = _lib.lib(); dev = torch.device("cuda:0")
P = lambda t: None if t is None else C.c_void_p(t.data_ptr())
S = lambda: C.c_void_p(torch.cuda.current_stream().cuda_stream)
B, T, Cc = 64, 16000, 64
A = torch.randn(B, T, Cc, device=dev).to(torch.float16)
W = (torch.randn(Cc, Cc, device=dev) / Cc ** 0.5).to(torch.float16)
dw = torch.randn(5, Cc, device=dev) * 0.3; bias = torch.randn(Cc, device=dev)
o1 = torch.empty(B, T, Cc, device=dev, dtype=torch.float16)
M = B * T
for _ in range(2):
    L.wv_op_gemm(P(A), Cc, P(W), Cc, M, Cc, Cc, None, None, P(o1), None, 1.0, 0, S())
    L.wv_op_gemm_dw5(P(A), P(W), B, T, Cc, Cc, P(dw), P(bias), None, None, P(o1), 1.0, S())
torch.cuda.synchronize()
torch.cuda.profiler.start()
L.wv_op_gemm(P(A), Cc, P(W), Cc, M, Cc, Cc, None, None, P(o1), None, 1.0, 0, S())
L.wv_op_gemm_dw5(P(A), P(W), B, T, Cc, Cc, P(dw), P(bias), None, None, P(o1), 1.0, S())
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("done")
