"""Fused resblock kernel (one launch) against the two-launch form (1x1 + dw5 + ELU ; 1x1 + dw5 + residual + 2 outputs)
at the layer shapes where the fused kernel applies.  L2 flushed before every timed call; min of 10."""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from waveverify_b200 import _lib  # noqa: E402

L = _lib.lib()
dev = torch.device("cuda:0")
P = lambda t: None if t is None else C.c_void_p(t.data_ptr())
S = lambda: C.c_void_p(torch.cuda.current_stream().cuda_stream)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, reps=10):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.zero_()
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return min(ts) * 1e3


shapes = [(64, 16000, 32), (64, 16000, 64), (64, 16000, 96), (64, 8000, 128)]
if len(sys.argv) > 1:
    shapes = [s for s in shapes if s[2] in [int(v) for v in sys.argv[1:]]]
for (B, T, Cc) in shapes:
    X = torch.randn(B, T, Cc, device=dev).to(torch.float16)
    A = torch.nn.functional.elu(X.float() * 0.9).to(torch.float16)
    W1 = (torch.randn(Cc, Cc, device=dev) / Cc ** 0.5).to(torch.float16); W2 = (torch.randn(Cc, Cc, device=dev) / Cc ** 0.5).to(torch.float16)
    k1 = torch.randn(5, Cc, device=dev) * 0.3; k2 = torch.randn(5, Cc, device=dev) * 0.3
    b1 = torch.randn(Cc, device=dev); b2 = torch.randn(Cc, device=dev)
    H = torch.empty_like(X); o1 = torch.empty_like(X); o2 = torch.empty_like(X)
    mb = B * T * Cc * 2 / 1e6
    t_h1 = timeit(lambda: L.wv_op_gemm_dw5(P(A), P(W1), B, T, Cc, Cc, P(k1), P(b1), None, None, P(H), 1.0, S()))
    t_out = timeit(lambda: L.wv_op_gemm_dw5(P(H), P(W2), B, T, Cc, Cc, P(k2), P(b2), P(X), P(o1), P(o2), 0.7, S()))
    t_out1 = timeit(lambda: L.wv_op_gemm_dw5(P(H), P(W2), B, T, Cc, Cc, P(k2), P(b2), P(X), None, P(o2), 0.7, S()))
    t_f2 = timeit(lambda: L.wv_op_resblock(P(X), P(A), P(W1), P(k1), P(b1), P(W2), P(k2), P(b2), B, T, Cc, 0.9, P(o1), P(o2), 0.7, S()))
    t_f1 = timeit(lambda: L.wv_op_resblock(P(X), P(A), P(W1), P(k1), P(b1), P(W2), P(k2), P(b2), B, T, Cc, 0.9, None, P(o2), 0.7, S()))
    print(f"C={Cc:4d} T={T:6d}: two launches {t_h1:6.1f} + {t_out:6.1f} = {t_h1 + t_out:6.1f} us ({6 * mb / (t_h1 + t_out):5.0f} GB/s) | fused {t_f2:6.1f} us ({4 * mb / t_f2:5.0f} GB/s)"
          f" || act only: {t_h1:6.1f} + {t_out1:6.1f} = {t_h1 + t_out1:6.1f} | fused {t_f1:6.1f} us ({3 * mb / t_f1:5.0f} GB/s)", flush=True)
