"""Micro-benchmark: plain staged GEMM (taps=1) vs fused depthwise epilogue (taps=5) at layer shapes."""
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


for (B, T, Cc) in [(64, 16000, 64), (64, 16000, 96), (64, 8000, 128), (64, 8000, 192), (64, 2000, 256), (64, 2000, 384), (64, 400, 512), (64, 400, 768)]:
    A = torch.randn(B, T, Cc, device=dev).to(torch.float16)
    W = (torch.randn(Cc, Cc, device=dev) / Cc ** 0.5).to(torch.float16)
    dw = torch.randn(5, Cc, device=dev) * 0.3
    bias = torch.randn(Cc, device=dev)
    R = torch.randn(B, T, Cc, device=dev).to(torch.float16)
    o1 = torch.empty(B, T, Cc, device=dev, dtype=torch.float16)
    o2 = torch.empty(B, T, Cc, device=dev, dtype=torch.float16)
    M = B * T
    mb = M * Cc * 2 / 1e6
    t_plain_raw = timeit(lambda: L.wv_op_gemm(P(A), Cc, P(W), Cc, M, Cc, Cc, None, None, P(o1), None, 1.0, 0, S()))
    t_plain_act = timeit(lambda: L.wv_op_gemm(P(A), Cc, P(W), Cc, M, Cc, Cc, P(bias), None, None, P(o2), 1.0, 0, S()))
    t_plain_res = timeit(lambda: L.wv_op_gemm(P(A), Cc, P(W), Cc, M, Cc, Cc, P(bias), P(R), P(o1), P(o2), 0.8, 0, S()))
    t_dw_act = timeit(lambda: L.wv_op_gemm_dw5(P(A), P(W), B, T, Cc, Cc, P(dw), P(bias), None, None, P(o2), 1.0, S()))
    t_dw_res = timeit(lambda: L.wv_op_gemm_dw5(P(A), P(W), B, T, Cc, Cc, P(dw), P(bias), P(R), P(o1), P(o2), 0.8, S()))
    print(f"C={Cc:4d} T={T:6d}: plain raw {t_plain_raw:7.1f} us ({2 * mb / t_plain_raw:6.0f} GB/s) | plain bias+act {t_plain_act:7.1f} ({2 * mb / t_plain_act:6.0f}) | "
          f"plain res+2out {t_plain_res:7.1f} ({4 * mb / t_plain_res:6.0f}) | dw5 act {t_dw_act:7.1f} ({2 * mb / t_dw_act:6.0f}) | dw5 res+2out {t_dw_res:7.1f} ({4 * mb / t_dw_res:6.0f})", flush=True)
