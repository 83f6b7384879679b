"""Micro-benchmark: time vs problem size for the fused 1x1+dw5 GEMM (h1 / out variants) next to a
plain device copy of the same bytes (what the HBM system gives a trivially simple kernel)."""
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


def timeit(fn, reps=8):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.zero_()
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return min(ts) * 1e3


for Cc, T in [(96, 16000), (192, 8000), (64, 16000)]:
    for B in (4, 16, 64, 128):
        A = torch.randn(B, T, Cc, device=dev).to(torch.float16)
        W = (torch.randn(Cc, Cc, device=dev) / Cc ** 0.5).to(torch.float16)
        dw = torch.randn(5, Cc, device=dev) * 0.3
        bias = torch.randn(Cc, device=dev)
        R = torch.randn(B, T, Cc, device=dev).to(torch.float16)
        o1 = torch.empty_like(A); o2 = torch.empty_like(A)
        mb = B * T * Cc * 2 / 1e6
        t_h1 = timeit(lambda: L.wv_op_gemm_dw5(P(A), P(W), B, T, Cc, Cc, P(dw), P(bias), None, None, P(o2), 1.0, S()))
        t_out = timeit(lambda: L.wv_op_gemm_dw5(P(A), P(W), B, T, Cc, Cc, P(dw), P(bias), P(R), P(o1), P(o2), 0.8, S()))
        t_cp = timeit(lambda: o1.copy_(A))
        t_cp2 = timeit(lambda: (o1.copy_(A), o2.copy_(R)))
        print(f"C={Cc:3d} T={T:5d} B={B:3d} ({mb:6.1f} MB/tensor): h1 {t_h1:7.1f} us {2 * mb / t_h1:6.0f} GB/s | out {t_out:7.1f} us {4 * mb / t_out:6.0f} GB/s | "
              f"copy {t_cp:7.1f} us {2 * mb / t_cp:6.0f} GB/s | 2 copies {t_cp2:7.1f} us {4 * mb / t_cp2:6.0f} GB/s", flush=True)
