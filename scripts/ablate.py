"""Ablation timing of the fused 1x1+dw5 GEMM (-DWV_TIMELINE build): full kernel vs no global stores vs
no epilogue math vs no drain.  WV_TIMELINE=10+mode selects the mode without the clock probes."""
import ctypes as C, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from waveverify_b200 import _lib
L = _lib.lib(); dev = torch.device("cuda:0")
P = lambda t: None if t is None else C.c_void_p(t.data_ptr())
S = lambda: C.c_void_p(torch.cuda.current_stream().cuda_stream)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def timeit(fn, reps=6):
    fn(); torch.cuda.synchronize(); ts = []
    for _ in range(reps):
        flush.zero_(); a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return min(ts) * 1e3
for Cc, T in [(96, 16000), (64, 16000), (128, 8000), (192, 8000)]:
    B = 64
    A = torch.randn(B, T, Cc, device=dev).to(torch.float16); W = (torch.randn(Cc, Cc, device=dev) / Cc ** 0.5).to(torch.float16)
    dw = torch.randn(5, Cc, device=dev) * 0.3; bias = torch.randn(Cc, device=dev); R = torch.randn(B, T, Cc, device=dev).to(torch.float16)
    o1 = torch.empty_like(A); o2 = torch.empty_like(A)
    row = []
    for mode in (0, 1, 2, 4, 6, 5):
        os.environ["WV_TIMELINE"] = str(10 + mode)
        t_h1 = timeit(lambda: L.wv_op_gemm_dw5(P(A), P(W), B, T, Cc, Cc, P(dw), P(bias), None, None, P(o2), 1.0, S()))
        t_out = timeit(lambda: L.wv_op_gemm_dw5(P(A), P(W), B, T, Cc, Cc, P(dw), P(bias), P(R), P(o1), P(o2), 0.8, S()))
        row.append(f"mode{mode}: h1 {t_h1:6.1f} out {t_out:6.1f}")
    print(f"C={Cc:3d} T={T}: " + " | ".join(row), flush=True)
