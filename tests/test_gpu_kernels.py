"""GPU: each hand-written kernel against a plain fp32 PyTorch statement of the same op, called
through the C ABI (wv_op_*)."""
import ctypes as C

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _lib():
    from waveverify_b200 import _lib as L
    return L.lib()


def P(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def S():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def half_close(out, ref, rel=2 ** -9, abs_=2e-3, mid=None):
    """Tolerance of the 16-bit path.  Activations are fp16 (11 significant bits, half-ulp 2^-12
    relative); the staged epilogue rounds the GEMM tile to fp16, then taps / bias / residual / scale
    each round once more (packed half2 arithmetic), and the half2 ELU carries <= 1e-3 absolute error
    on its (-1, 0] branch: rel = 2^-9 and abs >= 2e-3 cover that chain with margin."""
    err = (out.float() - ref).abs()
    tol = rel * ref.abs() + abs_
    if mid is not None:          # fp16 rounding of the staged accumulator, before bias / residual
        tol = tol + 2 * rel * mid.abs()
    assert bool((err <= tol).all()), f"max err {err.max().item()} (ref max {ref.abs().max().item()})"


GEMM_CASES = [
    # M, N, K, bias, residual, act
    (128, 64, 64, False, False, False),
    (1, 32, 32, True, False, True),            # single row, locator width
    (1000, 64, 64, True, True, True),
    (4096, 256, 256, False, False, False),
    (300, 96, 192, True, False, True),          # decoder widths, BLOCK_N=96
    (777, 1536, 128, False, True, False),       # 6 N tiles
    (5000, 128, 33, False, True, True),         # spec 1x1: K=33 zero-filled to 64 by TMA
    (513, 1024, 513, False, True, True),        # K tail 513 -> 9 k-blocks
    (333, 768, 1536, True, False, False),       # long K
    (64000, 192, 192, False, True, True),       # many M tiles per CTA (persistent loop, both TMEM stages)
    (40000, 384, 384, False, False, True),
    # CTA-pair mode (cluster of 2, tcgen05 cta_group::2): K >= 256, an even number of n tiles, enough tile pairs
    (64000, 256, 256, True, True, True),        # one wide n tile
    (40000, 512, 512, False, True, True),       # two wide n tiles, odd number of M tiles (the last one is computed twice)
    (30000, 768, 768, True, False, True),       # three wide n tiles
    (25600, 1024, 512, False, False, True),
]


@pytest.mark.parametrize("M,N,K,bias,res,act", GEMM_CASES)
def test_gemm_tcgen05(M, N, K, bias, res, act):
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(M * 7 + N * 3 + K)
    lda = (K + 7) // 8 * 8
    A = torch.zeros(M, lda, dtype=torch.float16)
    A[:, :K] = torch.randn(M, K, generator=g).to(torch.float16)
    W = torch.zeros(N, lda, dtype=torch.float16)
    W[:, :K] = (torch.randn(N, K, generator=g) / K ** 0.5).to(torch.float16)
    b = torch.randn(N, generator=g).to(dev) if bias else None
    R = torch.randn(M, N, generator=g).to(torch.float16).to(dev) if res else None
    A, W = A.to(dev), W.to(dev)
    out = torch.full((M, N), float("nan"), dtype=torch.float16, device=dev)
    outa = torch.full((M, N), float("nan"), dtype=torch.float16, device=dev) if act else None
    rc = _lib().wv_op_gemm(P(A), lda, P(W), lda, M, N, K, P(b), P(R), P(out), P(outa), 0.8, 0, S())
    assert rc == 0, _lib().wv_last_error()
    torch.cuda.synchronize()
    acc = A[:, :K].double() @ W[:, :K].double().t()
    ref = acc
    if bias:
        ref = ref + b.double()
    if res:
        ref = ref + R.double()
    ref = ref.float()
    half_close(out, ref, mid=acc.float())
    if act:
        half_close(outa, F.elu(ref * 0.8), abs_=3e-3, mid=acc.float())


def test_gemm_fp16_operands():
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(1)
    M, N, K = 3000, 128, 128
    A = torch.randn(M, K, generator=g).half().to(dev)
    W = (torch.randn(N, K, generator=g) / K ** 0.5).half().to(dev)
    out = torch.empty(M, N, dtype=torch.float16, device=dev)
    rc = _lib().wv_op_gemm(P(A), K, P(W), K, M, N, K, None, None, P(out), None, 1.0, 1, S())
    assert rc == 0, _lib().wv_last_error()
    torch.cuda.synchronize()
    half_close(out, (A.double() @ W.double().t()).float())


def _cl(x):  # [B,C,T] fp32 -> channels-last fp16 [B,T,C]
    return x.transpose(1, 2).contiguous().to(torch.float16)


@pytest.mark.parametrize("B,T,C", [(2, 1000, 64), (1, 7, 32), (3, 4097, 96), (1, 50, 1536), (2, 333, 384)])
@pytest.mark.parametrize("mode", ["act", "res_both", "raw_nobias"])
def test_dw5(B, T, C, mode):
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(T + C)
    x = torch.randn(B, C, T, generator=g)
    w = torch.randn(C, 1, 5, generator=g) * 0.4
    b = None if mode == "raw_nobias" else torch.randn(C, generator=g)
    r = torch.randn(B, C, T, generator=g) if mode == "res_both" else None
    xin = _cl(x).to(dev)
    wk = w[:, 0, :].t().contiguous().to(dev)             # [5][C]
    out_raw = torch.empty(B, T, C, dtype=torch.float16, device=dev) if mode != "act" else None
    out_act = torch.empty(B, T, C, dtype=torch.float16, device=dev) if mode != "raw_nobias" else None
    rin = _cl(r).to(dev) if r is not None else None
    rc = _lib().wv_op_dw5(P(xin), P(wk), P(b.to(dev)) if b is not None else None, P(rin), P(out_raw), P(out_act),
                          0.7, B, T, C, S())
    assert rc == 0, _lib().wv_last_error()
    torch.cuda.synchronize()
    xq = xin.float().cpu().transpose(1, 2)
    ref = F.conv1d(F.pad(xq, (4, 0)), w, b, groups=C)
    if r is not None:
        ref = ref + rin.float().cpu().transpose(1, 2)
    ref = ref.transpose(1, 2)
    if out_raw is not None:
        half_close(out_raw.cpu(), ref)
    if out_act is not None:
        half_close(out_act.cpu(), F.elu(ref * 0.7), abs_=2e-3)


@pytest.mark.parametrize("B,Tin,C,r", [(2, 1000, 128, 2), (1, 16001, 64, 4), (3, 401, 256, 5), (2, 77, 1024, 8), (1, 3, 64, 8)])
@pytest.mark.parametrize("film", [False, True])
def test_down_conv_film(B, Tin, C, r, film):
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(Tin + C + r)
    x = torch.randn(B, C, Tin, generator=g)
    w = torch.randn(C, 1, 2 * r, generator=g) * 0.3
    b = torch.randn(C, generator=g)
    xin = _cl(x).to(dev)
    wk = w[:, 0, :].t().contiguous().to(dev)
    To = -(-Tin // r)
    ft = torch.randn(B, 3, 4, 2, generator=g) if film else None    # [B, scales, bands, 2]; use scale 1
    out_raw = torch.empty(B, To, C, dtype=torch.float16, device=dev)
    out_act = torch.empty(B, To, C, dtype=torch.float16, device=dev)
    ftd = ft.to(dev) if film else None
    fptr = C_void(ftd[:, 1]) if film else None
    rc = _lib().wv_op_down(P(xin), P(wk), P(b.to(dev)), fptr, 3 * 4 * 2, 4, P(out_raw), P(out_act), 0.9,
                           B, Tin, C, r, S())
    assert rc == 0, _lib().wv_last_error()
    torch.cuda.synchronize()
    xq = xin.float().cpu().transpose(1, 2)
    extra = (To - 1) * r + 2 * r - r - Tin                       # modules/conv.py:160-203
    ref = F.conv1d(F.pad(xq, (r, max(0, extra))), w, b, stride=r, groups=C)
    assert ref.shape[-1] == To
    if film:
        gam = ft[:, 1, :, 0].repeat_interleave(C // 4, dim=1).unsqueeze(-1)
        bet = ft[:, 1, :, 1].repeat_interleave(C // 4, dim=1).unsqueeze(-1)
        ref = ref * gam + bet
    ref = ref.transpose(1, 2)
    half_close(out_raw.cpu(), ref, abs_=2e-3)
    half_close(out_act.cpu(), F.elu(ref * 0.9), abs_=3e-3)


def C_void(t):
    # pointer to the first element of a (possibly strided) view
    return C.c_void_p(t.data_ptr())


@pytest.mark.parametrize("B,Tin,C,r", [(2, 50, 1536, 8), (1, 400, 768, 5), (2, 1001, 384, 4), (1, 8000, 192, 2), (1, 1, 96, 2)])
def test_up_conv_transposed(B, Tin, C, r):
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(Tin + C + r)
    x = torch.randn(B, C, Tin, generator=g)
    w = torch.randn(C, 1, 2 * r, generator=g) * 0.3
    xin = _cl(x).to(dev)
    wk = w[:, 0, :].t().contiguous().to(dev)
    out = torch.empty(B, Tin * r, C, dtype=torch.float16, device=dev)
    rc = _lib().wv_op_up(P(xin), P(wk), P(out), B, Tin, C, r, S())
    assert rc == 0, _lib().wv_last_error()
    torch.cuda.synchronize()
    xq = xin.float().cpu().transpose(1, 2)
    ref = F.conv_transpose1d(xq, w, None, stride=r, groups=C)[..., : Tin * r]   # modules/conv.py:838-874
    half_close(out.cpu(), ref.transpose(1, 2), abs_=2e-3)


@pytest.mark.parametrize("B,T,N,K", [(2, 1000, 64, 64), (1, 50, 1536, 128), (3, 401, 96, 96), (2, 124, 256, 256),
                                     (1, 125, 192, 192), (2, 4097, 384, 384), (1, 3, 32, 32), (1, 16000, 128, 128),
                                     (32, 2000, 256, 256), (33, 401, 512, 512), (64, 400, 768, 768)])   # CTA-pair mode
@pytest.mark.parametrize("mode", ["act", "res_both"])
def test_gemm_with_fused_depthwise_epilogue(B, T, N, K, mode):
    """1x1 conv -> causal depthwise k=5 (+bias, +residual) -> raw / ELU outputs in ONE kernel;
    per-clip tiles of 128 rows with a 4-row halo (first tile starts at t=-4: TMA zero fill)."""
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(T + N + K)
    A = torch.randn(B, T, K, generator=g).to(torch.float16).to(dev)
    W = (torch.randn(N, K, generator=g) / K ** 0.5).to(torch.float16).to(dev)
    dw = (torch.randn(N, 1, 5, generator=g) * 0.4)
    b = torch.randn(N, generator=g)
    R = torch.randn(B, T, N, generator=g).to(torch.float16).to(dev) if mode == "res_both" else None
    wk = dw[:, 0, :].t().contiguous().to(dev)
    out_raw = torch.full((B, T, N), float("nan"), dtype=torch.float16, device=dev) if mode == "res_both" else None
    out_act = torch.full((B, T, N), float("nan"), dtype=torch.float16, device=dev)
    rc = _lib().wv_op_gemm_dw5(P(A), P(W), B, T, N, K, P(wk), P(b.to(dev)), P(R), P(out_raw), P(out_act), 0.75, S())
    assert rc == 0, _lib().wv_last_error()
    torch.cuda.synchronize()
    G = (A.double() @ W.double().t()).to(torch.float16).float().cpu()      # the kernel stages the GEMM tile in fp16
    ref = F.conv1d(F.pad(G.transpose(1, 2), (4, 0)), dw, b, groups=N).transpose(1, 2)
    # a staged element may round the other way than the emulation (fp32 summation order): bound by
    # one fp16 ulp of |G| pushed through |w| (bound kept generous)
    mid = F.conv1d(F.pad(G.abs().transpose(1, 2), (4, 0)), dw.abs(), None, groups=N).transpose(1, 2)
    if R is not None:
        ref = ref + R.float().cpu()
        half_close(out_raw.cpu(), ref, abs_=4e-3, mid=mid)
    half_close(out_act.cpu(), F.elu(ref * 0.75), abs_=4e-3, mid=mid)


@pytest.mark.parametrize("B,T,C", [(2, 1000, 64), (1, 120, 96), (3, 401, 96), (1, 7, 32), (2, 16000, 96), (5, 241, 64),
                                   (1, 121, 128), (64, 2000, 32)])
@pytest.mark.parametrize("mode", ["raw", "act", "both"])
def test_fused_resblock(B, T, C, mode):
    """One SEANet residual block (modules/seanet.py:245-281) as ONE kernel: the intermediate h never leaves
    shared memory; tiles of 128 rows with an 8-row causal halo (120 outputs), two tiles in flight per CTA."""
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(B * 131 + T * 7 + C)
    X = torch.randn(B, T, C, generator=g).to(torch.float16)
    W1 = (torch.randn(C, C, generator=g) / C ** 0.5).to(torch.float16)
    W2 = (torch.randn(C, C, generator=g) / C ** 0.5).to(torch.float16)
    dw1 = torch.randn(C, 1, 5, generator=g) * 0.4
    dw2 = torch.randn(C, 1, 5, generator=g) * 0.3
    b1 = torch.randn(C, generator=g) * 0.5
    b2 = torch.randn(C, generator=g) * 0.5
    pre, s_act = 0.866, 0.7071
    out_raw = torch.full((B, T, C), float("nan"), dtype=torch.float16, device=dev) if mode != "act" else None
    out_act = torch.full((B, T, C), float("nan"), dtype=torch.float16, device=dev) if mode != "raw" else None
    k1 = dw1[:, 0, :].t().contiguous().to(dev); k2 = dw2[:, 0, :].t().contiguous().to(dev)
    A = F.elu(X.float() * pre).to(torch.float16)          # the activated stream as the producing launch stores it
    Xd, Ad, W1d, W2d, b1d, b2d = X.to(dev), A.to(dev), W1.to(dev), W2.to(dev), b1.to(dev), b2.to(dev)   # keep every buffer alive
    rc = _lib().wv_op_resblock(P(Xd), P(Ad), P(W1d), P(k1), P(b1d), P(W2d), P(k2), P(b2d), B, T, C, pre,
                               P(out_raw), P(out_act), s_act, S())
    assert rc == 0, _lib().wv_last_error()
    torch.cuda.synchronize()

    def dw5(u, w, b):
        return F.conv1d(F.pad(u.transpose(1, 2), (4, 0)), w, b, groups=C).transpose(1, 2)

    x = X.double()
    a = A.double()
    h = F.elu(dw5(a @ W1.double().t(), dw1.double(), b1.double()))
    ref = (dw5(h @ W2.double().t(), dw2.double(), b2.double()) + x).float()
    # error budget: fp16 rounding of a, of both staged GEMM tiles and of h, the half2 tap sums, and the
    # <= 1e-3 absolute error of the half2 ELU on its exponential branch, pushed through |W2| and the taps
    mag = (dw5((h.abs() @ W2.double().abs().t()), dw2.double().abs(), None)).float()
    if out_raw is not None:
        half_close(out_raw.cpu(), ref, rel=2 ** -9, abs_=6e-3, mid=mag * 0.5)
    if out_act is not None:
        half_close(out_act.cpu(), F.elu(ref * s_act), rel=2 ** -9, abs_=6e-3, mid=mag * 0.5)
