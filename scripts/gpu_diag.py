"""First-contact diagnostics on a B200: per-kernel and end-to-end errors vs torch / golden.
Prints numbers, asserts nothing (the pytest suite carries the tolerances)."""
import ctypes as C
import os
import sys
import time
import traceback

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
from waveverify_b200 import _lib, Generator, Detector, Locator  # noqa: E402
from helpers import golden_cases, load_case, fixture_weights, BASE_KW, snr_db  # noqa: E402

dev = torch.device("cuda:0")
L = _lib.lib()
P = lambda t: None if t is None else C.c_void_p(t.data_ptr())
S = lambda: C.c_void_p(torch.cuda.current_stream().cuda_stream)


def gemm_case(M, N, K, bias, res, act, fp16=False):
    g = torch.Generator(device="cpu").manual_seed(M + N + K)
    dt = torch.float16 if fp16 else torch.float16
    lda = (K + 7) // 8 * 8
    A = torch.zeros(M, lda, dtype=dt); A[:, :K] = torch.randn(M, K, generator=g).to(dt)
    W = torch.zeros(N, lda, dtype=dt); W[:, :K] = (torch.randn(N, K, generator=g) / K ** 0.5).to(dt)
    b = torch.randn(N, generator=g) if bias else None
    R = torch.randn(M, N, generator=g).to(torch.float16) if res else None
    A, W = A.to(dev), W.to(dev)
    b = b.to(dev) if bias else None
    R = R.to(dev) if res else None
    out = torch.empty(M, N, dtype=torch.float16, device=dev)
    outa = torch.empty(M, N, dtype=torch.float16, device=dev) if act else None
    rc = L.wv_op_gemm(P(A), lda, P(W), lda, M, N, K, P(b), P(R), P(out), P(outa), 0.8, int(fp16), S())
    torch.cuda.synchronize()
    if rc != 0:
        return f"rc={rc} {L.wv_last_error().decode()}"
    ref = A[:, :K].float() @ W[:, :K].float().t()
    if bias: ref = ref + b
    if res: ref = ref + R.float()
    e = (out.float() - ref).abs().max().item()
    s = f"max|err| {e:.4f} (ref max {ref.abs().max().item():.2f})"
    if act:
        ra = torch.nn.functional.elu(ref * 0.8)
        s += f" act err {(outa.float() - ra).abs().max().item():.4f}"
    return s


def main():
    print(torch.cuda.get_device_name(0), "lib version", L.wv_version())
    for (M, N, K, bias, res, act) in [(128, 64, 64, False, False, False), (1000, 64, 64, True, True, True),
                                      (4096, 256, 256, False, False, False), (300, 96, 192, True, False, True),
                                      (777, 1536, 128, False, True, False), (5000, 128, 33, False, True, True),
                                      (20000, 96, 96, False, False, False), (333, 768, 1536, True, False, False),
                                      (64000, 192, 192, False, True, True)]:
        try:
            print(f"gemm M={M} N={N} K={K} bias={bias} res={res} act={act}:", gemm_case(M, N, K, bias, res, act), flush=True)
        except Exception:
            traceback.print_exc()
    print("gemm fp16:", gemm_case(1000, 128, 128, False, False, False, fp16=True), flush=True)

    import wv_oracle as O
    from helpers import oracle_cfg
    for path in golden_cases():
        z = load_case(path)
        name = os.path.basename(path)[:-4]
        zi, ws = bool(z["zero_init"]), int(z["wseed"])
        try:
            mods = {}
            for kind, cls in (("generator", Generator), ("detector", Detector), ("locator", Locator)):
                c, sd = fixture_weights(kind, zi, ws)
                m = cls(**{**BASE_KW[kind], "bias": True, "zero_init": zi})
                m.load_state_dict(sd)
                mods[kind] = m.to(dev)
            x = torch.from_numpy(z["x"]).to(dev); msg = torch.from_numpy(z["msg"]).to(dev)
            t0 = time.time()
            wm, y, lat = mods["generator"].embed_batch(x, msg, want_latent=True)
            torch.cuda.synchronize()
            print(f"[{name}] G: {time.time()-t0:.3f}s launches {mods['generator'].launches(*z['x'].shape[::2])} ws {mods['generator'].workspace_bytes()/1e6:.1f}MB", flush=True)
            e_wm = np.abs(wm.cpu().numpy() - z["wm"]).max()
            print(f"[{name}] wm max|err| {e_wm:.5f} snr {snr_db(z['wm'], wm.cpu().numpy()):.1f} dB  (wm rms {np.sqrt((z['wm']**2).mean()):.4f}); latent max|err| {np.abs(lat.cpu().numpy()-z['latent']).max():.4f} snr {snr_db(z['latent'], lat.cpu().numpy()):.1f} dB", flush=True)
            yg = torch.from_numpy(z["y"]).to(dev)
            d = mods["detector"].detect_batch(yg, want_logits=True)
            torch.cuda.synchronize()
            dd = int(z["det_decim"])
            lg = d["logits"][:, :, ::dd].cpu().numpy()
            print(f"[{name}] D logits max|err| {np.abs(lg - z['det_logits_decim']).max():.4f} snr {snr_db(z['det_logits_decim'], lg):.1f} dB; avg max|err| {np.abs(d['avg'].cpu().numpy()-z['det_avg']).max():.5f}; bits mismatch {(d['bits'].cpu().numpy()!=z['det_bits']).sum()} of {z['det_bits'].size}; min margin {np.abs(z['det_avg']-0.5).min():.5f}", flush=True)
            l = mods["locator"].locate_batch(yg, want_logits=True, want_probs=True)
            torch.cuda.synchronize()
            ll = l["logits"].cpu().numpy(); mk = l["mask"].cpu().numpy()
            mism = mk != z["loc_mask"]
            print(f"[{name}] L logits max|err| {np.abs(ll - z['loc_logits']).max():.4f} snr {snr_db(z['loc_logits'], ll):.1f} dB; mask mismatch {mism.sum()} of {mism.size}; max |ref-0.5| among mismatches {np.abs(z['loc_logits'][mism]-0.5).max() if mism.any() else 0:.4f}", flush=True)
        except Exception:
            traceback.print_exc()


if __name__ == "__main__":
    main()
