"""CPU oracle for the WaveVerify embed/detect/locate hot path.

TEST INFRASTRUCTURE ONLY.  This is a plain torch-on-CPU functional restatement of the
reference algorithm (fp32 by default, fp64 on request).  Only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs may
import it; the product package `waveverify_b200` never does.

Parity pinning: the reference ships no tests / golden vectors for this path (SURVEY
F2), so this oracle is pinned against OUTPUTS OF THE REFERENCE ITSELF, run in the build
container by `oracle/make_golden.py` (reference modules imported unmodified from
/root/reference) and committed under `tests/golden/`.  `tests/test_oracle_golden.py`
re-checks the oracle against those fixtures everywhere (no reference needed).

Every function cites the reference file:line it restates (paths relative to the
reference root).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor

WAV_STD = 0.1122080159                                  # modules/seanet.py:631, 1044
SPEC_MEANS = [-4.554, -4.315, -4.021, -3.726, -3.477]  # modules/seanet.py:632
SPEC_STDS = [2.830, 2.837, 2.817, 2.796, 2.871]        # modules/seanet.py:633

BASE_CFG = {
    # conf/base.yml:5-112 with the SURVEY F5 fix-ups (bias=True; Locator has no nbits)
    "generator": dict(channels_enc=64, channels_dec=96, n_residual_enc=2, n_residual_dec=3,
                      strides=[8, 5, 4, 2], dimension=128, n_fft_base=64, res_scale=3 ** -0.5,
                      nbits=16, embedding_dim=64, embedding_layers=2, freq_bands=4),
    "detector": dict(channels_enc=64, n_residual_enc=2, strides=[8, 5, 4, 2], dimension=128,
                     n_fft_base=64, res_scale=3 ** -0.5, output_dim=32, nbits=16),
    "locator": dict(channels_enc=32, n_residual_enc=1, strides=[8, 4], dimension=64,
                    n_fft_base=64, res_scale=3 ** -0.5, output_dim=32, nbits=1),
}


# --------------------------------------------------------------------------------------
# weight folding
# --------------------------------------------------------------------------------------
def fold_state_dict(sd: Dict[str, Tensor], dtype=torch.float32) -> Dict[str, Tensor]:
    """Collapse re-parametrised conv weights into plain `<prefix>.weight` tensors.

    weight-norm (modules/conv.py:47-88 -> torch parametrizations.weight_norm, dim=0):
        W[o] = g[o] * v[o] / ||v[o]||_2   (norm over all dims but 0; same for ConvTranspose1d)
    weight-standardisation (modules/weight_standardization.py:108-147):
        W = g * scale * (v - mean) / sqrt(max(var * fan_in, 1e-7))
    Plain `weight` keys (checkpoints with parametrizations removed, scripts/train.py:1624-1629)
    pass through unchanged.
    """
    out: Dict[str, Tensor] = {}
    for k, v in sd.items():
        if k.endswith(".parametrizations.weight.original1"):
            p = k[: -len(".parametrizations.weight.original1")]
            g = sd[p + ".parametrizations.weight.original0"].to(torch.float64)
            vv = v.to(torch.float64)
            n = vv.flatten(1).norm(dim=1).view(-1, *([1] * (vv.dim() - 1)))
            out[p + ".weight"] = (g * vv / n).to(dtype)
        elif k.endswith(".parametrizations.weight.original0"):
            continue
        elif k.endswith(".weight_v"):
            p = k[: -len(".weight_v")]
            vv = v.to(torch.float64)
            g = sd[p + ".weight_g"].to(torch.float64)
            scale = sd.get(p + ".weight_scale")
            scale = 1.0 if scale is None else scale.to(torch.float64)
            flat = vv.flatten(1)
            mean = flat.mean(dim=1).view(-1, *([1] * (vv.dim() - 1)))
            var = flat.var(dim=1, unbiased=False).view(-1, *([1] * (vv.dim() - 1)))
            fan_in = flat.shape[1]
            out[p + ".weight"] = (g * scale * (vv - mean) / torch.sqrt(
                torch.clamp(var * fan_in, min=1e-7))).to(dtype)
        elif k.endswith(".weight_g") or k.endswith(".weight_scale"):
            continue
        else:
            out[k] = v.to(dtype) if v.is_floating_point() else v
    return out


# --------------------------------------------------------------------------------------
# primitive ops
# --------------------------------------------------------------------------------------
def causal_conv1d(x: Tensor, w: Tensor, b: Optional[Tensor], stride: int = 1,
                  groups: int = 1) -> Tensor:
    """SConv1d.forward with causal=True, dilation=1 (modules/conv.py:715-763) incl. the
    extra right padding of get_extra_padding_for_conv1d (modules/conv.py:160-203)."""
    k = w.shape[-1]
    pad_total = (k - 1) - (stride - 1)
    T = x.shape[-1]
    n_frames = (T - k + pad_total) / stride + 1
    ideal = (math.ceil(n_frames) - 1) * stride + (k - pad_total)
    extra = max(0, ideal - T)
    x = F.pad(x, (pad_total, extra))
    return F.conv1d(x, w, b, stride=stride, groups=groups)


def causal_convtr1d_dw(x: Tensor, w: Tensor, stride: int) -> Tensor:
    """SConvTranspose1d.forward, causal, trim_right_ratio=1, depthwise, no bias
    (modules/conv.py:838-881): full transposed conv then drop the last k - stride samples."""
    C = x.shape[1]
    k = w.shape[-1]
    y = F.conv_transpose1d(x, w, None, stride=stride, groups=C)
    return y[..., : y.shape[-1] - (k - stride)]


def elu(x: Tensor) -> Tensor:
    return F.elu(x, alpha=1.0)


def resblock(x: Tensor, W: Dict[str, Tensor], p: str, idx: int, res_scale: float) -> Tensor:
    """SEANetResnetBlock.forward (modules/seanet.py:245-281); block layout from
    dws_conv_block (modules/seanet.py:85-109): [ELU, 1x1 (no bias), dw k5 (bias)] x 2."""
    h = x * (1.0 + idx * res_scale ** 2) ** -0.5
    for a, d in ((1, 2), (4, 5)):
        h = elu(h)
        h = F.conv1d(h, W[f"{p}.block.{a}.conv.conv.weight"])
        wd = W[f"{p}.block.{d}.conv.conv.weight"]
        h = causal_conv1d(h, wd, W.get(f"{p}.block.{d}.conv.conv.bias"), groups=wd.shape[0])
    scale = res_scale
    rsp = W.get(f"{p}.res_scale_param")
    if rsp is not None:
        scale = scale * rsp
    return h * scale + x


def causal_stft_mag(wav: Tensor, weight: Tensor, hop: int) -> Tensor:
    """CausalSTFT.forward (modules/conv.py:1036-1080): left-pad n_fft-1 zeros, conv1d with the
    [(n_fft+2), 1, n_fft] DFT*hann matrix at stride hop, magnitude with clamp 1e-12."""
    n_fft = weight.shape[-1]
    xp = F.pad(wav, (n_fft - 1, 0))
    c = F.conv1d(xp, weight, None, stride=hop)
    B, C2, Fr = c.shape
    c = c.view(B, 2, C2 // 2, Fr)
    return c.square().sum(dim=1).clamp_min(1e-12).sqrt()


def spec_branch(x: Tensor, wav: Tensor, W: Dict[str, Tensor], p: str, hop: int, mean: float,
                std: float, res_scale: float, taps: Optional[dict] = None, tap_name: str = "") -> Tensor:
    """SpecBlock.forward (modules/seanet.py:463-507), compression='log', inout_norm."""
    y = causal_stft_mag(wav, W[f"{p}.spec.weight"], hop)
    y = y.clamp_min(1e-5).log()
    y = (y - mean) / std
    if taps is not None:
        taps[tap_name + "_y"] = y
    y = F.conv1d(y, W[f"{p}.layer.conv.conv.weight"])
    scale = res_scale
    sp = W.get(f"{p}.scale_param")
    if sp is not None:
        scale = sp * scale
    return x + y * scale


def msg_embedding(msg: Tensor, W: Dict[str, Tensor], p: str, n_layers: int) -> Tensor:
    """msg_embedding MLP (modules/seanet.py:830-839): Linear, then n x (Linear, ReLU)."""
    e = F.linear(msg, W[f"{p}.0.weight"], W[f"{p}.0.bias"])
    for i in range(n_layers):
        j = 1 + 2 * i
        e = F.relu(F.linear(e, W[f"{p}.{j}.weight"], W[f"{p}.{j}.bias"]))
    return e


def film_table(e: Tensor, W: Dict[str, Tensor], p: str, n_scales: int, bands: int) -> Tensor:
    """gamma/beta scalars of FiLM.forward (modules/seanet.py:535-550) -> [B, scales, bands, 2]."""
    out = e.new_zeros(e.shape[0], n_scales, bands, 2)
    for s in range(n_scales):
        for b in range(bands):
            q = f"{p}.{s}.{b}"
            out[:, s, b, 0] = F.linear(e, W[q + ".gamma_layer.weight"], W[q + ".gamma_layer.bias"])[:, 0]
            out[:, s, b, 1] = F.linear(e, W[q + ".beta_layer.weight"], W[q + ".beta_layer.bias"])[:, 0]
    return out


# --------------------------------------------------------------------------------------
# encoder / decoder
# --------------------------------------------------------------------------------------
def encoder_forward(wav: Tensor, msg: Optional[Tensor], W: Dict[str, Tensor], cfg: dict,
                    p: str = "encoder", taps: Optional[dict] = None) -> Tensor:
    """SEANetEncoder.forward (modules/seanet.py:883-976)."""
    ratios = list(reversed(cfg["strides"]))          # seanet.py:646
    n_res = cfg["n_residual_enc"]
    rs = cfg["res_scale"]
    x = wav * (1.0 / WAV_STD)                           # Scale, seanet.py:658
    w = W[f"{p}.conv_pre.1.conv.conv.weight"]
    x = causal_conv1d(x, w, W.get(f"{p}.conv_pre.1.conv.conv.bias"))
    film = None
    if msg is not None:
        e = msg_embedding(msg.to(wav.dtype), W, f"{p}.msg_embedding", cfg.get("embedding_layers", 2))
        film = film_table(e, W, f"{p}.film_layers", len(ratios), cfg.get("freq_bands", 4))
        if taps is not None:
            taps["film"] = film
    hop = 1
    for s, r in enumerate(ratios):
        for j in range(1, n_res + 1):                   # idx = j because spec != "" (seanet.py:684)
            x = resblock(x, W, f"{p}.blocks.{s}.{j - 1}", j, rs)
        if taps is not None:
            taps[f"enc_s{s}_res"] = x
        x = spec_branch(x, wav, W, f"{p}.spec_blocks.{s}", hop, SPEC_MEANS[s], SPEC_STDS[s], rs, taps, f"enc_s{s}")
        if taps is not None:
            taps[f"enc_s{s}_spec"] = x
        hop *= r
        # downsample: Scale, ELU, 1x1 (no bias), strided depthwise (bias)  seanet.py:745-771
        x = elu(x * (1.0 + n_res * rs ** 2) ** -0.5)
        x = F.conv1d(x, W[f"{p}.downsample.{s}.2.conv.conv.weight"])
        wd = W[f"{p}.downsample.{s}.3.conv.conv.weight"]
        x = causal_conv1d(x, wd, W.get(f"{p}.downsample.{s}.3.conv.conv.bias"), stride=r,
                          groups=wd.shape[0])
        if film is not None:                            # seanet.py:928-966
            B, C, _ = x.shape
            nb = film.shape[2]
            if C % nb != 0:
                raise ValueError("channels must be divisible by freq_bands")
            g = film[:, s, :, 0].repeat_interleave(C // nb, dim=1).unsqueeze(-1)
            b = film[:, s, :, 1].repeat_interleave(C // nb, dim=1).unsqueeze(-1)
            x = x * g + b
        if taps is not None:
            taps[f"enc_s{s}"] = x
    x = spec_branch(x, wav, W, f"{p}.spec_post", hop, SPEC_MEANS[-1], SPEC_STDS[-1], rs, taps, "enc_post")
    if taps is not None:
        taps["enc_post_spec"] = x
    # conv_post: ELU, dw k5 (no bias), 1x1 (bias), L2Norm * sqrt(dim)   seanet.py:797-823
    x = elu(x)
    wd = W[f"{p}.conv_post.1.conv.conv.weight"]
    x = causal_conv1d(x, wd, None, groups=wd.shape[0])
    x = F.conv1d(x, W[f"{p}.conv_post.2.conv.conv.weight"], W.get(f"{p}.conv_post.2.conv.conv.bias"))
    x = F.normalize(x, p=2.0, dim=1, eps=1e-12) * (cfg["dimension"] ** 0.5)   # seanet.py:288-318
    if taps is not None:
        taps["enc_latent"] = x
    return x


def decoder_forward(z: Tensor, W: Dict[str, Tensor], cfg: dict, p: str = "decoder.model",
                    taps: Optional[dict] = None) -> Tensor:
    """SEANetDecoder.forward (modules/seanet.py:1212-1227); layer list built at :1067-1204."""
    ratios = cfg["strides"]
    n_res = cfg["n_residual_dec"]
    rs = cfg["res_scale"]
    i = 0
    x = F.conv1d(z, W[f"{p}.{i}.conv.conv.weight"]); i += 1
    wd = W[f"{p}.{i}.conv.conv.weight"]
    x = causal_conv1d(x, wd, W.get(f"{p}.{i}.conv.conv.bias"), groups=wd.shape[0]); i += 1
    stage_scale = (1.0 + n_res * rs ** 2) ** -0.5
    for si, r in enumerate(ratios):
        if si > 0:
            x = x * stage_scale
        x = elu(x)
        i += 2                                           # scale_layer, act
        x = causal_convtr1d_dw(x, W[f"{p}.{i}.convtr.convtr.weight"], r); i += 1
        x = F.conv1d(x, W[f"{p}.{i}.conv.conv.weight"], W.get(f"{p}.{i}.conv.conv.bias")); i += 1
        for j in range(n_res):
            x = resblock(x, W, f"{p}.{i}", j, rs); i += 1
        if taps is not None:
            taps[f"dec_u{si}"] = x
    x = elu(x * stage_scale)
    i += 2
    x = causal_conv1d(x, W[f"{p}.{i}.conv.conv.weight"], W.get(f"{p}.{i}.conv.conv.bias"))
    x = x * WAV_STD
    return torch.tanh(x)


# --------------------------------------------------------------------------------------
# models
# --------------------------------------------------------------------------------------
def generator_forward(x: Tensor, msg: Tensor, W: Dict[str, Tensor], cfg: dict,
                      taps: Optional[dict] = None) -> Tensor:
    """Generator.forward (model/generator.py:360-423): returns the watermark RESIDUAL [B,1,T]."""
    T = x.shape[-1]
    z = encoder_forward(x, msg, W, cfg, "encoder", taps)
    if taps is not None:
        taps["latent"] = z
    wm = decoder_forward(z, W, cfg, "decoder.model", taps)
    return wm[..., :T]


def embed(x: Tensor, msg: Tensor, W: Dict[str, Tensor], cfg: dict) -> Tuple[Tensor, Tensor]:
    """AudioWatermarking._forward_audio_sample (model/watermarking.py:423-441)."""
    wm = generator_forward(x, msg, W, cfg)
    return wm, wm + x


def _head(z: Tensor, W: Dict[str, Tensor], T: int) -> Tensor:
    u = F.conv_transpose1d(z, W["reverse_convolution.weight"], W["reverse_convolution.bias"],
                           stride=W["reverse_convolution.weight"].shape[-1])
    u = u[:, :, :T]
    return F.conv1d(u, W["last_layer.weight"], W["last_layer.bias"])


def detector_forward(y: Tensor, W: Dict[str, Tensor], cfg: dict) -> Tensor:
    """Detector.forward/decode (model/detector.py:366-391, 278-318) -> logits [B,nbits,T]."""
    return _head(encoder_forward(y, None, W, cfg, "encoder"), W, y.shape[-1])


def locator_forward(y: Tensor, W: Dict[str, Tensor], cfg: dict, taps: Optional[dict] = None) -> Tensor:
    """Locator.forward/decode (model/locator.py:268-299, 228-265) -> logits [B,1,T]."""
    return _head(encoder_forward(y, None, W, cfg, "encoder", taps), W, y.shape[-1])


# --------------------------------------------------------------------------------------
# decode / metrics
# --------------------------------------------------------------------------------------
def decode_bits(logits: Tensor, mask: Optional[Tensor] = None, thr: float = 0.5):
    """waveverify/core.py:577-586 + waveverify/utils.py:385-401 (unmasked) and
    scripts/evaluate.py:471-494 (masked).  Returns (bits u8 [B,W], avg [B,W], conf [B], valid [B,W])."""
    p = torch.sigmoid(logits)
    B, Wb, T = logits.shape
    if mask is None:
        avg = p.mean(dim=2)
        valid = torch.ones(B, Wb, dtype=torch.bool)
    else:
        m = mask.to(p.dtype).expand(-1, Wb, -1)
        valid = m.sum(dim=2) > 0
        avg = (p * m).sum(dim=2) / (m.sum(dim=2) + 1e-8)
    bits = (avg >= thr).to(torch.uint8)
    return bits, avg, avg.mean(dim=1), valid


def detector_postprocess(logits: Tensor, thr: float = 0.5) -> Tensor:
    """Detector.postprocess (model/detector.py:320-364): softmax over bits -> mean_t -> sigmoid -> > thr."""
    r = torch.softmax(logits, dim=1).mean(dim=-1)
    return torch.gt(torch.sigmoid(r), thr).int()


def locator_mask(logits: Tensor) -> Tensor:
    """model/watermarking.py:717, 797: `locator_out > 0.5` on RAW logits."""
    return (logits > 0.5).to(torch.uint8)


def metric_counters(bits: Tensor, valid: Tensor, msg: Tensor, pred_mask: Tensor,
                    gt_mask: Tensor) -> List[int]:
    """Six exact integer counters whose ratios give BER (scripts/evaluate.py:498-505) and
    mIoU (scripts/evaluate.py:636-656): [bit_errors, valid_bits, I_fg, U_fg, I_bg, U_bg]."""
    err = ((bits != msg.to(bits.dtype)) & valid).sum().item()
    nv = valid.sum().item()
    p = pred_mask.bool(); g = gt_mask.bool()
    return [int(err), int(nv), int((p & g).sum()), int((p | g).sum()),
            int((~p & ~g).sum()), int((~p | ~g).sum())]


def ber_miou_from_counters(c: Sequence[int]) -> Tuple[float, float]:
    ber = c[0] / c[1] if c[1] > 0 else 0.0
    iou_fg = c[2] / c[3] if c[3] > 0 else 1.0
    iou_bg = c[4] / c[5] if c[5] > 0 else 1.0
    return ber, 0.5 * (iou_fg + iou_bg)
