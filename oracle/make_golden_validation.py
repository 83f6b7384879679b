"""Generate tests/golden/validation/validation.npz from the REFERENCE ITSELF (run in the build container only).

The reference's utils/localization_augmentation.py, utils/seq_augmentation.py and
utils/effect_augmentation.py are loaded unmodified by path.  Their module-level imports of
matplotlib, julius and audiotools (absent here; never called by the functions exercised) are
satisfied by empty stand-in modules.  Inputs are regenerated from seeds; the npz stores the seeds,
the reference's outputs and (for the noise / suppression effects) the reference's random draws.
"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

REF = os.environ.get("WV_REFERENCE_ROOT", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "validation", "validation.npz")
SR = 200          # small "sample rate": segment_length = 20 samples, shuffle segments of 100


def _stub(name, **attrs):
    if name in sys.modules:
        return
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m


def load_ref():
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__))))
    import ref_import
    ref_import.load()                                  # installs the audiotools stand-in
    _stub("matplotlib")
    _stub("matplotlib.pyplot")
    _stub("matplotlib.patches")
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.modules["matplotlib"].patches = sys.modules["matplotlib.patches"]
    _stub("julius")
    import logging
    logging.disable(logging.CRITICAL)
    mods = {}
    for name in ("localization_augmentation", "seq_augmentation", "effect_augmentation"):
        spec = importlib.util.spec_from_file_location(f"_wvref_{name}", os.path.join(REF, "utils", f"{name}.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        mods[name] = mod
    return mods


def inputs(seed, B, T):
    r = np.random.RandomState(seed)
    x = (0.1 * r.standard_normal((B, 1, T))).astype(np.float32)
    y = (x + 0.01 * r.standard_normal((B, 1, T))).astype(np.float32)
    return x, y


def main():
    mods = load_ref()
    Loc = mods["localization_augmentation"].LocalizationAugmentation
    Seq = mods["seq_augmentation"].SequenceAugmentation
    FX = mods["effect_augmentation"]
    out = {"sample_rate": np.int64(SR)}
    # localization: (seed, B, T); B = 1 has no cross substitution
    loc_cases = [(11, 3, 437), (12, 4, 1000), (13, 1, 450), (14, 2, 219), (15, 5, 2013)]
    out["loc_cases"] = np.array(loc_cases, np.int64)
    for i, (seed, B, T) in enumerate(loc_cases):
        x, y = inputs(seed, B, T)
        np.random.seed(seed)
        sig, gt, upd, stats = Loc(SR, 0.1)(torch.from_numpy(x), torch.from_numpy(y))
        out[f"loc{i}_wm"] = sig.audio_data.numpy()
        out[f"loc{i}_gt"] = gt.numpy().astype(np.uint8)
        out[f"loc{i}_upd"] = upd.numpy()
        out[f"loc{i}_stats"] = np.array([stats[k] for k in ("original_revert", "zero_replace", "cross_substitute", "unchanged")])
    # sequence: scan seeds until every reachable method is covered twice, at two lengths
    seq_cases = []
    seen = {}
    for T in (437, 150, 1000):
        for seed in range(100, 160):
            np.random.seed(seed)
            r = np.random.rand()
            m = "reverse" if r < 0.3 else "circular_shift" if r < 0.7 else "shuffle"
            if seen.get((T, m), 0) >= (2 if T != 150 else 1):
                continue
            seen[(T, m)] = seen.get((T, m), 0) + 1
            seq_cases.append((seed, 2, T))
    out["seq_cases"] = np.array(seq_cases, np.int64)
    methods = []
    for i, (seed, B, T) in enumerate(seq_cases):
        x, y = inputs(seed, B, T)
        gt = (np.random.RandomState(seed + 1).rand(B, 1, T) < 0.7).astype(np.float32)
        np.random.seed(seed)
        torch.manual_seed(seed)
        sig, upd, gto, stats, method = Seq(SR)(torch.from_numpy(x), torch.from_numpy(y), torch.from_numpy(gt))
        out[f"seq{i}_wm"] = sig.audio_data.numpy()
        out[f"seq{i}_upd"] = upd.numpy()
        out[f"seq{i}_gt"] = gto.numpy().astype(np.uint8)
        methods.append(method)
    out["seq_methods"] = np.array(methods)
    # effects through the reference's apply_effect
    x, _ = inputs(21, 3, 437)
    xt = torch.from_numpy(x)
    out["fx_scale"] = FX.apply_effect(xt.clone(), "amplitude_scaling", sample_rate=SR, scale=0.5)[0].numpy()
    for bd in (8, 16):
        out[f"fx_quant{bd}"] = FX.apply_effect(xt.clone(), "quantization", sample_rate=SR, bit_depth=bd)[0].numpy()
    for k in (3, 5, 9):
        out[f"fx_median{k}"] = FX.apply_effect(xt.clone(), "median_filter", sample_rate=SR, kernel_size=k)[0].numpy()
    torch.manual_seed(5)
    noise = torch.randn_like(xt)
    torch.manual_seed(5)
    out["fx_noise_draw"] = noise.numpy()
    out["fx_noise"] = FX.apply_effect(xt.clone(), "random_noise", sample_rate=SR, noise_std=0.01)[0].numpy()
    torch.manual_seed(6)
    idx = torch.stack([torch.randperm(437)[:43] for _ in range(3)])
    torch.manual_seed(6)
    mask = torch.ones_like(xt)
    ya, ma = FX.apply_effect(xt.clone(), "sample_suppression", sample_rate=SR, mask=mask, suppression_percentage=0.1)
    out["fx_suppress_idx"] = idx.numpy()
    out["fx_suppress"] = ya.numpy()
    out["fx_suppress_mask"] = ma.numpy().astype(np.uint8)
    np.savez_compressed(OUT, **out)
    print("wrote", os.path.normpath(OUT), os.path.getsize(OUT), "bytes; sequence methods:", methods)


if __name__ == "__main__":
    main()
