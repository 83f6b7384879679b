"""Shared helpers for the parity tests."""
import glob
import os

import numpy as np
import torch

from waveverify_b200 import params as P

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")

BASE_KW = {
    "generator": dict(), "detector": dict(),
    "locator": dict(dimension=64, channels_enc=32, n_residual_enc=1, strides=[8, 4]),
}


def net_config(kind, zero_init):
    return P.config_from_kwargs(kind, {**BASE_KW[kind], "bias": True, "zero_init": bool(zero_init)})


def oracle_cfg(c):
    return dict(strides=list(c.strides), n_residual_enc=c.n_residual_enc,
                n_residual_dec=c.n_residual_dec, res_scale=c.res_scale, dimension=c.dimension,
                embedding_layers=c.embedding_layers, freq_bands=c.freq_bands)


def golden_cases():
    return sorted(glob.glob(os.path.join(GOLDEN, "*.npz")))


def load_case(path):
    z = np.load(path)
    return {k: z[k] for k in z.files}


def fixture_weights(kind, zero_init, seed):
    c = net_config(kind, zero_init)
    return c, P.fixture_state_dict(c, int(seed))


def snr_db(ref, test):
    ref = np.asarray(ref, np.float64); test = np.asarray(test, np.float64)
    n = ((ref - test) ** 2).sum()
    return 10 * np.log10((ref ** 2).sum() / max(n, 1e-300))
