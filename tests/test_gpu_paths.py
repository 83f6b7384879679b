"""GPU: the alternative code paths behind the tuning knobs must agree with the default path.

The knobs are read once per process (init_device_once), so each configuration runs in a subprocess
that prints a digest of one embed -> detect -> locate pass; the digests are compared here.  Bound:
the paths differ only in fp16 rounding order (fused vs separate launches), so the watermark must
agree to >= 60 dB and the decoded bits exactly on this fixture; bit-identical paths are compared
bit for bit."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SCRIPT = r"""
import json, sys, os
import numpy as np, torch
sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, ROOT)
from helpers import BASE_KW, fixture_weights
from waveverify_b200 import Detector, Generator, Locator
dev = torch.device("cuda:0")
mods = {}
for kind, cls in (("generator", Generator), ("detector", Detector), ("locator", Locator)):
    c, sd = fixture_weights(kind, False, 5)
    m = cls(**{**BASE_KW[kind], "bias": True, "zero_init": False}); m.load_state_dict(sd); mods[kind] = m.to(dev)
rng = np.random.RandomState(11)
x = torch.from_numpy(0.1 * rng.standard_normal((3, 1, 20011)).astype(np.float32)).to(dev)
msg = torch.from_numpy(rng.randint(0, 2, (3, 16)).astype(np.float32)).to(dev)
wm, y, _ = mods["generator"].embed_batch(x, msg)
d = mods["detector"].detect_batch(y, want_logits=True)
l = mods["locator"].locate_batch(y, want_logits=True)
torch.cuda.synchronize()
np.savez(sys.argv[1], wm=wm.cpu().numpy(), logits=d["logits"].cpu().numpy(), bits=d["bits"].cpu().numpy(),
         avg=d["avg"].cpu().numpy(), ll=l["logits"].cpu().numpy())
"""


def run(tmp_path, name, env):
    out = str(tmp_path / f"{name}.npz")
    e = dict(os.environ)
    e.update(env)
    r = subprocess.run([sys.executable, "-c", f"ROOT={ROOT!r}\n" + SCRIPT, out], env=e, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    z = np.load(out)
    return {k: z[k] for k in z.files}


def snr(a, b):
    a = a.astype(np.float64); b = b.astype(np.float64)
    return 10 * np.log10((a ** 2).sum() / max(((a - b) ** 2).sum(), 1e-300))


def test_knob_paths_agree(tmp_path):
    base = run(tmp_path, "default", {})
    # bit-identical alternatives: math-warp organisation and residual recomputation do not change any arithmetic
    for name, env in (("groups1", {"WV_MATH_GROUPS": "1"}), ("groups2", {"WV_MATH_GROUPS": "2"}),
                      ("rows4", {"WV_ROWS6_BN": "0"}), ("pre", {"WV_PRE_FUSE": "1"}), ("prefetch", {"WV_A_PREFETCH": "4"}),
                      ("res_late", {"WV_RES_EARLY2": "0"}), ("epi4", {"WV_EPI_GROUPS": "4"}), ("res1", {"WV_RES1_KB": "128"}),
                      ("res_ldg", {"WV_RES_TMA": "0"}), ("res_tma3", {"WV_RES_TMA_MIN_STAGES": "3"}),
                      ("no_cta_pairs", {"WV_CG2_MIN_KB": "0", "WV_CG2_BN96": "0"}), ("no_graph", {"WV_GRAPH_MAX_SAMPLES": "0"})):
        alt = run(tmp_path, name, env)
        for k in base:
            assert np.array_equal(base[k], alt[k]), f"{name}: {k} differs"
    # alternatives that change the fusion level (rounding points move): tight tolerance
    for name, env in (("nospecfuse", {"WV_SPEC_FUSE_MAXC": "0"}), ("specfuse512", {"WV_SPEC_FUSE_MAXC": "512"}),
                      ("lastconv_cuda", {"WV_LAST_GEMM": "0"}), ("noupfuse", {"WV_UP_FUSE_MAXC": "0"}),
                      ("nopair", {"WV_PAIR_MIN_KB": "0"})):
        alt = run(tmp_path, name, env)
        assert snr(base["wm"], alt["wm"]) >= 55.0, f"{name}: wm {snr(base['wm'], alt['wm']):.1f} dB"
        assert snr(base["logits"], alt["logits"]) >= 50.0, name
        assert snr(base["ll"], alt["ll"]) >= 50.0, name
        assert np.abs(base["avg"] - alt["avg"]).max() <= 2e-4, name
        safe = np.abs(base["avg"] - 0.5) > 3e-4
        assert (base["bits"] == alt["bits"])[safe].all(), name


def test_cta_pair_path_agrees_at_a_size_where_it_is_selected(tmp_path):
    """The CTA-pair launches need >= 2 tile pairs per SM pair: 48 x 1 s selects them for the deep stages.  Same arithmetic
    per element (only the tile shapes change): the Generator's output is bit-identical with the mode switched off."""
    script = r"""
import sys, os
import numpy as np, torch
sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, ROOT)
from helpers import BASE_KW, fixture_weights
from waveverify_b200 import Generator
c, sd = fixture_weights("generator", False, 5)
m = Generator(**{**BASE_KW["generator"], "bias": True, "zero_init": False}); m.load_state_dict(sd); m = m.cuda()
rng = np.random.RandomState(12)
x = torch.from_numpy(0.1 * rng.standard_normal((48, 1, 16000)).astype(np.float32)).cuda()
msg = torch.from_numpy(rng.randint(0, 2, (48, 16)).astype(np.float32)).cuda()
wm, y, _ = m.embed_batch(x, msg)
torch.cuda.synchronize()
np.savez(sys.argv[1], wm=wm.cpu().numpy())
"""
    outs = {}
    for name, env in (("on", {}), ("off", {"WV_CG2_MIN_KB": "0", "WV_CG2_BN96": "0"})):
        out = str(tmp_path / f"cg2_{name}.npz")
        e = dict(os.environ); e.update(env)
        r = subprocess.run([sys.executable, "-c", f"ROOT={ROOT!r}\n" + script, out], env=e, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr[-2000:]
        outs[name] = np.load(out)["wm"]
    assert np.array_equal(outs["on"], outs["off"])
