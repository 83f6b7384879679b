"""TEST INFRASTRUCTURE (build container only: reads /root/reference).  Times the UNMODIFIED reference modules
(model/generator.py:360, detector.py:366, locator.py:268 + the add at watermarking.py:440) against the oracle port that
bench.py uses as its CPU arm, same weights, same clips, same thread count; writes profiles/r02_reference_vs_port_cpu.md."""
import os
import sys
import time

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE); sys.path.insert(0, ROOT)
import ref_import  # noqa: E402
import wv_oracle as O  # noqa: E402

threads = int(os.environ.get("WV_THREADS", os.cpu_count() or 8))
torch.set_num_threads(threads)
torch.manual_seed(0)
g, d, l, AS = ref_import.build(False)
rows = []
for B in (1, 8):
    T = 16000
    rng = np.random.RandomState(1)
    x = torch.from_numpy((0.1 * rng.standard_normal((B, 1, T))).astype(np.float32))
    msg = torch.from_numpy(rng.randint(0, 2, (B, 16)).astype(np.int64))
    W = {k: O.fold_state_dict(m.state_dict()) for k, m in (("g", g), ("d", d), ("l", l))}
    cfg = {"g": dict(strides=[8, 5, 4, 2], n_residual_enc=2, n_residual_dec=3, res_scale=0.5773502691896258, dimension=128, embedding_layers=2, freq_bands=4),
           "d": dict(strides=[8, 5, 4, 2], n_residual_enc=2, n_residual_dec=3, res_scale=0.5773502691896258, dimension=128, embedding_layers=2, freq_bands=4),
           "l": dict(strides=[8, 4], n_residual_enc=1, n_residual_dec=3, res_scale=0.5773502691896258, dimension=64, embedding_layers=2, freq_bands=4)}

    def ref_step():
        with torch.no_grad():
            wm = g(AS(x, 16000), msg).audio_data
            y = x + wm
            lg = d(AS(y, 16000)); ll = l(AS(y, 16000))
            return (torch.sigmoid(lg).mean(-1) >= 0.5), ll > 0.5

    def port_step():
        with torch.no_grad():
            wm = O.generator_forward(x, msg, W["g"], cfg["g"])
            y = x + wm
            lg = O.detector_forward(y, W["d"], cfg["d"]); ll = O.locator_forward(y, W["l"], cfg["l"])
            return O.decode_bits(lg)[0], O.locator_mask(ll)

    def timeit(fn, n=3):
        fn()
        ts = []
        for _ in range(n):
            t0 = time.perf_counter(); fn(); ts.append(time.perf_counter() - t0)
        return min(ts)

    tr, tp = timeit(ref_step), timeit(port_step)
    rows.append((B, tr, tp))
    print(B, tr, tp, flush=True)
out = ["# r02: the unmodified reference modules against the oracle port on the CPU (build container, %d threads)\n" % threads,
       "`oracle/time_reference_vs_port.py`: embed -> detect -> locate on B x 1 s clips, fp32, best of 3 after one warm-up, same weights\n"
       "and inputs.  The GPU box has no copy of the reference, so `bench.py`'s CPU arm is the port; this table gives the factor.\n",
       "| clips | reference modules (s / step) | audio-s/s | oracle port (s / step) | audio-s/s | port / reference |", "|---|---|---|---|---|---|"]
for B, tr, tp in rows:
    out.append(f"| {B} | {tr:.3f} | {B / tr:.2f} | {tp:.3f} | {B / tp:.2f} | {tr / tp:.2f} x faster |")
out.append("\nThe port folds weight-norm once (the reference recomputes `g * v / ||v||` in every forward) and does not materialise a\n"
           "padded copy per convolution; the arithmetic is the same (`tests/test_oracle_golden.py`: <= 2e-6 against the reference's outputs).\n"
           "A GPU / port ratio therefore understates the GPU / reference ratio by the last column.\n")
open(os.path.join(ROOT, "profiles", "r02_reference_vs_port_cpu.md"), "w").write("\n".join(out))
print("\n".join(out))
