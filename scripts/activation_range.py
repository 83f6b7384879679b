"""fp16 headroom: the largest |value| of every intermediate 16-bit tensor of the three nets (raw and
activated outputs of every launch, through wv_debug_tap) for full-scale inputs.  fp16 saturates at
65504; the report shows how far below that the path runs with the fixture weights."""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from waveverify_b200 import _lib  # noqa: E402

L = _lib.lib()
dev = torch.device("cuda:0")
mods = bench.make_models(dev)
B, T = 4, 16000
g = torch.Generator(device=dev).manual_seed(11)
inputs = {
    "0.1*randn (bench / golden inputs)": 0.1 * torch.randn(B, 1, T, device=dev, generator=g),
    "full-scale uniform noise |x|<=1": 2 * torch.rand(B, 1, T, device=dev, generator=g) - 1,
    "full-scale square wave 200 Hz": torch.sign(torch.sin(2 * 3.14159265 * 200 * torch.arange(T, device=dev) / 16000)).expand(B, 1, T).contiguous(),
}
msg = torch.randint(0, 2, (B, 16), device=dev).float()
buf = torch.empty(B * T * 1536 // 8, dtype=torch.float16, device=dev)
print("| input | net | largest \\|value\\| over all 16-bit tensors | where | headroom to 65504 |")
print("|---|---|---|---|---|")
for name, x in inputs.items():
    for kind, m in mods.items():
        m.set_profile(True)
        if kind == "generator":
            m.embed_batch(x, msg)
        elif kind == "detector":
            m.detect_batch(x)
        else:
            m.locate_batch(x)
        torch.cuda.synchronize()
        tags = [r["tag"] for r in m.profile_read()]
        m.set_profile(False)
        worst, where = 0.0, ""
        h = m._native().handle
        for tag in tags:
            for which in (0, 1):
                wr = C.c_size_t(0)
                rc = L.wv_debug_tap(h, C.c_void_p(x.data_ptr()), C.c_void_p(msg.data_ptr()) if kind == "generator" else None,
                                    B, T, tag.encode(), which, C.c_void_p(buf.data_ptr()), buf.numel() * 2, C.byref(wr))
                if rc != 0 or wr.value == 0 or "wav16" in tag or "film" in tag:
                    continue
                v = buf[: wr.value // 2].float()
                assert bool(torch.isfinite(v).all()), f"non-finite value in {kind} {tag}"
                mx = float(v.abs().max())
                if mx > worst:
                    worst, where = mx, f"{tag}[{which}]"
        print(f"| {name} | {kind} | {worst:.1f} | {where} | {65504 / max(worst, 1e-9):.0f}x |")
