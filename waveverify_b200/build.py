"""Build the sm_100a shared library in-tree (nvcc cross-compiles without a GPU)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "csrc", "wv_b200.cu")
DEPS = [os.path.join(HERE, "csrc", f) for f in ("wv_b200.cu", "gemm_sm100.cuh", "glue_kernels.cuh", "ptx_sm100.cuh", "validation_kernels.cuh", "resblock_sm100.cuh")]
DEPS.append(os.path.join(os.path.dirname(HERE), "include", "wv_b200.h"))
LIB = os.path.join(HERE, "libwv_b200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
    "-shared", "-Xcompiler", "-fPIC", "-Xptxas", "-v", "-lcudart",
]


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(d) > t for d in DEPS)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    extra = ["-DWV_TIMELINE"] if os.environ.get("WV_TIMELINE_BUILD") else []   # scripts/timeline.py probes
    out = os.environ.get("WV_LIB_OUT") or LIB          # A/B and probe builds go to their own file (loaded via WV_LIB_PATH)
    cmd = [nvcc] + NVCC_FLAGS + extra + ["-o", out, SRC]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed building libwv_b200.so")
    if out == LIB:
        with open(os.path.join(HERE, "csrc", "ptxas.log"), "w") as f:
            f.write(r.stdout + r.stderr)
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
