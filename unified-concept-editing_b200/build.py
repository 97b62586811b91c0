"""Compile libuce_b200.so (sm_100a) in-tree with nvcc.  No JIT cache: the .so travels with the tree."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libuce_b200.so")
SOURCES = ["uce_api.cu", "artifact.cu", "png.cu", "factor.cu", "factor_small.cu", "apply.cu", "apply_tc.cu", "apply_tc2.cu", "apply_tc3.cu", "apply_gemm3x.cu", "apply_gemm3x_ss.cu",
           "unet_gemm.cu", "unet_ops.cu", "unet_attn.cu", "unet_engine.cu", "vae_engine.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared", "-lcuda", "-lz"]


def _stale() -> bool:
    if not os.path.isfile(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", h) for h in ("uce_b200.h", "sd_unet_b200.h", "sd_vae_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps if os.path.isfile(d))


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not _stale():
        return LIB
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.isfile(nvcc):
        raise RuntimeError("nvcc not found: cannot build libuce_b200.so")
    cmd = [nvcc] + NVCC_FLAGS + [os.path.join(CSRC, s) for s in SOURCES] + ["-o", LIB]
    if verbose:
        print(" ".join(cmd), file=sys.stderr)
    subprocess.run(cmd, check=True, cwd=CSRC)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
