"""Compile libuce_b200.so (sm_100a) in-tree with nvcc.  No JIT cache: the .so travels with the tree.

Every source is compiled to its own object (in parallel, only when it or a header is newer) and the objects are linked into one
shared library: a one-file change rebuilds in seconds instead of a minute."""
from __future__ import annotations

import concurrent.futures
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libuce_b200.so")
SOURCES = ["uce_api.cu", "artifact.cu", "png.cu", "factor.cu", "factor_small.cu", "apply.cu", "apply_tc3.cu", "apply_ab.cu", "apply_gemm3x.cu",
           "unet_gemm.cu", "unet_ops.cu", "unet_attn.cu", "unet_engine.cu", "vae_engine.cu", "clip_text.cu", "clip_vision.cu"]
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
CFLAGS = ARCH + ["-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC"]
LDFLAGS = ARCH + ["-shared", "-Xcompiler", "-fPIC", "-lcuda", "-lz"]


def _headers():
    inc = os.path.join(HERE, "..", "include")
    hs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))]
    hs += [os.path.join(inc, f) for f in os.listdir(inc) if f.endswith(".h")]
    return hs + [os.path.abspath(__file__)]


def _newest(paths):
    return max(os.path.getmtime(p) for p in paths if os.path.isfile(p))


def build(force: bool = False, verbose: bool = False) -> str:
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    t_hdr = _newest(_headers())
    todo = []
    for s in SOURCES:
        src, obj = os.path.join(CSRC, s), os.path.join(OBJ, s[:-3] + ".o")
        if force or not os.path.isfile(obj) or os.path.getmtime(obj) < max(os.path.getmtime(src), t_hdr):
            todo.append((src, obj))
    objs = [os.path.join(OBJ, s[:-3] + ".o") for s in SOURCES]
    if not todo and os.path.isfile(LIB) and os.path.getmtime(LIB) >= _newest(objs):
        return LIB
    if not os.path.isfile(nvcc):
        raise RuntimeError("nvcc not found: cannot build libuce_b200.so")
    os.makedirs(OBJ, exist_ok=True)

    def one(job):
        src, obj = job
        cmd = [nvcc] + CFLAGS + ["-c", src, "-o", obj]
        if verbose:
            print(" ".join(cmd), file=sys.stderr)
        r = subprocess.run(cmd, cwd=CSRC, capture_output=True, text=True)
        return src, r

    with concurrent.futures.ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 4)) as ex:
        for src, r in ex.map(one, todo):
            if r.stderr.strip() and (verbose or r.returncode):
                print(r.stderr, file=sys.stderr)
            if r.returncode:
                raise RuntimeError(f"nvcc failed for {src}")
    cmd = [nvcc] + LDFLAGS + objs + ["-o", LIB]
    if verbose:
        print(" ".join(cmd), file=sys.stderr)
    subprocess.run(cmd, check=True, cwd=CSRC)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
