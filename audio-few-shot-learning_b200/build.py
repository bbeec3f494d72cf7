"""Build libafsl.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

    python -m afsl_b200.build            # or: python audio-few-shot-learning_b200/build.py

The shared object lands next to this file so that it travels with the source
tree (it is git-ignored, not gpurun-ignored).  Nothing is JIT-compiled at import.
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
INCLUDE = os.path.join(os.path.dirname(PKG_DIR), "include")
LIB_PATH = os.path.join(PKG_DIR, "libafsl.so")
OBJ_DIR = os.path.join(PKG_DIR, "build")

# IEEE fp32 math throughout (1e-5 parity bar): no --use_fast_math
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _digest(path: str) -> str:
    h = hashlib.sha256()
    for dep in [path] + sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cuh")) + \
            [os.path.join(INCLUDE, "afsl.h")]:
        with open(dep, "rb") as fh:
            h.update(fh.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def _compile(src: str, verbose: bool) -> str:
    os.makedirs(OBJ_DIR, exist_ok=True)
    obj = os.path.join(OBJ_DIR, os.path.basename(src)[:-3] + ".o")
    stamp = obj + ".sha"
    want = _digest(src)
    if os.path.exists(obj) and os.path.exists(stamp) and open(stamp).read() == want:
        return obj
    cmd = [_nvcc()] + NVCC_FLAGS + ["-I", INCLUDE, "-I", CSRC, "-c", src, "-o", obj]
    if verbose:
        print(" ".join(cmd), flush=True)
    subprocess.run(cmd, check=True)
    with open(stamp, "w") as fh:
        fh.write(want)
    return obj


def build(force: bool = False, verbose: bool = True) -> str:
    """Compile every csrc/*.cu for sm_100a and link libafsl.so; returns its path."""
    if force:
        for f in os.listdir(OBJ_DIR) if os.path.isdir(OBJ_DIR) else []:
            os.remove(os.path.join(OBJ_DIR, f))
    srcs = sources()
    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as pool:
        objs = list(pool.map(lambda s: _compile(s, verbose), srcs))
    newest = max(os.path.getmtime(o) for o in objs)
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < newest:
        cmd = [_nvcc(), "-shared", "-o", LIB_PATH] + objs + ["-gencode", "arch=compute_100a,code=sm_100a",
                                                               "-Xcompiler", "-fPIC", "-lcudart"]
        if verbose:
            print(" ".join(cmd), flush=True)
        subprocess.run(cmd, check=True)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
