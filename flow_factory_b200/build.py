"""Compile libffb200.so (sm_100a only) in-tree with nvcc.  No JIT cache, no fallback."""
from __future__ import annotations

import hashlib
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libffb200.so")
STAMP = LIB + ".srchash"
SOURCES = ["ffb200.cu", "gemm.cu", "attention.cu", "attention_d128.cu", "elementwise.cu", "final_step.cu", "engine.cu", "flux_engine.cu", "vae_conv.cu", "vae_elementwise.cu", "vae_engine.cu", "wan_elementwise.cu", "wan_engine.cu", "common.cuh", "kernels.h", "softmax.cuh"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-shared",
              "-Xcompiler", "-fPIC"]


def _nvcc() -> str:
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if c and (os.path.isabs(c) and os.path.exists(c) or not os.path.isabs(c)):
            return c
    return "nvcc"


def source_hash() -> str:
    """sha256 over the compiler flags and every source the unity build includes (names + contents)."""
    h = hashlib.sha256(" ".join(NVCC_FLAGS).encode())
    deps = [os.path.join(CSRC, s) for s in SOURCES] + [os.path.join(os.path.dirname(HERE), "include", "ffb200.h")]
    for d in sorted(deps):
        h.update(os.path.basename(d).encode())
        with open(d, "rb") as f:
            h.update(f.read())
    return h.hexdigest()


def needs_build() -> bool:
    """True unless libffb200.so was built from exactly the sources in the tree: the hash of the sources is stamped next to the library at
    build time (no mtime trust - a stale or foreign .so is rebuilt, 20 s)."""
    if not os.path.exists(LIB) or not os.path.exists(STAMP):
        return True
    with open(STAMP) as f:
        return f.read().strip() != source_hash()


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    cmd = [_nvcc(), *NVCC_FLAGS, "-o", LIB, os.path.join(CSRC, "ffb200.cu")]
    if verbose:
        print(" ".join(cmd))
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + r.stdout + r.stderr)
    with open(STAMP, "w") as f:
        f.write(source_hash() + "\n")
    return LIB


if __name__ == "__main__":
    print(build(force=True, verbose=True))
