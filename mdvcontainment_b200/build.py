"""In-tree build of libmdvc.so (hand-written sm_100a CUDA behind the C-ABI of include/mdvc.h).

nvcc cross-compiles without a GPU; the shared object lands next to this file so it travels to
the GPU box with the repository snapshot (it is git-ignored, not gpurun-ignored).
"""
from __future__ import annotations

import os
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
LIB = os.path.join(PKG, "libmdvc.so")
SOURCES = ("bits.cu", "frame.cu", "gridops.cu", "hostgraph.cpp")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC,-O2", "--use_fast_math=false"]


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(PKG, "..", "include", "mdvc.h")]
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build_library(force: bool = False, verbose: bool = False, out: str | None = None) -> str:
    """``out``: build a variant (with MDVC_NVCC_EXTRA defines) into another file instead of libmdvc.so."""
    if out is None and not force and not needs_build():
        return LIB
    objs = []
    flags = [f for f in NVCC_FLAGS if not f.startswith("--use_fast_math")]
    flags += os.environ.get("MDVC_NVCC_EXTRA", "").split()          # e.g. -DSTRIP_TY=128 for tuning experiments
    for src in SOURCES:
        obj = os.path.join(CSRC, os.path.splitext(src)[0] + (".o" if out is None else ".variant.o"))
        cmd = [_nvcc(), *flags, "-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
            print(" ".join(cmd), file=sys.stderr)
        subprocess.check_call(cmd)
        objs.append(obj)
    cmd = [_nvcc(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", *objs, "-o", out or LIB, "-lcudart"]
    subprocess.check_call(cmd)
    return out or LIB


if __name__ == "__main__":
    out = sys.argv[sys.argv.index("--out") + 1] if "--out" in sys.argv else None
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv, out=out))
