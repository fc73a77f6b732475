"""In-tree build of the CUDA library (sm_100a only) and of the host C++ programs.

  libxfb.so        csrc/xfb_row.cu + xfb_col.cu + xfb_api.cu  (the C ABI of include/xfb.h)

nvcc cross-compiles without a GPU; the .so is git-ignored but travels with gpurun snapshots.
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
ROOT = os.path.dirname(HERE)
BUILD = os.path.join(HERE, "_build")
LIB = os.path.join(HERE, "libxfb.so")

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xptxas", "-v"]
CUDA_SOURCES = ["xfb_row.cu", "xfb_col.cu", "xfb_api.cu"]
HEADERS = ["xfb_fft.cuh", "xfb_row.cuh", "xfb_col.cuh", "xfb_internal.h", os.path.join(ROOT, "include", "xfb.h")]


def _mtime(p):
    return os.path.getmtime(p) if os.path.exists(p) else 0.0


def _compile(src):
    obj = os.path.join(BUILD, os.path.splitext(src)[0] + ".o")
    log = os.path.join(BUILD, os.path.splitext(src)[0] + ".ptxas.log")
    deps = [os.path.join(CSRC, src)] + [h if os.path.isabs(h) else os.path.join(CSRC, h) for h in HEADERS]
    if _mtime(obj) >= max(_mtime(d) for d in deps):
        return obj
    r = subprocess.run([NVCC, *NVCC_FLAGS, "-c", os.path.join(CSRC, src), "-o", obj], capture_output=True, text=True)
    with open(log, "w") as fh:
        fh.write(r.stdout + r.stderr)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError(f"nvcc failed on {src}")
    return obj


def build_library(force: bool = False) -> str:
    os.makedirs(BUILD, exist_ok=True)
    if force:
        for s in CUDA_SOURCES:
            o = os.path.join(BUILD, os.path.splitext(s)[0] + ".o")
            if os.path.exists(o):
                os.remove(o)
    with ThreadPoolExecutor(max_workers=len(CUDA_SOURCES)) as ex:
        objs = list(ex.map(_compile, CUDA_SOURCES))
    if _mtime(LIB) < max(_mtime(o) for o in objs):
        r = subprocess.run([NVCC, "-shared", "-o", LIB, *objs, "-gencode", "arch=compute_100a,code=sm_100a"],
                           capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("link failed")
    return LIB


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv))
