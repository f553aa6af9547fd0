"""Build libcoopcap.so (hand-written sm_100a kernels + C ABI) in-tree with nvcc.

nvcc cross-compiles for sm_100a without a GPU; the resulting .so sits next to this file so it
travels with the repo snapshot to the GPU box (it is git-ignored, not gpurun-ignored).
"""
from __future__ import annotations

import concurrent.futures as cf
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
BUILD = os.path.join(HERE, "csrc", "build")
LIB = os.path.join(HERE, "libcoopcap.so")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
    "-I", INCLUDE, "-I", CSRC,
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; libcoopcap.so cannot be built")


def _sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _deps_mtime() -> float:
    m = 0.0
    for root in (CSRC, INCLUDE):
        for f in os.listdir(root):
            if f.endswith((".cuh", ".h")):
                m = max(m, os.path.getmtime(os.path.join(root, f)))
    return m


def _compile_one(nvcc, src, obj, verbose):
    cmd = [nvcc, *NVCC_FLAGS, "-c", os.path.join(CSRC, src), "-o", obj]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
    return src, r.stderr


def build_library(force: bool = False, verbose: bool = False) -> str:
    nvcc = _nvcc()
    os.makedirs(BUILD, exist_ok=True)
    hdr_m = _deps_mtime()
    jobs, objs = [], []
    for src in _sources():
        obj = os.path.join(BUILD, src[:-3] + ".o")
        objs.append(obj)
        src_m = max(os.path.getmtime(os.path.join(CSRC, src)), hdr_m)
        if force or not os.path.exists(obj) or os.path.getmtime(obj) < src_m:
            jobs.append((src, obj))
    if jobs:
        with cf.ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            for src, log in ex.map(lambda j: _compile_one(nvcc, j[0], j[1], verbose), jobs):
                if verbose and log:
                    print(f"--- {src}\n{log}", file=sys.stderr)
    if jobs or not os.path.exists(LIB):
        cmd = [nvcc, "-shared", "-cudart", "static", "-gencode", "arch=compute_100a,code=sm_100a",
               "-Xcompiler", "-fPIC", "-o", LIB, *objs]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
