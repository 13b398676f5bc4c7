"""Build libegot2.so (hand-written sm_100a CUDA + the C ABI of include/egot2.h) in-tree with nvcc.

    python -m egot2_b200.build            # incremental
    python -m egot2_b200.build --force
    EGOT2_BUILD_TAG=timeline EGOT2_CFLAGS=-DEGOT2_TIMELINE python -m egot2_b200.build    # -> lib/libegot2_timeline.so

The library lands in egot2_b200/lib/libegot2.so (git-ignored, shipped to the GPU box by gpurun).
"""
from __future__ import annotations

import concurrent.futures as cf
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
# EGOT2_BUILD_TAG=<tag>: an instrumented / experimental variant (e.g. EGOT2_CFLAGS=-DEGOT2_TIMELINE) gets its own object
# directory and library name (lib/libegot2_<tag>.so, loaded with EGOT2_LIB=...), so the product build is never disturbed
_TAG = os.environ.get("EGOT2_BUILD_TAG", "")
OBJDIR = os.path.join(HERE, "build" + ("_" + _TAG if _TAG else ""))
LIB = os.path.join(LIBDIR, "libegot2" + ("_" + _TAG if _TAG else "") + ".so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
CFLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xcompiler", "-Wall", "--expt-relaxed-constexpr"]
CFLAGS += os.environ.get("EGOT2_CFLAGS", "").split()      # e.g. EGOT2_CFLAGS=-DEGOT2_FFN_TRACE for a debugging build


def sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _newest_header() -> float:
    t = 0.0
    for d in (CSRC, os.path.join(os.path.dirname(HERE), "include")):
        for f in os.listdir(d):
            if f.endswith((".h", ".cuh")):
                t = max(t, os.path.getmtime(os.path.join(d, f)))
    return t


def _compile(src: str, verbose: bool) -> str:
    obj = os.path.join(OBJDIR, src[:-3] + ".o")
    cmd = [NVCC, *ARCH, *CFLAGS, "-c", os.path.join(CSRC, src), "-o", obj]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
    if verbose:
        sys.stderr.write(r.stderr)
    return obj


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(LIBDIR, exist_ok=True)
    os.makedirs(OBJDIR, exist_ok=True)
    hdr = _newest_header()
    todo, objs = [], []
    for s in sources():
        obj = os.path.join(OBJDIR, s[:-3] + ".o")
        objs.append(obj)
        src_t = max(os.path.getmtime(os.path.join(CSRC, s)), hdr)
        if force or not os.path.exists(obj) or os.path.getmtime(obj) < src_t:
            todo.append(s)
    if todo:
        with cf.ThreadPoolExecutor(max_workers=min(8, len(todo))) as ex:
            list(ex.map(lambda s: _compile(s, verbose), todo))
    if todo or not os.path.exists(LIB):
        cmd = [NVCC, *ARCH, "-shared", "-o", LIB, *objs]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
