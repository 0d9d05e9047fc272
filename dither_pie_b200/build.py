"""Build libditherpie_b200.so in-tree with nvcc for sm_100a.

    python -m dither_pie_b200.build [--force] [--verbose]

nvcc cross-compiles without a GPU; the .so is git-ignored but travels to the GPU box.
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")
LIB = os.path.join(HERE, "libditherpie_b200.so")
OBJ_DIR = os.path.join(HERE, "_obj")

NVCC_FLAGS = [
    "-O3", "-std=c++17", "-lineinfo",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-Xcompiler", "-fPIC",
    "--fmad=false",          # every float multiply/add is separately rounded unless fmaf() is spelled
    "-I", INCLUDE,
]


def sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _newest_header() -> float:
    t = os.path.getmtime(os.path.join(INCLUDE, "ditherpie_b200.h"))
    for f in os.listdir(CSRC):
        if f.endswith(".cuh"):
            t = max(t, os.path.getmtime(os.path.join(CSRC, f)))
    return t


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJ_DIR, exist_ok=True)
    hdr_t = _newest_header()
    objs = []
    procs = []
    for src in sources():
        s = os.path.join(CSRC, src)
        o = os.path.join(OBJ_DIR, src[:-3] + ".o")
        objs.append(o)
        if (not force and os.path.exists(o)
                and os.path.getmtime(o) >= max(os.path.getmtime(s), hdr_t)):
            continue
        cmd = ["nvcc", *NVCC_FLAGS, "-c", s, "-o", o]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
            print(" ".join(cmd))
        procs.append((src, subprocess.Popen(cmd)))
    failed = [src for src, p in procs if p.wait() != 0]
    if failed:
        raise RuntimeError(f"nvcc failed for {failed}")
    if (force or procs or not os.path.exists(LIB)
            or any(os.path.getmtime(o) > os.path.getmtime(LIB) for o in objs)):
        cmd = ["nvcc", "-shared", "-gencode", "arch=compute_100a,code=sm_100a",
               "-o", LIB, *objs, "-lcudart", "-ldl"]
        if verbose:
            print(" ".join(cmd))
        subprocess.check_call(cmd)
    return LIB


def build_timing() -> str:
    """Debug variant for tools/wave_timing.py: the wavefront kernel with per-phase clock64()
    counters (-DDP_WAVE_TIMING).  Written next to the objects, never loaded by the product."""
    build()
    out = os.path.join(OBJ_DIR, "libditherpie_b200_timing.so")
    timed = {"dp_diffusion.cu": "-DDP_WAVE_TIMING", "dp_kmeans.cu": "-DDP_KM_TIMING"}   # tools/km_timing.py
    tobjs = []
    for src, flag in timed.items():
        tobjs.append(os.path.join(OBJ_DIR, src[:-3] + "_timing.o"))
        subprocess.check_call(["nvcc", *NVCC_FLAGS, flag, "-c", os.path.join(CSRC, src), "-o", tobjs[-1]])
    others = [os.path.join(OBJ_DIR, s[:-3] + ".o") for s in sources() if s not in timed]
    subprocess.check_call(["nvcc", "-shared", "-gencode", "arch=compute_100a,code=sm_100a",
                           "-o", out, *tobjs, *others, "-lcudart", "-ldl"])
    return out


if __name__ == "__main__":
    if "--timing" in sys.argv:
        print(build_timing())
    else:
        print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
