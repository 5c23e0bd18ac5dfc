"""Builds libflk.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

The .so is git-ignored but travels to the GPU box with the gpurun snapshot.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libflk.so")
SOURCES = ["flk_kernels.cu", "flk_slots.cu", "flk_api.cu", "flk_dist.cu"]
HEADERS = ["flk_device.cuh", "flk_kernels.cuh", "flk_step.cuh", "flk_slots.cuh", "flk_internal.cuh", os.path.join("..", "..", "include", "flk.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17", "--extended-lambda",
    # exact-f32 contract: no FMA contraction, IEEE div/sqrt, denormals kept
    "-fmad=false", "-prec-div=true", "-prec-sqrt=true", "-ftz=false",
    "-Xcompiler", "-fPIC,-fvisibility=hidden,-O2",
    "-shared",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libflk.so cannot be built (there is no CPU fallback)")


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    extra = os.environ.get("FLK_NVCC_EXTRA", "").split()   # tuning experiments only, e.g. -DFLK_SMEM_UNROLL=4
    out = os.environ.get("FLK_LIB_OUT", LIB)
    cmd = [_nvcc()] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-o", out] + SOURCES
    env = dict(os.environ)
    # the image exports CC=/opt/gcc/bin/gcc (no libgomp spec); nvcc's host compiler must be the system one
    res = subprocess.run(cmd + ["-ccbin", "/usr/bin/g++"] if os.path.exists("/usr/bin/g++") else cmd,
                         cwd=CSRC, env=env, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    if verbose:
        sys.stderr.write(res.stderr)
    return LIB


HOST_DIR = os.path.join(HERE, "host")
HOST_DEMO = os.path.join(HOST_DIR, "flockers_demo")


def build_host_demo(force: bool = False) -> str | None:
    """C++ host mirror of the krABMaga surface (examples_b200/host): compile its demo against libflk.so."""
    src = os.path.join(HOST_DIR, "flockers_main.cpp")
    if not os.path.exists(src):
        return None
    hdr = os.path.join(HOST_DIR, "krabmaga_gpu.hpp")
    if not force and os.path.exists(HOST_DEMO) and os.path.getmtime(HOST_DEMO) > max(os.path.getmtime(src), os.path.getmtime(hdr)):
        return HOST_DEMO
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    cmd = [cxx, "-O2", "-std=c++17", "-I", os.path.join(HERE, "..", "include"), "-o", HOST_DEMO, src,
           "-L", HERE, "-lflk", "-Wl,-rpath,$ORIGIN/.."]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("host demo build failed:\n" + res.stdout + res.stderr)
    return HOST_DEMO


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
