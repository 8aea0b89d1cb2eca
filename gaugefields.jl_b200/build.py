"""Builds libgfb200.so (hand-written CUDA for sm_100a + the C ABI) in-tree with nvcc.

    python gaugefields.jl_b200/build.py [--force]

nvcc cross-compiles without a GPU; the .so is git-ignored but travels to the GPU box.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
SOURCES = ["kernels.cu", "tmarch.cu", "stout.cu", "primitives.cu", "general.cu", "heatbath.cu", "api.cu"]
HEADERS = ["su3.cuh", "lattice.cuh", "stencil.cuh", "tmarch_geom.h", "gfb_internal.h", os.path.join("..", "..", "include", "gfb200.h")]
# GFB200_VARIANT=<name> + GFB200_NVCC_EXTRA="-D..." build a tuning variant next to the default library
VARIANT = os.environ.get("GFB200_VARIANT", "")
LIB = os.path.join(HERE, "libgfb200%s.so" % (("_" + VARIANT) if VARIANT else ""))
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "-ccbin", "/usr/bin/g++",
]


def _stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force=False, verbose=False):
    if not force and not _stale():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    objs = []
    procs = []
    bdir = os.path.join(HERE, "build", VARIANT or "default")
    os.makedirs(bdir, exist_ok=True)
    extra = os.environ.get("GFB200_NVCC_EXTRA", "").split()
    for s in SOURCES:
        src = os.path.join(CSRC, s)
        if not os.path.exists(src):
            continue
        obj = os.path.join(bdir, s.replace(".cu", ".o"))
        cmd = [nvcc] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
        procs.append((cmd, subprocess.Popen(cmd)))
        objs.append(obj)
    for cmd, p in procs:
        if p.wait() != 0:
            raise RuntimeError("nvcc failed: " + " ".join(cmd))
    link = [nvcc, "-shared", "-Wno-deprecated-gpu-targets", "-ccbin", "/usr/bin/g++", "-o", LIB] + objs + ["-lnccl"]
    subprocess.check_call(link)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
