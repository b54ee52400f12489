"""Compile libf2d.so (and the bit-exact variant) in-tree with nvcc for sm_100a.

    python -m fluids2d_b200.build            # both libraries
    python -m fluids2d_b200.build --force

libf2d.so        production build: FMA contraction on, one-division WENO weights
libf2d_exact.so  -fmad=false -DF2D_EXACT: every floating-point operation in the
                 reference's order (used by the bit-exact per-kernel tests)
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
SOURCES = ["api.cu", "ops.cu", "step.cu", "mg.cu", "dist.cu"]
HEADERS = ["engine.cuh", "weno.cuh", "reduce.cuh", "mg_tiles.cuh", os.path.join("..", "..", "include", "f2d.h")]
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]


def lib_path(exact=False):
    return os.path.join(HERE, "libf2d_exact.so" if exact else "libf2d.so")


def _stale(target):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build_one(exact=False, force=False, verbose=False):
    out = lib_path(exact)
    if not force and not _stale(out):
        return out
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc, "-O3", "-std=c++17", "-lineinfo", "-shared", "-Xcompiler", "-fPIC",
           "-cudart", "static", *ARCH]
    if exact:
        cmd += ["-fmad=false", "-DF2D_EXACT"]
    if verbose:
        cmd += ["-Xptxas", "-v"]
    cmd += ["-o", out] + [os.path.join(CSRC, s) for s in SOURCES] + ["-ldl"]
    subprocess.check_call(cmd)
    return out


def build_all(force=False, verbose=False):
    return [build_one(False, force, verbose), build_one(True, force, verbose)]


if __name__ == "__main__":
    for p in build_all(force="--force" in sys.argv, verbose="-v" in sys.argv):
        print(p)
