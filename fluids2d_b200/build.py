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
HEADERS = ["engine.cuh", "weno.cuh", "reduce.cuh", "mg_tiles.cuh", "tma.cuh", os.path.join("..", "..", "include", "f2d.h")]
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]


def lib_path(exact=False):
    return os.path.join(HERE, "libf2d_exact.so" if exact else "libf2d.so")


def _stale(target):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def _compile(nvcc, flags, src, obj):
    subprocess.check_call([nvcc, *flags, "-c", "-o", obj, src])
    return obj


def build_one(exact=False, force=False, verbose=False):
    """one object per translation unit (compiled in parallel, rebuilt only when the
    source or a header is newer), linked into the shared library"""
    out = lib_path(exact)
    if not force and not _stale(out):
        return out
    from concurrent.futures import ThreadPoolExecutor
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    flags = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", *ARCH]
    if exact:
        flags += ["-fmad=false", "-DF2D_EXACT"]
    if verbose:
        flags += ["-Xptxas", "-v"]
    objdir = os.path.join(HERE, "build", "exact" if exact else "prod")
    os.makedirs(objdir, exist_ok=True)
    hdr_t = max(os.path.getmtime(os.path.join(CSRC, h)) for h in HEADERS)
    hdr_t = max(hdr_t, os.path.getmtime(os.path.abspath(__file__)))
    jobs = []
    with ThreadPoolExecutor(max_workers=len(SOURCES)) as pool:
        objs = []
        for src in SOURCES:
            sp, ob = os.path.join(CSRC, src), os.path.join(objdir, src[:-3] + ".o")
            objs.append(ob)
            if force or not os.path.exists(ob) or os.path.getmtime(ob) < max(os.path.getmtime(sp), hdr_t):
                jobs.append(pool.submit(_compile, nvcc, flags, sp, ob))
        for j in jobs:
            j.result()
    subprocess.check_call([nvcc, "-shared", "-cudart", "static", *ARCH, "-o", out, *objs, "-ldl"])
    return out


def build_all(force=False, verbose=False):
    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(max_workers=2) as pool:
        jobs = [pool.submit(build_one, e, force, verbose) for e in (False, True)]
        return [j.result() for j in jobs]


if __name__ == "__main__":
    for p in build_all(force="--force" in sys.argv, verbose="-v" in sys.argv):
        print(p)
