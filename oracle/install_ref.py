"""ORACLE -- TEST / BENCH INFRASTRUCTURE ONLY.

Install the UNMODIFIED reference package into ``baseline/_ref`` (git-ignored, NOT
gpurun-ignored: it travels to the GPU box) so that ``bench.py --impl reference``
times the reference itself (numba kernels + numpy + scipy SuperLU) on the box's
host cores.  Run in the build container, where /root/reference exists:

    python oracle/install_ref.py

First choice is the contract's offline pip install.  The reference's build backend
(hatchling, pyproject.toml:1-3) is not in the image's wheelhouse, so pip cannot
build the wheel here; the package is pure Python in hatch's ``src/`` layout, and
the wheel would contain exactly the directory ``src/fluids2d`` -- the fallback
unpacks that directory as the wheel would.  Nothing of it enters the git history.
"""
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
DEST = os.path.join(ROOT, "baseline", "_ref")


def install(verbose=True):
    if not os.path.isdir(os.path.join(REF, "src", "fluids2d")):
        return "reference tree not present (GPU box): using what baseline/_ref already holds"
    if os.path.isdir(os.path.join(DEST, "fluids2d")):
        return "already installed"
    os.makedirs(DEST, exist_ok=True)
    tmp = "/tmp/f2d_refcopy"
    shutil.rmtree(tmp, ignore_errors=True)
    shutil.copytree(REF, tmp)                 # /root/reference is read-only; builds write into the tree
    cmd = [sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--no-deps",
           "--find-links", "/opt/wheelhouse", "--target", DEST, tmp]
    r = subprocess.run(cmd, capture_output=True, text=True)
    how = "pip install --no-index --no-deps --target baseline/_ref"
    if r.returncode != 0 or not os.path.isdir(os.path.join(DEST, "fluids2d")):
        # no hatchling offline: place the pure-Python package the way its wheel would
        shutil.copytree(os.path.join(tmp, "src", "fluids2d"), os.path.join(DEST, "fluids2d"), dirs_exist_ok=True)
        how = "pip could not build the wheel (hatchling missing offline); unpacked src/fluids2d as the wheel would"
    shutil.rmtree(tmp, ignore_errors=True)
    if verbose:
        print("baseline/_ref:", how)
    return how


if __name__ == "__main__":
    print(install())
