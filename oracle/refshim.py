"""ORACLE -- TEST INFRASTRUCTURE ONLY.

Import the *live reference* (/root/reference/src/fluids2d) inside the build
container, stubbing the two host-only dependencies the image lacks
(matplotlib, netCDF4).  Used by ``tests/golden/make_golden.py`` to write the
golden vectors and by the optional in-container cross-checks.  The reference
tree does not exist on the GPU box; callers must test ``available()`` first.
"""
import os
import sys
import types

REF_SRC = "/root/reference/src"
# where oracle/install_ref.py puts the UNMODIFIED reference package (git-ignored,
# travels to the GPU box with the snapshot): bench.py --impl reference runs it
VENDORED = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "baseline", "_ref")


def available():
    return os.path.isdir(os.path.join(REF_SRC, "fluids2d"))


def vendored_available():
    return os.path.isdir(os.path.join(VENDORED, "fluids2d"))


def load(src=None):
    """Return the reference's ``fluids2d`` package (imports numba kernels, ~8 s).
    `src`: directory that holds the package (default: the live tree)."""
    src = src or REF_SRC
    if not os.path.isdir(os.path.join(src, "fluids2d")):
        raise RuntimeError("reference package not present under " + src)

    def stub(name, **attrs):
        if name in sys.modules:
            return sys.modules[name]
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    class _Dataset:
        def __init__(self, *a, **k):
            raise RuntimeError("netCDF4 stub: history output is disabled in the oracle harness")

    stub("netCDF4", Dataset=_Dataset)
    mpl = stub("matplotlib", rc=lambda *a, **k: None)
    mpl.pyplot = stub("matplotlib.pyplot")
    stub("mpl_toolkits")
    stub("mpl_toolkits.axes_grid1", make_axes_locatable=lambda ax: None)
    stub("PIL", Image=None)
    if src not in sys.path:
        sys.path.insert(0, src)
    import fluids2d  # noqa: E402
    return fluids2d
