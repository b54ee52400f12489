"""The three kernel wrappers of the reference (src/fluids2d/weno.py:412-446),
same signatures, running on the device: host arrays are copied in, one CUDA
kernel is launched, the first argument is copied back.  They exist for
script-level compatibility and per-kernel parity tests; the time step itself
never round-trips through the host."""
from ._cabi import Engine, METHODS

_engine = None


def bind(engine):
    """kernels run on this engine's device/stream (set by Mesh)"""
    global _engine
    _engine = engine


def _e():
    if _engine is None:
        raise RuntimeError("no device engine bound: create a Model first")
    return _engine


def compflux(flx, U, q, o, s, funcname, nthreads=1):
    _e().compflux(flx, U, q, o, s, funcname)


def vortexforce(du, V, omega, o, s, s2, sign, funcname, nthreads=1):
    _e().vortexforce(du, V, omega, o, s, s2, sign, funcname)


def innerproduct(ke, U, u, o, s, funcname, nthreads=1):
    _e().innerproduct(ke, U, u, o, s, funcname)


class _Registry(dict):
    """CompFlux[m](flx, U, q, o, s, i0, i1)-style access (weno.py:439-446)."""

    def __init__(self, kind):
        super().__init__()
        for m in ("weno", "upwind", "centered", "cweno"):
            self[m] = self._make(kind, m)

    @staticmethod
    def _make(kind, m):
        def call(first, a, b, o, *args):
            shape = first.shape
            if kind == "compflux":
                s, i0, i1 = args
                assert (i0, i1) == (0, first.size), "partial intervals are not supported"
                _e().compflux(first, a, b, o, s, m)
            elif kind == "vortexforce":
                s, s2, sign, i0, i1 = args
                assert (i0, i1) == (0, first.size), "partial intervals are not supported"
                _e().vortexforce(first, a, b, o, s, s2, sign, m)
            else:
                s, i0, i1 = args
                assert (i0, i1) == (0, first.size), "partial intervals are not supported"
                _e().innerproduct(first, a, b, o, s, m)
            assert first.shape == shape
        return call


CompFlux = _Registry("compflux")
VortexForce = _Registry("vortexforce")
InnerProduct = _Registry("innerproduct")
