"""Equation-set helpers that user scripts import (reference:
src/fluids2d/equations.py:4-6, 229-238).  The model right-hand sides themselves
are CUDA kernels (csrc/step.cu); this module only keeps the host-side hooks."""


def fill(mesh, *variables):
    for v in variables:
        mesh.fill(v)


def addforcingterm(param, mesh, rhs, forcing):
    """wrap a (device) rhs with a host forcing callback ``forcing(param, mesh, s, ds)``"""
    print("[INFO] add a forcing term")

    def newrhs(s, ds):
        rhs(s, ds)
        forcing(param, mesh, s, ds)

    newrhs.__doc__ = "\n".join([rhs.__doc__ or "", "with forcing term"])
    return newrhs
